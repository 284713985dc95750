// Collisions.h:5 / Forces.h:11 include "external\ArcSim\mesh.hpp" (Windows separator): forward to the reference's own header
#include "external/ArcSim/mesh.hpp"
