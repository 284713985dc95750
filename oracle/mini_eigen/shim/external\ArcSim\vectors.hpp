// conversions.h:5 includes "external\ArcSim\vectors.hpp" (Windows separator): forward to the reference's own header
#include "external/ArcSim/vectors.hpp"
