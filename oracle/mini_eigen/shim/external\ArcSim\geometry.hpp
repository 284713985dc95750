// UtilEOL.cpp:3 includes "external\ArcSim\geometry.hpp" (Windows separator): forward to the reference's own header
#include "external/ArcSim/geometry.hpp"
