// Forces.cpp:23 includes "external\ArcSim\util.hpp" (Windows separator): forward to the reference's own header
#include "external/ArcSim/util.hpp"
