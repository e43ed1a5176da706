// oracle/ref_mock: stand-in for the OpenCV header of this name (OpenCV C++ headers do not exist in this image).
// TEST INFRASTRUCTURE ONLY -- see cvmini.hpp.
#pragma once
#include "../../cvmini.hpp"
