// Oracle-only OpenCV stand-in: see ../opencv.hpp
#pragma once
#include "../opencv.hpp"
