// stand-in for <opencv/cv.h>: see cv_standin.hpp (TEST INFRASTRUCTURE)
#pragma once
#include "cv_standin.hpp"
