// stand-in for <opencv2/opencv.hpp>: see cv_standin.hpp (TEST INFRASTRUCTURE)
#pragma once
#include "cv_standin.hpp"
