// stand-in for <opencv2/highgui/highgui.hpp>: see cv_standin.hpp (TEST INFRASTRUCTURE)
#pragma once
#include "cv_standin.hpp"
