// stand-in for <opencv2/features2d/features2d.hpp>: see cv_standin.hpp (TEST INFRASTRUCTURE)
#pragma once
#include "cv_standin.hpp"
