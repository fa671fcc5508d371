// stand-in for <opencv2/core/core.hpp>: see cv_standin.hpp (TEST INFRASTRUCTURE)
#pragma once
#include "cv_standin.hpp"
