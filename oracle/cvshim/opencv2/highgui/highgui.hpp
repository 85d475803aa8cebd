// TEST INFRASTRUCTURE ONLY — see core/core.hpp.  DisparityNCorr.cpp:6 includes highgui and uses nothing of it.
#pragma once
#include <opencv2/core/core.hpp>
