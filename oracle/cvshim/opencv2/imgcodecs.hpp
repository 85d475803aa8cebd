// TEST INFRASTRUCTURE ONLY — see core/core.hpp.  common/include/common/BasicConfig.h names
// cv::imread in an inline member that the SSD translation unit never calls; a declaration suffices.
#pragma once
#include <opencv2/core/core.hpp>
namespace cv {
enum ImreadModes { IMREAD_UNCHANGED = -1 };
Mat imread(const std::string& path, int flags);
} // namespace cv
