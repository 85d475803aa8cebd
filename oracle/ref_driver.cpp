// TEST INFRASTRUCTURE ONLY (oracle/): C entry point around the reference's own, unmodified
// serial::disparitySSD (/root/reference/ProblemSets/ps2_cpp/lib/DisparitySSD.cpp:9-62), which the
// Makefile compiles in place against oracle/cvshim.  Built into oracle/_ref/libref_ssd.so; used by
// tests/ to pin the C restatement (oracle/stereo_oracle.c) and by bench.py's cpu_baseline /
// --impl reference leg.  Never linked into the product library.
#include <opencv2/core/core.hpp>
#include <opencv2/imgcodecs.hpp>
#include <spdlog/spdlog.h>
#include <spdlog/sinks/null_sink.h>

#include <cstring>
#include <mutex>

// The reference's header (ProblemSets/ps2_cpp/include/DisparitySSD.h:38-43)
namespace serial {
void disparitySSD(const cv::Mat& left, const cv::Mat& right, const size_t windowRad,
                  const int minDisparity, const int maxDisparity, cv::Mat& disparity);
}

namespace cv {
Mat imread(const std::string&, int) { return Mat(); }   // declared by the shim, never called
}

static void ensure_logger() {
    static std::once_flag once;
    std::call_once(once, [] {
        // DisparitySSD.cpp:17-18 dereferences spdlog::get("file_logger") unconditionally.
        auto sink = std::make_shared<spdlog::sinks::null_sink_mt>();
        spdlog::register_logger(std::make_shared<spdlog::logger>("file_logger", sink));
    });
}

extern "C" int ref_serial_disparity_ssd(const float* left, const float* right, int rows, int cols,
                                        int windowRad, int minDisparity, int maxDisparity,
                                        signed char* disparity_out) {
    if (!left || !right || !disparity_out || rows <= 0 || cols <= 0 || windowRad < 0) return 1;
    ensure_logger();
    cv::Mat l(rows, cols, CV_32FC1, const_cast<float*>(left), size_t(cols) * 4);
    cv::Mat r(rows, cols, CV_32FC1, const_cast<float*>(right), size_t(cols) * 4);
    cv::Mat d;
    serial::disparitySSD(l, r, size_t(windowRad), minDisparity, maxDisparity, d);
    for (int y = 0; y < rows; ++y)
        std::memcpy(disparity_out + size_t(y) * cols, d.data + size_t(y) * d.step, size_t(cols));
    return 0;
}
