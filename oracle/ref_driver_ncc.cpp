// TEST INFRASTRUCTURE ONLY (oracle/): C entry point around the reference's own, unmodified
// serial::disparityNCorr (/root/reference/ProblemSets/ps2_cpp/lib/DisparityNCorr.cpp:12-71), which the
// Makefile compiles in place against oracle/cvshim into oracle/_ref/libref_ncc.so.  The reference's loop
// runs verbatim; cv::matchTemplate / cv::minMaxLoc come from the shim (cvshim/opencv2/imgproc/imgproc.hpp:
// OpenCV's TM_CCORR_NORMED arithmetic restated, pinned separately against executed OpenCV).  Used by tests/
// to pin oracle/stereo_oracle.c's NCC restatement and by bench.py's NCC cpu_baseline leg.  Never linked into
// the product library.
#include <opencv2/core/core.hpp>
#include <opencv2/imgcodecs.hpp>
#include <spdlog/spdlog.h>
#include <spdlog/sinks/null_sink.h>

#include <cstring>
#include <mutex>

// The reference's header (ProblemSets/ps2_cpp/include/DisparityNCorr.h:39-44)
namespace serial {
void disparityNCorr(const cv::Mat& left, const cv::Mat& right, const size_t windowRad,
                    const int minDisparity, const int maxDisparity, cv::Mat& disparity);
}

namespace cv {
Mat imread(const std::string&, int) { return Mat(); }   // declared by the shim, never called
}

static void ensure_logger() {
    static std::once_flag once;
    std::call_once(once, [] {
        // DisparityNCorr.cpp:20-21 dereferences spdlog::get("file_logger") unconditionally.
        if (!spdlog::get("file_logger")) {
            auto sink = std::make_shared<spdlog::sinks::null_sink_mt>();
            spdlog::register_logger(std::make_shared<spdlog::logger>("file_logger", sink));
        }
    });
}

// Returns 0, 1 (bad argument) or 2 (the reference threw: a pixel without any candidate window makes
// cv::Mat::operator()(Rect) / matchTemplate fail in real OpenCV too).
extern "C" int ref_serial_disparity_ncorr(const float* left, const float* right, int rows, int cols,
                                          int windowRad, int minDisparity, int maxDisparity,
                                          signed char* disparity_out) {
    if (!left || !right || !disparity_out || rows <= 0 || cols <= 0 || windowRad < 0) return 1;
    ensure_logger();
    cv::Mat l(rows, cols, CV_32FC1, const_cast<float*>(left), size_t(cols) * 4);
    cv::Mat r(rows, cols, CV_32FC1, const_cast<float*>(right), size_t(cols) * 4);
    cv::Mat d;
    try {
        serial::disparityNCorr(l, r, size_t(windowRad), minDisparity, maxDisparity, d);
    } catch (const std::exception&) {
        return 2;
    }
    for (int y = 0; y < rows; ++y)
        std::memcpy(disparity_out + size_t(y) * cols, d.data + size_t(y) * d.step, size_t(cols));
    return 0;
}
