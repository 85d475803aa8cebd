// TEST INFRASTRUCTURE ONLY — see core/core.hpp.  The two OpenCV imgproc/core calls that
// /root/reference/ProblemSets/ps2_cpp/lib/DisparityNCorr.cpp makes (:60 matchTemplate, :64 minMaxLoc),
// so that the reference's NCC translation unit compiles in place, unmodified, into oracle/_ref/libref_ncc.so.
//
// What this pins and what it does not: the reference's OWN code — padding, template/search rectangles and
// their clamping (:44-53), result size (:57), the disparity formula (:67), the char store (:68) — runs
// verbatim.  The arithmetic INSIDE cv::matchTemplate belongs to OpenCV imgproc 3.4.1 (absent here in C++);
// it is restated below after modules/imgproc/src/templmatch.cpp (crossCorr + common_matchTemplate) and is
// itself pinned against executed OpenCV (python cv2.matchTemplate) by oracle/ncc_cv2.py and
// tests/golden/make_golden.py.
#pragma once
#include <opencv2/core/core.hpp>
#include <cfloat>
#include <cmath>

namespace cv {

enum TemplateMatchModes { TM_SQDIFF = 0, TM_SQDIFF_NORMED = 1, TM_CCORR = 2, TM_CCORR_NORMED = 3, TM_CCOEFF = 4, TM_CCOEFF_NORMED = 5 };

// TM_CCORR_NORMED, CV_32FC1 only.  result is (H-h+1) x (W-w+1) CV_32FC1.
//   num      = sum T*I, delivered by OpenCV as float32 (crossCorr writes a CV_32F plane);
//              here: the double sum rounded once to float32
//   templNorm= sqrt(sum T^2)               (double; templmatch.cpp: norm(templ, NORM_L2) path, templSum2)
//   wndSum2  = sum I^2 over the window     (double integral image in OpenCV)
//   t        = sqrt(max(wndSum2,0)) * templNorm, or 0 when wndSum2 <= min(0.5, 10*FLT_EPSILON*wndSum2)
//   |num| < t -> num/t ; |num| < 1.125 t -> +-1 ; else 0 ; stored as float32
inline void matchTemplate(const Mat& image, const Mat& templ, Mat& result, int method) {
    if (method != TM_CCORR_NORMED || image.type() != CV_32FC1 || templ.type() != CV_32FC1)
        throw std::invalid_argument("cvshim::matchTemplate: only TM_CCORR_NORMED on CV_32FC1");
    const int H = image.rows, W = image.cols, h = templ.rows, w = templ.cols;
    if (H < h || W < w) throw std::out_of_range("cvshim::matchTemplate: template larger than image");
    result.create(H - h + 1, W - w + 1, CV_32FC1);
    double templSum2 = 0;
    for (int r = 0; r < h; ++r)
        for (int c = 0; c < w; ++c) { const double v = templ.at<float>(r, c); templSum2 += v * v; }
    const double templNorm = std::sqrt(templSum2);
    for (int y = 0; y + h <= H; ++y)
        for (int x = 0; x + w <= W; ++x) {
            double acc = 0, wndSum2 = 0;
            for (int r = 0; r < h; ++r)
                for (int c = 0; c < w; ++c) {
                    const double iv = image.at<float>(y + r, x + c);
                    acc += double(templ.at<float>(r, c)) * iv;
                    wndSum2 += iv * iv;
                }
            double num = double(float(acc));
            const double diff2 = wndSum2 > 0 ? wndSum2 : 0;
            const double t = (diff2 <= std::fmin(0.5, 10 * double(FLT_EPSILON) * wndSum2)) ? 0 : std::sqrt(diff2) * templNorm;
            if (std::fabs(num) < t) num /= t;
            else if (std::fabs(num) < t * 1.125) num = num > 0 ? 1 : -1;
            else num = 0;
            result.at<float>(y, x) = float(num);
        }
}

// First minimum / first maximum in row-major scan order (core/src/minmax.cpp: strict comparisons).
inline void minMaxLoc(const Mat& src, double* minVal, double* maxVal, Point* minLoc, Point* maxLoc) {
    if (src.type() != CV_32FC1 || src.empty()) throw std::invalid_argument("cvshim::minMaxLoc: CV_32FC1 only");
    float mn = src.at<float>(0, 0), mx = mn;
    Point pmn(0, 0), pmx(0, 0);
    for (int r = 0; r < src.rows; ++r)
        for (int c = 0; c < src.cols; ++c) {
            const float v = src.at<float>(r, c);
            if (v < mn) { mn = v; pmn = Point(c, r); }
            if (v > mx) { mx = v; pmx = Point(c, r); }
        }
    if (minVal) *minVal = mn;
    if (maxVal) *maxVal = mx;
    if (minLoc) *minLoc = pmn;
    if (maxLoc) *maxLoc = pmx;
}

} // namespace cv
