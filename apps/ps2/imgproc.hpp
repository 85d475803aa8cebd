// The image operations the ps2 driver performs around the matcher (ProblemSets/ps2_cpp/src/main.cpp):
// grayscale conversion, convertTo(CV_32FC1), Gaussian noise, contrast gain, min-max normalisation to
// 8 bits and the "inverted" copy.  Each function restates the arithmetic of the OpenCV call the
// reference makes (OpenCV 3.4.1 is an un-vendored dependency, README.md:36); tests/test_ps2_app.py pins
// them against executed OpenCV (python cv2) in this image.
#pragma once
#include "image_io.hpp"
#include "../../include/stereo_b200.hpp"

#include <cfloat>
#include <cmath>
#include <cstdint>

namespace ps2 {

// cv::cvtColor(img, gray, cv::COLOR_RGB2GRAY, 1) applied to what imread returned, i.e. to B,G,R data
// (main.cpp:114-117): channel 0 gets the R weight.  8-bit path of OpenCV's RGB2Gray, fixed point, round to
// nearest.  `shift` selects the coefficient set: 14 = OpenCV <= 3.4.1, the version the reference pins
// (README.md:36): R2Y 4899, G2Y 9617, B2Y 1868; 15 = OpenCV >= 3.4.2 / 4.x: 9798, 19235, 3735 (the two
// differ by one grey level on ~0.2 % of pixels).
inline sb::Mat rgb2gray_on_bgr_as_float(const Image8& img, int shift = 14) {
    const int cr = shift == 15 ? 9798 : 4899, cg = shift == 15 ? 19235 : 9617, cb = shift == 15 ? 3735 : 1868;
    sb::Mat out(img.rows, img.cols, sb::F32C1);
    for (int r = 0; r < img.rows; ++r) {
        const uint8_t* s = img.row(r);
        float* d = out.ptr<float>(r);
        if (img.channels == 1) { for (int c = 0; c < img.cols; ++c) d[c] = float(s[c]); continue; }
        for (int c = 0; c < img.cols; ++c) {
            const uint8_t* p = s + size_t(c) * img.channels;
            d[c] = float((p[0] * cr + p[1] * cg + p[2] * cb + (1 << (shift - 1))) >> shift);   // then convertTo(CV_32FC1), unscaled
        }
    }
    return out;
}

// Mat::convertTo(dst, CV_32FC1) of a single-channel 8-bit image (main.cpp:87-88): no scaling.
inline sb::Mat to_float(const Image8& img) {
    if (img.channels != 1) throw std::runtime_error("expected a single-channel image (the reference asserts CV_32FC1, main.cpp:27)");
    sb::Mat out(img.rows, img.cols, sb::F32C1);
    for (int r = 0; r < img.rows; ++r) { const uint8_t* s = img.row(r); float* d = out.ptr<float>(r); for (int c = 0; c < img.cols; ++c) d[c] = float(s[c]); }
    return out;
}

// cv::RNG (multiply-with-carry) + cv::randn's float path (modules/core/src/rand.cpp: Ziggurat normal
// generator, then x*sigma + mean in float).  The reference never seeds it, so the state starts at
// OpenCV's default 0xffffffff and runs on across problems 3 and 4 (main.cpp:146-152).
class CvRng {
public:
    explicit CvRng(uint64_t state = 0xffffffffu) : state_(state ? state : 0xffffffffu) { init_tables(); }
    void fill_normal(sb::Mat& m, float mean, float sigma) {
        for (int r = 0; r < m.rows; ++r) { float* d = m.ptr<float>(r); for (int c = 0; c < m.cols; ++c) d[c] = next_normal() * sigma + mean; }
    }

private:
    uint64_t state_;
    unsigned kn_[128]; float wn_[128], fn_[128];
    uint32_t next() { const uint32_t lo = uint32_t(state_); state_ = uint64_t(lo) * 4164903690u + (state_ >> 32); return lo; }
    // NB: OpenCV reads the current state first and advances afterwards
    void init_tables() {
        const double m1 = 2147483648.0;
        double dn = 3.442619855899, tn = dn, vn = 9.91256303526217e-3;
        const double q = vn / std::exp(-.5 * dn * dn);
        kn_[0] = unsigned((dn / q) * m1); kn_[1] = 0;
        wn_[0] = float(q / m1); wn_[127] = float(dn / m1);
        fn_[0] = 1.f; fn_[127] = float(std::exp(-.5 * dn * dn));
        for (int i = 126; i >= 1; --i) {
            dn = std::sqrt(-2. * std::log(vn / dn + std::exp(-.5 * dn * dn)));
            kn_[i + 1] = unsigned((dn / tn) * m1);
            tn = dn;
            fn_[i] = float(std::exp(-.5 * dn * dn));
            wn_[i] = float(dn / m1);
        }
    }
    float next_normal() {
        const float r = 3.442620f, rng_flt = 2.3283064365386962890625e-10f;
        for (;;) {
            const int hz = int(next());
            const int iz = hz & 127;
            float x = hz * wn_[iz];
            if (unsigned(std::abs(hz)) < kn_[iz]) return x;
            if (iz == 0) {                                   // base strip: sample the tail
                float y;
                do {
                    x = next() * rng_flt;
                    y = next() * rng_flt;
                    x = float(-std::log(x + FLT_MIN) * 0.2904764);
                    y = float(-std::log(y + FLT_MIN));
                } while (y + y < x * x);
                return hz > 0 ? r + x : -r - x;
            }
            const float y = next() * rng_flt;               // wedges of the other strips
            if (fn_[iz] + y * (fn_[iz - 1] - fn_[iz]) < std::exp(-.5 * x * x)) return x;
        }
    }
};

// addNoise (main.cpp:140-153): first + N(mean, sigma), then second + N(mean, sigma), unclipped floats.
inline void add_noise(CvRng& rng, const sb::Mat& first, const sb::Mat& second, float mean, float sigma, sb::Mat& first_noisy, sb::Mat& second_noisy) {
    auto one = [&](const sb::Mat& src, sb::Mat& dst) {
        sb::Mat noise(src.rows, src.cols, sb::F32C1);
        rng.fill_normal(noise, mean, sigma);
        dst.create(src.rows, src.cols, sb::F32C1);
        for (int r = 0; r < src.rows; ++r) { const float *a = src.ptr<float>(r), *n = noise.ptr<float>(r); float* d = dst.ptr<float>(r); for (int c = 0; c < src.cols; ++c) d[c] = a[c] + n[c]; }
    };
    one(first, first_noisy);
    one(second, second_noisy);
}

// `left * contrastFactor` (main.cpp:191-193): float multiply.
inline sb::Mat scaled(const sb::Mat& src, float gain) {
    sb::Mat out(src.rows, src.cols, sb::F32C1);
    for (int r = 0; r < src.rows; ++r) { const float* a = src.ptr<float>(r); float* d = out.ptr<float>(r); for (int c = 0; c < src.cols; ++c) d[c] = a[c] * gain; }
    return out;
}

// cv::normalize(disp, disp, 0, 255, cv::NORM_MINMAX, CV_8UC1) on a CV_8SC1 map (main.cpp:94):
// scale = 255 / (max - min) (0 if the map is constant), shift = -min*scale, both double; the
// conversion is saturate_cast<uchar>(cvRound(src*float(scale) + float(shift))).
inline sb::Mat normalize_minmax_u8(const sb::Mat& disp) {
    int lo = 127, hi = -128;
    for (int r = 0; r < disp.rows; ++r) { const int8_t* s = disp.ptr<int8_t>(r); for (int c = 0; c < disp.cols; ++c) { lo = std::min<int>(lo, s[c]); hi = std::max<int>(hi, s[c]); } }
    const double scale = 255.0 * ((hi - lo) > DBL_EPSILON ? 1.0 / (hi - lo) : 0.0), shift = 0.0 - lo * scale;
    const float a = float(scale), b = float(shift);
    sb::Mat out(disp.rows, disp.cols, sb::U8C1);
    for (int r = 0; r < disp.rows; ++r) {
        const int8_t* s = disp.ptr<int8_t>(r);
        uint8_t* d = out.ptr<uint8_t>(r);
        for (int c = 0; c < disp.cols; ++c) {
            const long v = std::lrintf(float(s[c]) * a + b);          // cvRound: nearest, ties to even
            d[c] = uint8_t(v < 0 ? 0 : (v > 255 ? 255 : v));
        }
    }
    return out;
}

// `ones * 255 - img` (main.cpp:126-128)
inline sb::Mat inverted(const sb::Mat& img) {
    sb::Mat out(img.rows, img.cols, sb::U8C1);
    for (int r = 0; r < img.rows; ++r) { const uint8_t* s = img.ptr<uint8_t>(r); uint8_t* d = out.ptr<uint8_t>(r); for (int c = 0; c < img.cols; ++c) d[c] = uint8_t(255 - s[c]); }
    return out;
}

} // namespace ps2
