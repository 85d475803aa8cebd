// Device-side preprocessing of the ps2 executable: what main.cpp does to an image between imread and the
// disparity call, as kernels on images that stay on the device across the two directions of a pair and across problems.
//   convertTo(CV_32FC1) without scaling                      main.cpp:87-88
//   cvtColor(COLOR_RGB2GRAY) applied to BGR data, then convertTo   main.cpp:114-117  (OpenCV's fixed-point coefficients:
//       gray = (c0*R2Y + c1*G2Y + c2*B2Y + half) >> shift with the channels in memory order, shift 14 (3.4.1) or 15)
//   first + noise / left * 1.1f in float32                   main.cpp:146-152, 191-193
// The Gaussian noise itself stays on the host: cv::randn draws from ONE sequential generator (multiply-with-carry state,
// ziggurat with data-dependent rejections), and the reference's results depend on that exact stream.
#pragma once
#include "common.cuh"

namespace sb {

__global__ void gray_f32_kernel(const uint8_t* __restrict__ img, size_t step, int rows, int cols, int channels, int shift,
                                float* __restrict__ out, size_t out_step) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= cols || y >= rows) return;
    const uint8_t* p = img + size_t(y) * step + size_t(x) * channels;
    float v;
    if (channels == 1) v = float(p[0]);
    else {
        const int cr = shift == 15 ? 9798 : 4899, cg = shift == 15 ? 19235 : 9617, cb = shift == 15 ? 3735 : 1868;
        v = float((p[0] * cr + p[1] * cg + p[2] * cb + (1 << (shift - 1))) >> shift);
    }
    reinterpret_cast<float*>(reinterpret_cast<char*>(out) + size_t(y) * out_step)[x] = v;
}

// out = a * scale (+ add): separately rounded float32 operations, like the cv::Mat expressions they replace
__global__ void scale_add_f32_kernel(const float* __restrict__ a, size_t a_step, const float* __restrict__ add, size_t add_step,
                                     float scale, int rows, int cols, float* __restrict__ out, size_t out_step) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= cols || y >= rows) return;
    float v = reinterpret_cast<const float*>(reinterpret_cast<const char*>(a) + size_t(y) * a_step)[x];
    if (scale != 1.f) v = __fmul_rn(v, scale);
    if (add) v = __fadd_rn(v, reinterpret_cast<const float*>(reinterpret_cast<const char*>(add) + size_t(y) * add_step)[x]);
    reinterpret_cast<float*>(reinterpret_cast<char*>(out) + size_t(y) * out_step)[x] = v;
}

} // namespace sb
