// Reference-GPU-semantics mode: the function the reference's own GTX-1080 kernels compute,
// which is NOT what its CPU code computes (the parity target of everything else in this library):
//   * window = (2R+1) rows x 2R columns: columns x-R .. x+R-1            (lib/DisparitySSD.cu:84,129; DisparityNCorr.cu:99,157)
//   * clamp-to-edge addressing instead of replicate padding + clamped candidate ranges; every d in [dmin, dmax] is tried
//   * float32 column sums, built for the first row of every 40-row block and then ROLLED (subtract the leaving row, add
//     the entering one) down the block - with the multiply-adds contracted to FMAs as nvcc does by default (:71-72,105-110)
//   * horizontal sum of the 2R column sums left to right in float32 (:85-87)
//   * SSD: a candidate wins if its cost is < the best so far, which starts at 5e6 (:16,88-91); NCC: score =
//     sum / sqrtf(a_t * a_i), wins if > the best so far, which starts at 0 (DisparityNCorr.cu:16,104-111); NaN never wins
//   * pixels nothing won keep -1 (:177); disparities are stored as `char`
// Every float operation happens in the reference's order, so the results are the ones its kernels produce (bit for bit
// against the C restatement the tests hold; the reference's own outputs are Git-LFS stubs and its .cu files need
// texture references that CUDA 12 no longer has).  One CTA = 64 + 2R threads, one shared-memory column sum each; the first
// 64 also own an output column.  No textures, no per-disparity read-modify-write of global memory.
#pragma once
#include "common.cuh"

namespace sb {

constexpr int CG_TILE = 64;         // output columns per CTA            (TILE_SIZE_X, DisparitySSD.cu:152)
constexpr int CG_ROWS = 40;         // rows per CTA                      (ROWS_PER_THREAD, :17)

template <int COST>
__global__ void refgpu_kernel(const float* __restrict__ L, size_t l_step, const float* __restrict__ Rt, size_t r_step, int rows,
                              int cols, int R, int dmin, int dmax, int8_t* __restrict__ disp_out, size_t disp_step,
                              float* __restrict__ best_out, size_t best_step) {
    extern __shared__ float cs[];                 // [3][CG_TILE + 2R]: products (or squared differences), template and image energies
    const int ncol = CG_TILE + 2 * R;
    float* ct = cs + ncol;
    float* ci = cs + 2 * ncol;
    const int t = threadIdx.x;                    // shared column t <-> image column x0 + t - R
    const int x0 = blockIdx.x * CG_TILE, gy = blockIdx.y * CG_ROWS;
    const int cx = x0 + t - R;
    const int xo = x0 + t;                        // output column of threads t < CG_TILE
    const bool owns = t < CG_TILE && xo < cols;
    auto at = [&](const float* img, size_t step, int x, int y) -> float {     // clamp-to-edge point sampling (tex2D, :19-20)
        const int xx = min(max(x, 0), cols - 1), yy = min(max(y, 0), rows - 1);
        return reinterpret_cast<const float*>(reinterpret_cast<const char*>(img) + size_t(yy) * step)[xx];
    };
    constexpr bool SSD = COST == STEREO_COST_SSD;
    const int nrow = min(CG_ROWS, rows - gy);     // output rows of this CTA
    // best-so-far lives in registers: one entry per output row of the block
    float best[CG_ROWS];
    int bestd[CG_ROWS];
#pragma unroll
    for (int r = 0; r < CG_ROWS; ++r) { best[r] = SSD ? 5000000.f : 0.f; bestd[r] = -1; }
    for (int d = dmin; d <= dmax; ++d) {
        float s = 0.f, st = 0.f, si = 0.f;
        // the first 2R+1 rows of the block, top to bottom
        for (int i = 0; i <= 2 * R; ++i) {
            const float a = at(L, l_step, cx, gy - R + i), b = at(Rt, r_step, cx + d, gy - R + i);
            if (SSD) { const float df = __fsub_rn(a, b); s = __fmaf_rn(df, df, s); }
            else { s = __fmaf_rn(a, b, s); st = __fmaf_rn(a, a, st); si = __fmaf_rn(b, b, si); }
        }
#pragma unroll 1
        for (int row = 0; row < CG_ROWS; ++row) {
            if (row > 0) {
                // the reference keeps rolling while row + gy < rows + R (:101); rows it never outputs are not needed here
                if (row >= nrow) break;
                const int yo = gy - R + row - 1, yn = yo + 2 * R + 1;
                const float a0 = at(L, l_step, cx, yo), b0 = at(Rt, r_step, cx + d, yo);
                const float a1 = at(L, l_step, cx, yn), b1 = at(Rt, r_step, cx + d, yn);
                if (SSD) {
                    const float d0 = __fsub_rn(a0, b0), d1 = __fsub_rn(a1, b1);
                    s = __fmaf_rn(-d0, d0, s);
                    s = __fmaf_rn(d1, d1, s);
                } else {
                    s = __fmaf_rn(-a0, b0, s); st = __fmaf_rn(-a0, a0, st); si = __fmaf_rn(-b0, b0, si);
                    s = __fmaf_rn(a1, b1, s); st = __fmaf_rn(a1, a1, st); si = __fmaf_rn(b1, b1, si);
                }
            }
            cs[t] = s;
            if (!SSD) { ct[t] = st; ci[t] = si; }
            __syncthreads();
            if (owns && row < nrow) {
                float acc = 0.f, acct = 0.f, acci = 0.f;
                for (int i = 0; i < 2 * R; ++i) {
                    acc = __fadd_rn(acc, cs[t + i]);
                    if (!SSD) { acct = __fadd_rn(acct, ct[t + i]); acci = __fadd_rn(acci, ci[t + i]); }
                }
                if (SSD) {
                    if (acc < best[row]) { best[row] = acc; bestd[row] = d; }
                } else {
                    const float sc = __fdiv_rn(acc, __fsqrt_rn(__fmul_rn(acct, acci)));
                    if (sc > best[row]) { best[row] = sc; bestd[row] = d; }
                }
            }
            __syncthreads();
        }
    }
    if (owns) {
#pragma unroll 1
        for (int row = 0; row < nrow; ++row) {
            reinterpret_cast<int8_t*>(reinterpret_cast<char*>(disp_out) + size_t(gy + row) * disp_step)[xo] = int8_t(bestd[row]);
            if (best_out) reinterpret_cast<float*>(reinterpret_cast<char*>(best_out) + size_t(gy + row) * best_step)[xo] = best[row];
        }
    }
}

} // namespace sb
