// Packed u8 path: the hot kernels of the library (sm_100a).
//
// For images that are exactly 8-bit the window cost is an integer and
//     SSD(x,d) = EL(x) + ER(x+d) - 2*C(x,d),   C = sum over the window of l*r,
//     NCC(x,d) = C(x,d) / sqrt(EL(x) * ER(x+d)),
// so the per-disparity work is ONE box-filtered cross term C; the energies are box sums computed
// once per image, not per disparity (north-star item (b)).  The kernel computes C with separable
// running sums held entirely in registers:
//
//   * vertical:   col[m][c] += l_new*r_new - l_old*r_old      ONE IDP.2A (dp2a, s16 x u8) per unit
//                 (the new and the leaving row are packed into one operand pair by the prep kernels)
//   * horizontal: s[m] += col[m][c+2R+1] - col[m][c]          ONE IADD3 per unit
//   * WTA:        key = E2[pos] + 256*s (cost*128 + position) ONE IMAD/LEA, then VIMNMX3 in-thread
//                 and REDUX.MIN across the 32 lanes of the warp (lanes = disparities)
//
// The cost volume never exists in memory: a thread owns K=24 pixels x M=4 disparities, a warp
// 24 pixels x 128 disparities, and only the winning key per pixel and 128-disparity group leaves
// the SM.  Operand rows are staged in shared memory with TMA bulk copies (cp.async.bulk, SASS
// UBLKCP) through a 4-stage mbarrier ring; all warps of a CTA share the staged rows.
//
// Reference semantics reproduced: replicate padding, the clamped candidate
// range in padded coordinates, the flat-index row wrap of the SSD target reads (realised by building
// the target operand rows from an "extended" image whose out-of-row columns come from the
// neighbouring padded row / zero guard), first-minimum (SSD) and first-maximum (NCC) tie-breaks.
#pragma once
#include "common.cuh"
#include "exact.cuh"
#include <climits>
#include <cstdlib>

namespace sb {

constexpr int FK_DEFAULT = 24;  // pixels per thread (strip width); template parameter K of the kernels
constexpr int FM = 4;           // disparities per thread
constexpr int FGROUP = 32 * FM; // disparities per warp ("group")
constexpr int FKEY_BITS = 7;    // log2(FGROUP): low bits of a key order candidates inside a group
constexpr int FRPS = 8;         // operand rows per pipeline stage
constexpr int FNST = 8;         // pipeline stages
constexpr int FWARPS_MAX = 12;   // warps per CTA: 8 (K=24, <=255 regs) or 12 (K=16, <=168 regs)
constexpr int FMAXR = 7;        // largest window radius with 32-bit keys: 128*(2R+1)^2*255^2 < 2^31
constexpr uint32_t KEY_INVALID = 0xFFFFFFFFu;
constexpr int FFREE_MASK_R = 5;   // largest radius for which invalid candidates lose through the key alone

// Keys are unsigned:  key = BIAS + 128*(ER - 2C) + position,  BIAS = 128*Emax, Emax = (2R+1)^2*255^2,
// so valid keys lie in [0, 256*Emax + position].  A candidate whose centre is not a legal search
// position carries E2 = KEY_INVALID; its key KEY_INVALID - 256*C stays above every valid key as long
// as 512*Emax < 2^32 (R <= 5), i.e. border masking costs no instruction there.
__host__ __device__ static inline uint32_t key_emax(int R) { return uint32_t((2 * R + 1) * (2 * R + 1)) * 65025u; }
__host__ __device__ static inline uint32_t key_bias(int R) { return key_emax(R) << FKEY_BITS; }
__host__ __device__ static inline uint32_t key_invalid_threshold(int R) {
    return R <= FFREE_MASK_R ? KEY_INVALID - (key_emax(R) << (FKEY_BITS + 1)) : KEY_INVALID;
}

// ---------------------------------------------------------------------------------------------------
// Geometry shared by host and device
// ---------------------------------------------------------------------------------------------------
struct FastGeom {
    // problem
    int rows, cols, R, dmin, dmax, cost;
    int rb, re;            // output band
    int ar0, ar1;          // image rows present in the caller's buffers (full image: 0, rows); reads clamp into it
    int K;                 // pixels per thread (strip width): 24 or 16
    int nw;                // warps per CTA: 8 (K=24) or 12 (K=16)
    // derived
    int G;                 // number of 128-disparity groups
    int gc;                // groups per CTA (1 or 2)
    int spc;               // strips per CTA = FWARPS / gc
    int nstrips;           // ceil(cols / FK)
    int tilesX;            // ceil(nstrips / spc)
    int gblocks;           // ceil(G / gc)
    int base_y;            // step row of operand row 0 (multiple of FRPS, <= rb - (2R+1))
    int J;                 // operand rows (multiple of FRPS)
    int qoff, eoff;        // column offsets of the RQ / E2 arrays
    int lp_pitch, rq_pitch, e2_pitch;   // words
    int lpw, rqw, e2w;     // tile widths in words (multiples of 4)
    int wpart;             // partial-key map width (= tilesX*spc*FK)
    int nrows;             // re - rb
    int ctas;              // grid size
    long long total;       // tile-rows
    long long L;           // tile-rows per CTA
    // valid centre columns (unpadded coordinates)
    int cmin, cmax;
};

static inline int round_up(int v, int m) { return (v + m - 1) / m * m; }
static inline int floor_div(int a, int b) { int q = a / b; if ((a % b != 0) && ((a < 0) != (b < 0))) --q; return q; }

struct FastArrays {
    int32_t* LP;       // [J][lp_pitch]   s16x2: (-l(y+R), +l(y-R-1))
    uint32_t* RQ;      // [J/2][rq_pitch] u8x4 : (r(ye+R), r(ye-R-1), r(ye+1+R), r(ye-R))
    int32_t* E2;       // [J][e2_pitch]   BIAS + 128*ER + position, or KEY_INVALID
    int32_t* PART;     // [G][nrows][wpart] winning keys
    int32_t* V;        // [nrows][vpitch] vertical (2R+1)-sums of squares of the extended target image
    float* RS;         // NCC: 1/sqrt(ER) per position   [J][e2_pitch]
    float* SC;         // NCC: [nstrips][nrows] magic = 2^ceil(log2 sqrt(max EL of the strip row))
};

// ---------------------------------------------------------------------------------------------------
// Device helpers
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ int dp2a_lo(int a, unsigned b, int c) {
    int d; asm("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d;
}
__device__ __forceinline__ int dp2a_hi(int a, unsigned b, int c) {
    int d; asm("dp2a.hi.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d;
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
// TMA bulk copy global -> shared, completion on an mbarrier (SASS: UBLKCP).
__device__ __forceinline__ void tma_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

__device__ __forceinline__ int4 lds128(const int* p) { return *reinterpret_cast<const int4*>(p); }

// The target image the reference actually reads: padded row i+R of the
// replicate-padded image, extended by R columns either side that alias the neighbouring padded row
// through the flat index of the unchecked cv::Mat::at (DisparitySSD.cpp:50).  `i` is the unpadded row
// index (may lie outside [0,rows)), `e` = unpadded column + 2R.  Reads outside the padded buffer
// (row -1 / row Hp) return 0 (the oracle's zero guard).
__device__ __forceinline__ int bext(const uint8_t* __restrict__ B, size_t step, int rows, int cols, int R, int i, int e,
                                    int ar0, int ar1) {
    const int Wp = cols + 2 * R;
    const int c = e - R;                 // padded column, may be < 0 or >= Wp
    int src_row = i, src_col;
    if (c < 0) {                         // previous padded row, right padding
        if (i + R - 1 < 0) return 0;
        src_row = i - 1; src_col = cols - 1;
    } else if (c >= Wp) {                // next padded row, left padding
        if (i + R + 1 > rows + 2 * R - 1) return 0;
        src_row = i + 1; src_col = 0;
    } else {
        src_col = clampi(c - R, 0, cols - 1);
    }
    src_row = clampi(clampi(src_row, 0, rows - 1), ar0, ar1 - 1);
    return B[size_t(src_row) * step + src_col];
}

// ---------------------------------------------------------------------------------------------------
// Prep kernels: operand rows in the layout the hot loop consumes
// ---------------------------------------------------------------------------------------------------
// LP[j][p]: 4 columns per thread, 16-byte stores.
__global__ void __launch_bounds__(256) prep_lp_kernel(const uint8_t* __restrict__ A, size_t step, FastGeom g, int32_t* __restrict__ LP) {
    const int p4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const int j = blockIdx.y;
    if (p4 >= g.lp_pitch) return;
    const int y = g.base_y + j;
    const uint8_t* rnew = A + size_t(clampi(clampi(y + g.R, 0, g.rows - 1), g.ar0, g.ar1 - 1)) * step;
    const uint8_t* rold = A + size_t(clampi(clampi(y - g.R - 1, 0, g.rows - 1), g.ar0, g.ar1 - 1)) * step;
    int v[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        const int col = clampi(p4 + t - g.R, 0, g.cols - 1);
        const int lnew = rnew[col], lold = rold[col];
        // SSD: (-l_new, +l_old) so the running sums hold -C; NCC: (+l_new, -l_old), sums hold +C
        const int a_new = g.cost == STEREO_COST_SSD ? -lnew : lnew, a_old = g.cost == STEREO_COST_SSD ? lold : -lold;
        v[t] = int(uint32_t(uint16_t(int16_t(a_new))) | (uint32_t(uint16_t(int16_t(a_old))) << 16));
    }
    *reinterpret_cast<int4*>(LP + size_t(j) * g.lp_pitch + p4) = make_int4(v[0], v[1], v[2], v[3]);
}

// RQ[jp][q]: 4 positions per thread, 16-byte stores.
__global__ void __launch_bounds__(256) prep_rq_kernel(const uint8_t* __restrict__ B, size_t step, FastGeom g, uint32_t* __restrict__ RQ) {
    const int q4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const int jp = blockIdx.y;
    if (q4 >= g.rq_pitch) return;
    const int ye = g.base_y + 2 * jp;
    uint32_t v[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        const int e = q4 + t - g.qoff;
        const uint32_t b0 = bext(B, step, g.rows, g.cols, g.R, ye + g.R, e, g.ar0, g.ar1);
        const uint32_t b1 = bext(B, step, g.rows, g.cols, g.R, ye - g.R - 1, e, g.ar0, g.ar1);
        const uint32_t b2 = bext(B, step, g.rows, g.cols, g.R, ye + 1 + g.R, e, g.ar0, g.ar1);
        const uint32_t b3 = bext(B, step, g.rows, g.cols, g.R, ye - g.R, e, g.ar0, g.ar1);
        v[t] = b0 | (b1 << 8) | (b2 << 16) | (b3 << 24);
    }
    *reinterpret_cast<uint4*>(RQ + size_t(jp) * g.rq_pitch + q4) = make_uint4(v[0], v[1], v[2], v[3]);
}

// Window energies of the extended target image, separably:
//   pass 1 (prep_v_kernel): V[yy][x] = sum_{j=-R..R} bext(y+j, e)^2, a running sum down the rows;
//           one thread per column and PV_ROWS-row chunk, no synchronisation.  Column x <-> e = x + vbase.
//   pass 2 (prep_e2_kernel): ER = sum_{t=-R..R} V[yy][centre + t] from a shared-memory tile, then
//           E2[j][q2] = BIAS + 128*ER + q2 for legal centres (KEY_INVALID otherwise); RS (NCC) = 1/sqrt(ER).
constexpr int PV_ROWS = 32;
__global__ void __launch_bounds__(128) prep_v_kernel(const uint8_t* __restrict__ B, size_t step, FastGeom g,
                                                    int32_t* __restrict__ V, int vpitch, int vbase) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= vpitch) return;
    const int e = x + vbase;
    const int yy0 = blockIdx.y * PV_ROWS, yy1 = min(g.nrows, yy0 + PV_ROWS);
    const int R = g.R;
    int v = 0;
    for (int i = g.rb + yy0 - R; i < g.rb + yy0 + R; ++i) { const int b = bext(B, step, g.rows, g.cols, R, i, e, g.ar0, g.ar1); v += b * b; }
    for (int yy = yy0; yy < yy1; ++yy) {
        const int y = g.rb + yy;
        const int bn = bext(B, step, g.rows, g.cols, R, y + R, e, g.ar0, g.ar1);
        v += bn * bn;
        V[size_t(yy) * vpitch + x] = v;
        const int bo = bext(B, step, g.rows, g.cols, R, y - R, e, g.ar0, g.ar1);
        v -= bo * bo;
    }
}

constexpr int PE_COLS = 256;
constexpr int PE_ROWS = 8;
__global__ void __launch_bounds__(PE_COLS) prep_e2_kernel(const int32_t* __restrict__ V, int vpitch, FastGeom g,
                                                         int32_t* __restrict__ E2, float* __restrict__ RS) {
    extern __shared__ int pe_smem[];                       // [PE_ROWS][PE_COLS + 2R]
    const int R = g.R, tw = PE_COLS + 2 * R;
    const int tid = threadIdx.x;
    const int q20 = blockIdx.x * PE_COLS;
    const int yy0 = blockIdx.y * PE_ROWS;
    const int nr = min(PE_ROWS, g.nrows - yy0);
    // V column x <-> ext column e = x + vbase with vbase = -eoff + R, so centre q2 (e_c = q2 - eoff + 2R) is
    // V column q2 + R and its window is V columns q2 .. q2 + 2R.
    for (int idx = tid; idx < nr * tw; idx += PE_COLS) {
        const int r = idx / tw, c = idx - r * tw;
        pe_smem[idx] = V[size_t(yy0 + r) * vpitch + q20 + c];
    }
    __syncthreads();
    const int q2 = q20 + tid;
    if (q2 >= g.e2_pitch) return;
    const int uc = q2 - g.eoff;
    const bool valid = uc >= g.cmin && uc <= g.cmax;
    for (int r = 0; r < nr; ++r) {
        int er = 0;
        for (int t = 0; t <= 2 * R; ++t) er += pe_smem[r * tw + tid + t];
        const size_t o = size_t(g.rb + yy0 + r - g.base_y) * g.e2_pitch + q2;
        if (g.cost == STEREO_COST_SSD) {
            E2[o] = int(valid ? key_bias(R) + (uint32_t(er) << FKEY_BITS) + uint32_t(q2) : KEY_INVALID);
        } else {
            E2[o] = valid ? er : -1;
            RS[o] = (valid && er > 0) ? float(1.0 / sqrt(double(er))) : 0.f;
        }
    }
}

// NCC: per strip (K pixels) and output row, the power of two just above sqrt(max EL) — the binade the
// fixed-point keys of that strip row live in.  V holds the vertical (2R+1)-sums of squares of the
// replicate-padded REFERENCE image: V column c = padded column c, so EL(x) = sum V[yy][x .. x+2R].
__global__ void __launch_bounds__(128) prep_scale_kernel(const int32_t* __restrict__ V, int vpitch, FastGeom g,
                                                        float* __restrict__ SC) {
    const int strip = blockIdx.x * blockDim.x + threadIdx.x;
    const int yy = blockIdx.y;
    if (strip >= g.tilesX * g.spc) return;
    const int x0 = strip * g.K;
    float magic = 1.f;
    if (x0 < g.cols) {
        const int32_t* v = V + size_t(yy) * vpitch + x0;
        const int x1 = min(g.K, g.cols - x0);
        int el = 0, elmax = 0;
        for (int t = 0; t < 2 * g.R; ++t) el += v[t];
        for (int x = 0; x < x1; ++x) { el += v[x + 2 * g.R]; elmax = max(elmax, el); el -= v[x]; }
        // smallest power of two strictly above sqrt(elmax) * (1 + 2^-20)  (C*rs <= sqrt(EL), rounding slack)
        const float bound = float(sqrt(double(elmax)) * (1.0 + 1.0 / 1048576.0));
        int e; frexpf(bound, &e);                                       // bound = f * 2^e, f in [0.5, 1)
        magic = elmax > 0 ? ldexpf(1.f, e) : 1.f;                       // 2^e > bound
    }
    SC[size_t(strip) * g.nrows + yy] = magic;
}

// ---------------------------------------------------------------------------------------------------
// The hot kernel
// ---------------------------------------------------------------------------------------------------
struct FastKernelParams {
    FastGeom g;
    const int32_t* LP;
    const uint32_t* RQ;
    const int32_t* E2;   // SSD: key offsets; NCC: the RS array (f32 bit patterns)
    int32_t* PART;
    const float* SC;     // NCC: [strip][output row] power-of-two magic (see fast_row)
};

template <int R, int K>
struct RowShape {
    static constexpr int NC = K + 2 * R;              // columns whose sums a thread keeps
    static constexpr int NC4 = (NC + 3) / 4 * 4;
    static constexpr int NQ = NC + FM - 1;             // target positions a thread touches
    static constexpr int NQ4 = (NQ + 3) / 4 * 4;
    static constexpr int NE = K + FM - 1;              // centre positions
    static constexpr int NE4 = (NE + 3) / 4 * 4;
};

// One operand row for one warp.
//   MODE 0: warm-up (add the entering row only, no output)
//   MODE 1: regular row; SSD: invalid search positions lose through their E2 entry alone (R <= 5);
//           NCC: every candidate of the block is legal
//   MODE 2: MODE 1 + whole lanes beyond max_disp are excluded (one LOP3 per pixel)
//   MODE 3: explicit per-candidate selects (partially valid lanes, border positions)
// PAR selects the byte pair of the RQ words (even/odd step row).
// The column updates (IDP.2A, FMA-heavy pipe) are interleaved with the horizontal slide / WTA of the
// same row (IADD3, VIMNMX on the ALU pipe) so that a single warp keeps both half-rate pipes busy.
//
// NCC keys.  v = C * RS[pos] (f32, C an exact integer) orders the candidates of one pixel; the oracle's
// first-maximum rule (cv::minMaxLoc, DisparityNCorr.cpp:62-64) needs the full f32 precision of v AND the
// position in one 32-bit key, which an f32 bit pattern cannot hold.  So v is turned into a 23-bit
// fixed-point number first: r = v + magic with magic = the power of two just above sqrt(max EL) of the
// strip row (C*RS <= sqrt(EL) by Cauchy-Schwarz), i.e. r lies in the binade [magic, 2*magic) and its
// mantissa IS round(v * 2^23 / magic).  key = (bits(r) << 9) + (reversed position << 2 | 3): 23 value
// bits, 7 position bits, low bits 11 so that 0 can mean "no legal candidate".  Unsigned max.
// For (2R+1)^2*255^2 < 2^23 (R <= 5) the running sums carry the float bias 0x4B000000, i.e. they ARE
// the float 2^23 + C, and v = fma(2^23 + C, rs, -2^23*rs) is the correctly rounded product with no
// conversion instruction; larger windows convert with I2F and fold the magic add into the FFMA.
#ifndef SB_KEY_LEA_MASK
#define SB_KEY_LEA_MASK 0
#endif
constexpr int NCC_BIAS_MAX_R = 5;
constexpr int NCC_FLOAT_BIAS = 0x4B000000;        // bit pattern of 8388608.0f
constexpr int NCC_KEY_SHIFT = 9;                  // mantissa -> bits 9..31
constexpr uint32_t NCC_KEY_NONE = 0u;             // "no legal candidate" (loses every unsigned max)

template <int R, int K, int PAR, int MODE, int COST>
__device__ __forceinline__ void fast_row(int (&col)[FM][RowShape<R, K>::NC], const int* __restrict__ lp_row,
                                         const int* __restrict__ rq_row, const int* __restrict__ e2_row,
                                         int32_t* __restrict__ out_row, int mmax, uint32_t lane_or, int lane, int cbase,
                                         int cols, float magic) {
    using S = RowShape<R, K>;
    constexpr bool NCC = (COST == STEREO_COST_NCORR);
    constexpr bool BIASED = NCC && (R <= NCC_BIAS_MAX_R);
    int lpv[S::NC4];
    int rqv[S::NQ4];
#pragma unroll
    for (int i = 0; i < S::NC4 / 4; ++i) {
        const int4 v = lds128(lp_row + 4 * i);
        lpv[4 * i] = v.x; lpv[4 * i + 1] = v.y; lpv[4 * i + 2] = v.z; lpv[4 * i + 3] = v.w;
    }
#pragma unroll
    for (int i = 0; i < S::NQ4 / 4; ++i) {
        const int4 v = lds128(rq_row + 4 * i);
        rqv[4 * i] = v.x; rqv[4 * i + 1] = v.y; rqv[4 * i + 2] = v.z; rqv[4 * i + 3] = v.w;
    }
    auto update = [&](int c) {
        const int a = (MODE == 0) ? (lpv[c] & 0xFFFF) : lpv[c];      // warm-up: entering row only
#pragma unroll
        for (int m = 0; m < FM; ++m)
            col[m][c] = PAR ? dp2a_hi(a, unsigned(rqv[c + m]), col[m][c]) : dp2a_lo(a, unsigned(rqv[c + m]), col[m][c]);
    };
    if (MODE == 0) {
#pragma unroll
        for (int c = 0; c < S::NC; ++c) update(c);
        return;
    }
    int e2v[S::NE4];
#pragma unroll
    for (int i = 0; i < S::NE4 / 4; ++i) {
        const int4 v = lds128(e2_row + 4 * i);
        e2v[4 * i] = v.x; e2v[4 * i + 1] = v.y; e2v[4 * i + 2] = v.z; e2v[4 * i + 3] = v.w;
    }
    int s[FM];
#pragma unroll
    for (int m = 0; m < FM; ++m) s[m] = BIASED ? NCC_FLOAT_BIAS : 0;
#pragma unroll
    for (int c = 0; c < 2 * R; ++c) {
        update(c);
#pragma unroll
        for (int m = 0; m < FM; ++m) s[m] += col[m][c];
    }
    uint32_t res[4];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        update(k + 2 * R);
        uint32_t key[FM];
#pragma unroll
        for (int m = 0; m < FM; ++m) {
            s[m] = s[m] + col[m][k + 2 * R] - (k > 0 ? col[m][k - 1] : 0);
            if (!NCC) {
                // s = -C (the packed operand carries -l): key = BIAS + 128*(ER - 2C) + position.
                // literal multiplier: ptxas emits the immediate-form IMAD / LEA (2 register reads; the
                // register file delivers ~2 operands per cycle per SMSP, tools/microbench/rf.cu)
                uint32_t kv;
                if (SB_KEY_LEA_MASK & (1 << m)) {      // shift-add on the ALU pipe
                    asm("{.reg .b32 t; shl.b32 t, %1, 8; add.s32 %0, t, %2;}" : "=r"(kv) : "r"(s[m]), "r"(e2v[k + m]));
                } else {                                // IMAD (immediate) on the FMA-heavy pipe
                    kv = uint32_t(e2v[k + m]) + uint32_t(s[m]) * uint32_t(2 << FKEY_BITS);
                }
                if (MODE == 3) kv = (uint32_t(e2v[k + m]) == KEY_INVALID || m > mmax) ? KEY_INVALID : kv;
                key[m] = kv;
            } else {
                const float rs = __int_as_float(e2v[k + m]);
                const float r = BIASED ? __fadd_rn(__fmaf_rn(__int_as_float(s[m]), rs, rs * -8388608.0f), magic)
                                       : __fmaf_rn(__int2float_rn(s[m]), rs, magic);
                // lane_or carries ((127 - 4*lane) << 2) | 3: reversed position of the lane's first candidate
                uint32_t kv = (uint32_t(__float_as_int(r)) << NCC_KEY_SHIFT) + (lane_or - 4u * m);
                if (MODE == 3) kv = (unsigned(cbase + k + m) >= unsigned(cols) || m > mmax) ? NCC_KEY_NONE : kv;
                key[m] = kv;
            }
        }
        uint32_t best;
        if (!NCC) {
            best = min(min(key[0], key[1]), min(key[2], key[3]));
            if (MODE == 2) best |= lane_or;
            res[k & 3] = __reduce_min_sync(0xffffffffu, best);
        } else {
            best = max(max(key[0], key[1]), max(key[2], key[3]));
            if (MODE == 2) best = mmax < 0 ? NCC_KEY_NONE : best;
            res[k & 3] = __reduce_max_sync(0xffffffffu, best);
        }
        if ((k & 3) == 3 && lane == 0)
            *reinterpret_cast<uint4*>(out_row + k - 3) = make_uint4(res[0], res[1], res[2], res[3]);
    }
}

template <int R, int K, int NW, int COST>
__global__ void __launch_bounds__(NW * 32, 1) fast_cost_kernel(const FastKernelParams P) {
    constexpr bool NCC = (COST == STEREO_COST_NCORR);
    using S = RowShape<R, K>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const FastGeom& g = P.g;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    const int lp_stage = FRPS * g.lpw, rq_stage = (FRPS / 2) * g.rqw, e2_stage = FRPS * g.e2w;   // words
    const int stage_words = lp_stage + rq_stage + e2_stage;
    int* smem = reinterpret_cast<int*>(smem_raw);
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem_raw + size_t(FNST) * stage_words * 4);
    const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + FNST);

    if (tid == 0) {
        for (int i = 0; i < FNST; ++i) { mbar_init(full0 + 8 * i, 1); mbar_init(empty0 + 8 * i, NW); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // ---- this CTA's share of the (tile, row) space -------------------------------------------------
    const long long lin_begin = (long long)blockIdx.x * g.L;
    long long lin_end = lin_begin + g.L;
    if (lin_end > g.total) lin_end = g.total;
    if (lin_begin >= lin_end) return;
    const int w = 2 * R + 1;

    // producer state (thread 0 only): iterates the same stage sequence, FNST-2 stages ahead
    long long p_lin = lin_begin;   // start of the producer's current segment
    int p_sj = 0, p_sj_end = -1;   // stage range of the producer's segment
    int p_tile = 0;
    int pn = 0;                    // loads issued
    auto producer_open_segment = [&]() {
        p_tile = int(p_lin / g.nrows);
        const int r0 = int(p_lin % g.nrows);
        long long rem = lin_end - p_lin;
        const int r1 = (rem > g.nrows - r0) ? g.nrows : r0 + int(rem);
        const int js = g.rb + r0 - w - g.base_y, je = g.rb + r1 - g.base_y;
        p_sj = js / FRPS; p_sj_end = (je - 1) / FRPS;
        p_lin += r1 - r0;
    };
    auto producer_issue = [&]() -> bool {     // returns false when nothing is left
        if (p_sj > p_sj_end) {
            if (p_lin >= lin_end) return false;
            producer_open_segment();
        }
        const int slot = pn % FNST;
        if (pn >= FNST) mbar_wait(empty0 + 8 * slot, ((pn / FNST) - 1) & 1);
        const int xt = p_tile % g.tilesX, gb = p_tile / g.tilesX;
        const int p0 = xt * g.spc * K;
        const int q0 = p0 + g.dmin + FGROUP * gb * g.gc + g.R + g.qoff;
        const int q20 = p0 + g.dmin + FGROUP * gb * g.gc + g.eoff;
        const uint32_t bar = full0 + 8 * slot;
        int* st = smem + size_t(slot) * stage_words;
        mbar_expect_tx(bar, uint32_t(stage_words) * 4u);
        const int j0 = p_sj * FRPS;
#pragma unroll 1
        for (int r = 0; r < FRPS; ++r)
            tma_load_1d(smem_u32(st + r * g.lpw), P.LP + size_t(j0 + r) * g.lp_pitch + p0, uint32_t(g.lpw) * 4u, bar);
#pragma unroll 1
        for (int r = 0; r < FRPS / 2; ++r)
            tma_load_1d(smem_u32(st + lp_stage + r * g.rqw), P.RQ + size_t(j0 / 2 + r) * g.rq_pitch + q0, uint32_t(g.rqw) * 4u, bar);
#pragma unroll 1
        for (int r = 0; r < FRPS; ++r)
            tma_load_1d(smem_u32(st + lp_stage + rq_stage + r * g.e2w), P.E2 + size_t(j0 + r) * g.e2_pitch + q20, uint32_t(g.e2w) * 4u, bar);
        ++pn; ++p_sj;
        return true;
    };
    // The producer is lane 0 of the LAST warp: the SMSP arbiter favours higher warp ids, so that warp
    // tends to run ahead and operand rows are requested as early as the ring allows.
    const bool is_producer = (tid == (NW - 1) * 32);
    if (is_producer) {
        for (int i = 0; i < FNST - 2; ++i) if (!producer_issue()) break;
    }

    // ---- consumers -----------------------------------------------------------------------------------
    int n = 0;                      // stages consumed
    long long lin = lin_begin;
    int col[FM][S::NC];
    while (lin < lin_end) {
        const int tile = int(lin / g.nrows);
        const int r0 = int(lin % g.nrows);
        const long long rem = lin_end - lin;
        const int r1 = (rem > g.nrows - r0) ? g.nrows : r0 + int(rem);
        lin += r1 - r0;
        const int xt = tile % g.tilesX, gb = tile / g.tilesX;
        const int strip = xt * g.spc + warp / g.gc;
        const int grp = gb * g.gc + warp % g.gc;
        const int x0 = strip * K;
        const bool active = (x0 < g.cols) && (grp < g.G);
        const int y0 = g.rb + r0, y1 = g.rb + r1;
        const int js = y0 - w - g.base_y, je = y1 - g.base_y, jreg = y0 - g.base_y;
        // which flavour of candidate masking this warp's (24 pixels x 128 disparities) block needs
        const int dlo = g.dmin + FGROUP * grp;                         // first disparity of the group
        const bool pos_invalid = (x0 + dlo < g.cmin) || (x0 + K - 1 + dlo + FGROUP - 1 > g.cmax);
        const bool lane_invalid = dlo + FGROUP - 1 > g.dmax;
        const bool partial_lane = lane_invalid && (((g.dmax - dlo + 1) % FM) != 0);
        const int mode = (partial_lane || (pos_invalid && (NCC || R > FFREE_MASK_R))) ? 3 : (lane_invalid ? 2 : 1);
        const int mmax = g.dmax - dlo - FM * lane;                      // m <= mmax are inside [dmin, dmax]
        // SSD: OR-mask that invalidates a whole lane; NCC: reversed position of the lane's first candidate
        const uint32_t lane_or = NCC ? (uint32_t(FGROUP - 1 - FM * lane) << 2 | 3u) : (mmax < 0 ? KEY_INVALID : 0u);
        const float* sc_row = NCC ? P.SC + size_t(strip) * g.nrows - (g.rb - g.base_y) : nullptr;   // indexed by operand row j
        const int cbase = x0 + dlo + FM * lane;                          // centre column of candidate (k=0, m=0)
        const int lp_off = (warp / g.gc) * K;
        const int rq_off = (warp / g.gc) * K + FGROUP * (warp % g.gc) + FM * lane;
        int32_t* part = P.PART + (size_t(grp) * g.nrows) * g.wpart + x0;
#pragma unroll
        for (int m = 0; m < FM; ++m)
#pragma unroll
            for (int c = 0; c < S::NC; ++c) col[m][c] = 0;

        for (int sj = js / FRPS; sj <= (je - 1) / FRPS; ++sj, ++n) {
            if (is_producer) producer_issue();
            const int slot = n % FNST;
            mbar_wait(full0 + 8 * slot, (n / FNST) & 1);
            if (active) {
                const int* st = smem + size_t(slot) * stage_words;
                const int jlo = max(js, sj * FRPS), jhi = min(je, sj * FRPS + FRPS);
                for (int j = jlo; j < jhi; ++j) {
                    const int r = j - sj * FRPS;
                    const int* lp_row = st + r * g.lpw + lp_off;
                    const int* rq_row = st + lp_stage + (r >> 1) * g.rqw + rq_off;
                    const int* e2_row = st + lp_stage + rq_stage + r * g.e2w + rq_off;
                    int32_t* out_row = part + size_t(j - (g.rb - g.base_y)) * g.wpart;
                    const int par = j & 1;
                    const float magic = (NCC && j >= jreg) ? __ldg(sc_row + j) : 0.f;
#define SB_ROW(P_, M_) fast_row<R, K, P_, M_, COST>(col, lp_row, rq_row, e2_row, out_row, mmax, lane_or, lane, cbase, g.cols, magic)
                    if (j < jreg)       { if (par) SB_ROW(1, 0); else SB_ROW(0, 0); }
                    else if (mode == 1) { if (par) SB_ROW(1, 1); else SB_ROW(0, 1); }
                    else if (mode == 2) { if (par) SB_ROW(1, 2); else SB_ROW(0, 2); }
                    else                { if (par) SB_ROW(1, 3); else SB_ROW(0, 3); }
#undef SB_ROW
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(empty0 + 8 * slot);
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// Merge: winning key per group -> disparity (+ cost), in the caller's layout
// ---------------------------------------------------------------------------------------------------
__global__ void fast_merge_ssd_kernel(const int32_t* __restrict__ PART, FastGeom g, const uint8_t* __restrict__ A,
                                      size_t a_step, void* disp_out, size_t disp_step, int elem, void* best_out,
                                      size_t best_step) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int yy = blockIdx.y;
    if (x >= g.cols) return;
    int bestc = INT_MAX, bestd = 0;
    bool found = false;
    const uint32_t thresh = key_invalid_threshold(g.R), bias = key_bias(g.R);
    for (int grp = 0; grp < g.G; ++grp) {
        const uint32_t key = uint32_t(PART[(size_t(grp) * g.nrows + yy) * g.wpart + x]);
        if (key >= thresh) continue;                               // no legal candidate in this group
        const uint32_t qlo = uint32_t(x + g.dmin + FGROUP * grp + g.eoff);   // position of the group's first candidate
        const uint32_t q2 = qlo + ((key - qlo) & uint32_t(FGROUP - 1));
        const int c = int(key - q2 - bias) >> FKEY_BITS;           // ER - 2C (exact: multiple of 128)
        if (!found || c < bestc) { bestc = c; bestd = int(q2) - g.eoff - x; found = true; }
    }
    int cost = 99999999;                                           // DisparitySSD.cpp:37
    // EL(x), the window energy of the reference image (replicate padding), is only needed to report
    // the cost and to honour the 99999999 threshold; with (2R+1)^2 * 255^2 < 99999999 (R <= 19) the
    // threshold can never bind, so the energy is computed only when the caller asked for costs.
    if (found && best_out) {
        int el = 0;
        const int y = g.rb + yy;
        for (int wy = -g.R; wy <= g.R; ++wy) {
            const uint8_t* row = A + size_t(clampi(clampi(y + wy, 0, g.rows - 1), g.ar0, g.ar1 - 1)) * a_step;
            for (int wx = -g.R; wx <= g.R; ++wx) { const int v = row[clampi(x + wx, 0, g.cols - 1)]; el += v * v; }
        }
        cost = bestc + el;
    }
    char* drow = reinterpret_cast<char*>(disp_out) + size_t(yy) * disp_step;
    if (elem == 1) reinterpret_cast<int8_t*>(drow)[x] = int8_t(uint8_t(uint32_t(bestd) & 0xFFu));
    else if (elem == 2) reinterpret_cast<int16_t*>(drow)[x] = int16_t(bestd);
    else reinterpret_cast<int32_t*>(drow)[x] = bestd;
    if (best_out) reinterpret_cast<int32_t*>(reinterpret_cast<char*>(best_out) + size_t(yy) * best_step)[x] = cost;
}

// NCC: winning key per group -> first maximum over the groups -> disparity with the reference's
// alignment rule (DisparityNCorr.cpp:67) and, on request, the winning score recomputed exactly with
// TM_CCORR_NORMED's arithmetic (float32 numerator, double energies; see ncorr_exact_kernel).
__global__ void fast_merge_ncc_kernel(const int32_t* __restrict__ PART, FastGeom g, const uint8_t* __restrict__ A,
                                      size_t a_step, const uint8_t* __restrict__ B, size_t b_step, void* disp_out,
                                      size_t disp_step, int elem, void* best_out, size_t best_step) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int yy = blockIdx.y;
    if (x >= g.cols) return;
    long long bestv = -1; int bestd = 0;
    for (int grp = 0; grp < g.G; ++grp) {
        const uint32_t key = uint32_t(PART[(size_t(grp) * g.nrows + yy) * g.wpart + x]);
        if (key == NCC_KEY_NONE) continue;                          // no legal candidate in this group
        const long long v = key >> NCC_KEY_SHIFT;                   // same strip row -> same magic -> comparable
        if (v > bestv) { bestv = v; bestd = g.dmin + FGROUP * grp + (FGROUP - 1 - int((key >> 2) & (FGROUP - 1))); }
    }
    const int centre = x + bestd;                                   // winning window centre (unpadded column)
    const int startc = max(0, x + g.dmin), endc = min(g.cols - 1, x + g.dmax);
    const bool right_aligned = (g.dmin <= 0 && g.dmax <= 0);
    const int disp = (centre - startc) - (right_aligned ? endc - startc : 0);
    char* drow = reinterpret_cast<char*>(disp_out) + size_t(yy) * disp_step;
    if (elem == 1) reinterpret_cast<int8_t*>(drow)[x] = int8_t(uint8_t(uint32_t(disp) & 0xFFu));
    else if (elem == 2) reinterpret_cast<int16_t*>(drow)[x] = int16_t(disp);
    else reinterpret_cast<int32_t*>(drow)[x] = disp;
    if (best_out) {
        const int y = g.rb + yy;
        int c = 0, el = 0, er = 0;
        for (int wy = -g.R; wy <= g.R; ++wy) {
            const int wr = clampi(clampi(y + wy, 0, g.rows - 1), g.ar0, g.ar1 - 1);
            const uint8_t* arow = A + size_t(wr) * a_step;
            const uint8_t* brow = B + size_t(wr) * b_step;
            for (int wx = -g.R; wx <= g.R; ++wx) {
                const int l = arow[clampi(x + wx, 0, g.cols - 1)], r = brow[clampi(centre + wx, 0, g.cols - 1)];
                c += l * r; el += l * l; er += r * r;
            }
        }
        double num = double(float(c));
        const double wnd = double(er);
        const double lim = fmin(0.5, 10 * double(FLT_EPSILON) * wnd);
        const double t = (wnd <= lim) ? 0 : sqrt(wnd) * sqrt(double(el));
        if (fabs(num) < t) num /= t;
        else if (fabs(num) < t * 1.125) num = num > 0 ? 1 : -1;
        else num = 0;
        reinterpret_cast<float*>(reinterpret_cast<char*>(best_out) + size_t(yy) * best_step)[x] = float(num);
    }
}

// ---------------------------------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------------------------------
static inline bool fast_supported(const Problem& p) {
    if (p.R > FMAXR) return false;
    if (p.cols < 1 || p.rows < 1) return false;
    if (p.dmax - p.dmin + 1 > 4096) return false;
    // positions must fit the key arithmetic comfortably
    if (p.cols + 2 * (p.dmax - p.dmin) + 1024 > (1 << 20)) return false;
    return true;
}

// Strip width: 24 pixels per thread by default; STEREO_FAST_K=16 selects the narrower variant (debug knob).
static inline int fast_pick_k(const Problem& p) {
    static const int forced = [] { const char* e = getenv("STEREO_FAST_K"); return e ? atoi(e) : 0; }();
    if (forced == 16 && (p.R == 4 || p.R == 5)) return 16;
    return FK_DEFAULT;
}

static inline void fast_geometry(const stereo_ctx* ctx, const Problem& p, FastGeom& g) {
    g.rows = p.rows; g.cols = p.cols; g.R = p.R; g.dmin = p.dmin; g.dmax = p.dmax; g.cost = p.cost;
    g.rb = p.row_begin; g.re = p.row_end; g.nrows = g.re - g.rb;
    g.ar0 = p.avail_begin; g.ar1 = p.avail_end;
    const int D = p.dmax - p.dmin + 1;
    g.K = fast_pick_k(p);
    g.G = (D + FGROUP - 1) / FGROUP;
    g.gc = (g.G % 2 == 0) ? 2 : 1;
    g.nw = (g.K == 16) ? 12 : 8;
    g.spc = g.nw / g.gc;
    g.nstrips = (p.cols + g.K - 1) / g.K;
    g.tilesX = (g.nstrips + g.spc - 1) / g.spc;
    g.gblocks = g.G / g.gc;
    const int w = 2 * p.R + 1;
    g.base_y = floor_div(g.rb - w, FRPS) * FRPS;
    g.J = round_up(g.re - g.base_y, FRPS);
    if (p.cost == STEREO_COST_SSD) { g.cmin = -p.R; g.cmax = p.cols - 1 + p.R; }
    else { g.cmin = 0; g.cmax = p.cols - 1; }
    // RQ column q = e + qoff with e = x0 + dl + R + (c+m); first index must be >= 0 and 4-aligned
    int qo = -(p.dmin + p.R); if (qo < 0) qo = 0;
    while (((p.dmin + p.R + qo) & 3) != 0) ++qo;
    g.qoff = qo;
    int eo = -p.dmin; if (eo < 0) eo = 0;
    while (((p.dmin + eo) & 3) != 0) ++eo;
    g.eoff = eo;
    const int tile_px = g.spc * g.K;
    g.lpw = round_up(tile_px + 2 * p.R, 4);
    g.rqw = round_up(tile_px + 2 * p.R + FGROUP * g.gc + FM, 4);
    g.e2w = round_up(tile_px + FGROUP * g.gc + FM, 4);
    g.wpart = g.tilesX * tile_px;
    g.lp_pitch = round_up((g.tilesX - 1) * tile_px + g.lpw, 64);
    const int last_p0 = (g.tilesX - 1) * tile_px;
    const int gmax = FGROUP * (g.gblocks - 1) * g.gc;
    g.rq_pitch = round_up(last_p0 + p.dmin + gmax + p.R + g.qoff + g.rqw, 64);
    g.e2_pitch = round_up(last_p0 + p.dmin + gmax + g.eoff + g.e2w, 64);
    g.total = (long long)g.tilesX * g.gblocks * g.nrows;
    // grid: one CTA per SM, but keep segments long enough that the (2R+1)-row warm-up stays small
    long long min_rows = 4LL * w; if (min_rows < 32) min_rows = 32;
    long long ctas = g.total / min_rows; if (ctas < 1) ctas = 1;
    if (ctas > ctx->sm_count) ctas = ctx->sm_count;
    g.L = (g.total + ctas - 1) / ctas;
    g.ctas = int((g.total + g.L - 1) / g.L);
}

static inline size_t fast_smem_bytes(const FastGeom& g) {
    const size_t stage_words = size_t(FRPS) * g.lpw + size_t(FRPS / 2) * g.rqw + size_t(FRPS) * g.e2w;
    return size_t(FNST) * stage_words * 4 + 2 * FNST * 8 + 16;
}

static inline size_t fast_scratch_bytes(stereo_ctx* ctx, const Problem& p) {
    FastGeom g; fast_geometry(ctx, p, g);
    size_t b = 0;
    auto add = [&](size_t bytes) { b += ((bytes + 255) & ~size_t(255)); };
    add(size_t(g.J) * g.lp_pitch * 4);
    add(size_t(g.J / 2) * g.rq_pitch * 4);
    add(size_t(g.J) * g.e2_pitch * 4);
    add(size_t(g.G) * g.nrows * g.wpart * 4);
    add(size_t(g.nrows) * round_up(g.e2_pitch + 2 * g.R + 256, 64) * 4);
    if (p.cost == STEREO_COST_NCORR) { add(size_t(g.J) * g.e2_pitch * 4); add(size_t(g.tilesX) * g.spc * g.nrows * 4); }
    return b + 4096;
}

typedef void (*fast_kernel_fn)(const FastKernelParams);
template <int COST>
static inline fast_kernel_fn fast_pick_cost(int R, int K) {
    if (K == 16) {
        switch (R) {
        case 4: return fast_cost_kernel<4, 16, 12, COST>;
        case 5: return fast_cost_kernel<5, 16, 12, COST>;
        }
        return nullptr;
    }
    switch (R) {
    case 0: return fast_cost_kernel<0, 24, 8, COST>;
    case 1: return fast_cost_kernel<1, 24, 8, COST>;
    case 2: return fast_cost_kernel<2, 24, 8, COST>;
    case 3: return fast_cost_kernel<3, 24, 8, COST>;
    case 4: return fast_cost_kernel<4, 24, 8, COST>;
    case 5: return fast_cost_kernel<5, 24, 8, COST>;
    case 6: return fast_cost_kernel<6, 24, 8, COST>;
    case 7: return fast_cost_kernel<7, 24, 8, COST>;
    }
    return nullptr;
}
static inline fast_kernel_fn fast_pick(int cost, int R, int K) {
    return cost == STEREO_COST_SSD ? fast_pick_cost<STEREO_COST_SSD>(R, K) : fast_pick_cost<STEREO_COST_NCORR>(R, K);
}

static inline int fast_ctx_init(stereo_ctx*) {
    for (int cost = 0; cost <= 1; ++cost)
        for (int K = 16; K <= 24; K += 8)
            for (int R = 0; R <= FMAXR; ++R) {
                fast_kernel_fn fn = fast_pick(cost, R, K);
                if (!fn) continue;
                cudaError_t e = cudaFuncSetAttribute(reinterpret_cast<const void*>(fn),
                                                     cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
                if (e != cudaSuccess) { set_error("cudaFuncSetAttribute(smem): %s", cudaGetErrorString(e)); return STEREO_ERR_CUDA; }
            }
    return STEREO_OK;
}

static inline int run_fast(stereo_ctx* ctx, const Problem& p, cudaStream_t st) {
    FastGeom g; fast_geometry(ctx, p, g);
    FastArrays a{};
    a.LP = static_cast<int32_t*>(ctx->arena.take(size_t(g.J) * g.lp_pitch * 4));
    a.RQ = static_cast<uint32_t*>(ctx->arena.take(size_t(g.J / 2) * g.rq_pitch * 4));
    a.E2 = static_cast<int32_t*>(ctx->arena.take(size_t(g.J) * g.e2_pitch * 4));
    a.PART = static_cast<int32_t*>(ctx->arena.take(size_t(g.G) * g.nrows * g.wpart * 4));
    a.V = static_cast<int32_t*>(ctx->arena.take(size_t(g.nrows) * round_up(g.e2_pitch + 2 * g.R + PE_COLS, 64) * 4));
    if (p.cost == STEREO_COST_NCORR) {
        a.RS = static_cast<float*>(ctx->arena.take(size_t(g.J) * g.e2_pitch * 4));
        a.SC = static_cast<float*>(ctx->arena.take(size_t(g.tilesX) * g.spc * g.nrows * 4));
    }
    if (!a.LP || !a.RQ || !a.E2 || !a.PART || !a.V || (p.cost == STEREO_COST_NCORR && (!a.RS || !a.SC))) {
        set_error("scratch arena too small (internal)"); return STEREO_ERR_ALLOC;
    }
    const uint8_t* A = static_cast<const uint8_t*>(p.ref.ptr);
    const uint8_t* B = static_cast<const uint8_t*>(p.tgt.ptr);
    const dim3 tb(256);
    prep_lp_kernel<<<dim3(div_round_up(g.lp_pitch / 4, 256), g.J), tb, 0, st>>>(A, p.ref.step, g, a.LP);
    prep_rq_kernel<<<dim3(div_round_up(g.rq_pitch / 4, 256), g.J / 2), tb, 0, st>>>(B, p.tgt.step, g, a.RQ);
    const int vpitch = round_up(g.e2_pitch + 2 * g.R + PE_COLS, 64);
    prep_v_kernel<<<dim3(div_round_up(vpitch, 128), div_round_up(g.nrows, PV_ROWS)), 128, 0, st>>>(
        B, p.tgt.step, g, a.V, vpitch, -g.eoff + g.R);
    const size_t pe_smem = size_t(PE_ROWS) * (PE_COLS + 2 * g.R) * sizeof(int);
    prep_e2_kernel<<<dim3(div_round_up(g.e2_pitch, PE_COLS), div_round_up(g.nrows, PE_ROWS)), PE_COLS, pe_smem, st>>>(
        a.V, vpitch, g, a.E2, a.RS);
    const bool ncc = p.cost == STEREO_COST_NCORR;
    if (ncc) {   // window energies of the reference image -> per strip-row key binade (V is free again after prep_e2)
        prep_v_kernel<<<dim3(div_round_up(vpitch, 128), div_round_up(g.nrows, PV_ROWS)), 128, 0, st>>>(
            A, p.ref.step, g, a.V, vpitch, g.R);
        prep_scale_kernel<<<dim3(div_round_up(g.tilesX * g.spc, 128), g.nrows), 128, 0, st>>>(a.V, vpitch, g, a.SC);
        ctx->last_launches += 2;
    }
    FastKernelParams kp{g, a.LP, a.RQ, ncc ? reinterpret_cast<const int32_t*>(a.RS) : a.E2, a.PART, a.SC};
    const int hot = ctx->hot_used < stereo_ctx::HOT_EVENTS ? ctx->hot_used : -1;
    if (hot >= 0) cudaEventRecord(ctx->hot0[hot], st);
    fast_pick(p.cost, p.R, g.K)<<<g.ctas, g.nw * 32, fast_smem_bytes(g), st>>>(kp);
    if (hot >= 0) { cudaEventRecord(ctx->hot1[hot], st); ctx->hot_used++; }
    ctx->hot_total++;
    if (ncc)
        fast_merge_ncc_kernel<<<dim3(div_round_up(g.cols, 128), g.nrows), 128, 0, st>>>(
            a.PART, g, A, p.ref.step, B, p.tgt.step, p.disp.ptr, p.disp.step, p.disp.elem, p.best.ptr, p.best.step);
    else
        fast_merge_ssd_kernel<<<dim3(div_round_up(g.cols, 128), g.nrows), 128, 0, st>>>(
            a.PART, g, A, p.ref.step, p.disp.ptr, p.disp.step, p.disp.elem, p.best.ptr, p.best.step);
    ctx->last_launches += 6;
    SB_CUDA(cudaGetLastError());
    return STEREO_OK;
}

} // namespace sb
