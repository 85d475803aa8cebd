// Packed u8 path, host side + the streaming kernels around the hot kernel: operand preparation (prep_*),
// merge of the per-group winners, geometry and launch.  The hot kernel itself is in fast_kernel.cuh and is
// instantiated in fast_inst.cu.
#pragma once
#include "fast_kernel.cuh"
#include "exact.cuh"

namespace sb {

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
// One launch for the stages that feed the hot kernel
// ---------------------------------------------------------------------------------------------------
// A 511x640 pair spends 4-8 us in each of these stages and ~3 us between two dependent launches - as much as in its hot
// kernel - and the stages do not depend on each other: their grids are laid end to end along blockIdx.x of ONE launch
// (the longest-running stage first).
constexpr int PREP_STAGES = 4;
struct PrepStages {
    unsigned gx[PREP_STAGES], gy[PREP_STAGES], gz[PREP_STAGES];   // grid of stage i (gx = 0: absent)
    unsigned first[PREP_STAGES + 1];  // first flat CTA index of stage i; first[PREP_STAGES] = CTAs of the launch
    int rows_per_cta;                 // prep_tgt stage (float launches: rows of the AF / BF arrays)
    int border_ppr_log2;              // fused_border stage: log2(pixels per row of a CTA)
    unsigned fill_words;              // fill stage: 16-byte words of one partner's partial-key map
    int fill_value;                   // ... and the word they are set to
};
static inline void prep_stage_set(PrepStages& s, int i, unsigned gx, unsigned gy, unsigned gz) { s.gx[i] = gx; s.gy[i] = gy; s.gz[i] = gz; }
static inline unsigned prep_stage_finish(PrepStages& s) {
    unsigned at = 0;
    for (int i = 0; i < PREP_STAGES; ++i) { s.first[i] = at; at += s.gx[i] * s.gy[i] * s.gz[i]; }
    s.first[PREP_STAGES] = at;
    return at;
}
// STEREO_PREP_SPLIT=1 (profiling knob): every stage as its own launch of the same kernel, so a launch list times them apart.
template <typename Fn>
static inline int prep_launch(Fn fn, const FastKernelParams& kp, PrepStages& stg, cudaStream_t st) {
    static const bool split = [] { const char* e = getenv("STEREO_PREP_SPLIT"); return e && atoi(e) != 0; }();
    if (!split) { fn<<<prep_stage_finish(stg), 256, 0, st>>>(kp, stg); return 1; }
    int launches = 0;
    for (int i = 0; i < PREP_STAGES; ++i) {
        if (!stg.gx[i]) continue;
        PrepStages one = stg;
        for (int k = 0; k < PREP_STAGES; ++k) if (k != i) one.gx[k] = one.gy[k] = one.gz[k] = 0;
        fn<<<prep_stage_finish(one), 256, 0, st>>>(kp, one);
        ++launches;
    }
    return launches;
}
__device__ __forceinline__ uint3 prep_stage_block(const PrepStages& s, int i) {
    const unsigned b = blockIdx.x - s.first[i];
    return make_uint3(b % s.gx[i], (b / s.gx[i]) % s.gy[i], b / (s.gx[i] * s.gy[i]));
}

// ---------------------------------------------------------------------------------------------------
// Prep kernels: operand rows in the layout the hot loop consumes
// ---------------------------------------------------------------------------------------------------
// Every prep / merge kernel takes the launch's parameter block; the z index of its grid selects the job.  The stages that
// do not depend on each other are bodies taking their block index as an argument: prep_u8_kernel / prep_f32_kernel below lay
// their grids end to end in ONE launch.
//
// LP[j][p]: 4 columns per thread, 16-byte stores.  256 threads.
// Fused pairs: every strip merges into the partner's partial keys with RED.MIN (NCC: RED.MAX), which therefore start from
// "no candidate".  As a stage of the prep launch the stores ride under the other stages' latency (a memset in front of the
// launch: 270 MB = 45 us for four 4K pairs).  A CTA sets 256 x 4 16-byte words of partner bid.z's map.
constexpr unsigned FILL_WORDS_PER_CTA = 256 * 4;
__device__ __forceinline__ void prep_fill_body(const FastKernelParams& P, const uint3 bid, const PrepStages& s) {
    int4* __restrict__ dst = reinterpret_cast<int4*>(P.job[P.g.npairs + bid.z].PART);
    const int4 v = make_int4(s.fill_value, s.fill_value, s.fill_value, s.fill_value);
    const unsigned w0 = bid.x * FILL_WORDS_PER_CTA + threadIdx.x;
#pragma unroll
    for (int i = 0; i < 4; ++i) { const unsigned w = w0 + i * 256; if (w < s.fill_words) dst[w] = v; }
}

// A CTA takes LP_ROWS rows: every load of all of them is issued before the first store (a plain row loop waits for one
// memory round trip per row - the stores keep the next row's loads from moving up - and one row per CTA keeps only 8 bytes
// per thread in flight: 99 us for four 4K pairs where the bytes need 30).
#ifndef SB_LP_ROWS
#define SB_LP_ROWS 4
#endif
constexpr int LP_ROWS = SB_LP_ROWS;
__device__ __forceinline__ void prep_lp_body(const FastKernelParams& P, const uint3 bid) {
    const FastGeom& g = P.g;
    const FastJob& job = P.job[bid.z];
    const uint8_t* __restrict__ A = job.A; const size_t step = job.a_step;
    int32_t* __restrict__ LP = job.LP;
    const int p4 = (bid.x * 256 + threadIdx.x) * 4;
    if (p4 >= g.lp_pitch) return;
    // Fused pair launches: padded columns past the end of the row (the windows of the partner direction's candidates
    // centred in the right padding, DisparitySSD.cpp:39-40,50) alias the next padded row, exactly as in the
    // extended TARGET image of an unfused launch (bext).  The direction's own pixels never read those columns.
    const bool ext = g.npairs > 0 && p4 + 3 >= g.cols + 2 * g.R;
    int col[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) col[t] = clampi(p4 + t - g.R, 0, g.cols - 1);
    const int j0 = int(bid.y) * LP_ROWS;
    int lnew[LP_ROWS][4], lold[LP_ROWS][4];
#pragma unroll
    for (int r = 0; r < LP_ROWS; ++r) {
        const int y = g.base_y + min(j0 + r, g.J - 1);
        if (ext) {
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                lnew[r][t] = bext(A, step, g.rows, g.cols, g.R, y + g.R, p4 + t + g.R, g.ar0, g.ar1);
                lold[r][t] = bext(A, step, g.rows, g.cols, g.R, y - g.R - 1, p4 + t + g.R, g.ar0, g.ar1);
            }
        } else {
            const uint8_t* rnew = A + size_t(clampi(clampi(y + g.R, 0, g.rows - 1), g.ar0, g.ar1 - 1)) * step;
            const uint8_t* rold = A + size_t(clampi(clampi(y - g.R - 1, 0, g.rows - 1), g.ar0, g.ar1 - 1)) * step;
#pragma unroll
            for (int t = 0; t < 4; ++t) { lnew[r][t] = rnew[col[t]]; lold[r][t] = rold[col[t]]; }
        }
    }
    const bool ssd = g.cost == STEREO_COST_SSD;
#pragma unroll
    for (int r = 0; r < LP_ROWS; ++r) {
        if (j0 + r >= g.J) break;
        int v[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            // SSD: (-l_new, +l_old) so the running sums hold -C; NCC: (+l_new, -l_old), sums hold +C
            const int a_new = ssd ? -lnew[r][t] : lnew[r][t], a_old = ssd ? lold[r][t] : -lold[r][t];
            v[t] = int(uint32_t(uint16_t(int16_t(a_new))) | (uint32_t(uint16_t(int16_t(a_old))) << 16));
        }
        *reinterpret_cast<int4*>(LP + size_t(j0 + r) * g.lp_pitch + p4) = make_int4(v[0], v[1], v[2], v[3]);
    }
}

// RQ[jp][q]: 4 positions per thread, 16-byte stores.
__global__ void __launch_bounds__(256) prep_rq_kernel(const __grid_constant__ FastKernelParams P) {
    const FastGeom& g = P.g;
    const FastJob& job = P.job[blockIdx.z];
    const uint8_t* __restrict__ B = job.B; const size_t step = job.b_step;
    uint32_t* __restrict__ RQ = job.RQ;
    const int q4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const int jp = blockIdx.y;
    if (q4 >= g.rq_pitch || !RQ) return;
    const int ye = g.base_y + 2 * jp;
    uint32_t v[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        const int e = q4 + t - job.qoff;
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
// of_ref = 0: the job's target image, V column x <-> e = x - eoff + R;  of_ref = 1 (NCC scale pass): the job's
// reference image, V column x <-> e = x + R (the padded column x).
__global__ void __launch_bounds__(128) prep_v_kernel(const __grid_constant__ FastKernelParams P, int vpitch, int of_ref) {
    const FastGeom& g = P.g;
    const FastJob& job = P.job[blockIdx.z];
    const uint8_t* __restrict__ B = of_ref ? job.A : job.B; const size_t step = of_ref ? job.a_step : job.b_step;
    int32_t* __restrict__ V = job.V;
    const int vbase = of_ref ? g.R : -job.eoff + g.R;
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
__global__ void __launch_bounds__(PE_COLS) prep_e2_kernel(const __grid_constant__ FastKernelParams P, int vpitch) {
    const FastGeom& g = P.g;
    const FastJob& job = P.job[blockIdx.z];
    const int32_t* __restrict__ V = job.V;
    int32_t* __restrict__ E2 = job.E2; float* __restrict__ RS = job.RS;
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
    const int uc = q2 - job.eoff;
    const bool valid = uc >= job.cmin && uc <= job.cmax;
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

// Fused target-side prep (replaces prep_rq + prep_v + prep_e2 on the default path): ONE pass over the
// extended target image produces the packed RQ rows AND the window-energy rows E2 / RS.
//   * a thread owns one extended column (V column x <-> ext column e = x - eoff + R <-> RQ column x + delta,
//     delta = qoff - eoff + R, a multiple of 4 and >= 0) and walks down PT_ROWS operand rows of its CTA;
//     the four bytes of an RQ word (entering / leaving rows of two consecutive step rows) are exactly the
//     bytes the vertical running sum of squares needs, so each image byte is fetched once per role;
//   * every 4 rows the vertical sums go through a double-buffered shared-memory tile and come back as
//     horizontal (2R+1)-sums, 4 adjacent centres per thread (LDS.128 in, sliding adds, STG.128 out).
// A CTA of 256 columns yields PT_TS(R) = (256 - 2R) & ~3 centre columns.
constexpr int PT_THREADS = 256;
constexpr int PT_ROWS = 64;                   // operand rows per CTA (multiple of 4; the host halves it for small grids)
constexpr int PT_VSTRIDE = PT_THREADS + 16;   // words per shared-memory row (slack for the sliding window reads)
__host__ __device__ constexpr int pt_ts(int R) { return (PT_THREADS - 2 * R) & ~3; }

template <int R, bool INTERIOR>
__device__ __forceinline__ void prep_tgt_body(const FastGeom& g, const FastJob& job, const uint3 bid, int rows_per_cta, int (&vs)[2][4][PT_VSTRIDE]) {
    const uint8_t* __restrict__ B = job.B; const size_t step = job.b_step;
    uint32_t* __restrict__ RQ = job.RQ;
    constexpr int TS = pt_ts(R);
    constexpr int NV = 2 * R + 4;              // vertical sums a thread of the horizontal stage touches
    constexpr int NVEC = (NV + 3) / 4;
    const int t = threadIdx.x;
    const int delta = job.qoff - job.eoff + R;
    const int x0 = -delta + int(bid.x) * TS;               // first V column of this CTA (multiple of 4)
    const int x = x0 + t;
    // column mapping of bext(): which source column, and whether the flat index falls into a neighbouring row
    const int Wp = g.cols + 2 * R;
    const int c = x - job.eoff;                            // padded column (e - R)
    int shift = 0, scol;
    if (c < 0) { shift = -1; scol = g.cols - 1; }
    else if (c >= Wp) { shift = 1; scol = 0; }
    else scol = clampi(c - R, 0, g.cols - 1);
    const uint8_t* __restrict__ Bc = B + scol;
    const int rows = g.rows, ar0 = g.ar0, ar1 = g.ar1;
    auto load = [&](int i) -> int {                        // == bext(B, step, rows, cols, R, i, x - eoff + R, ar0, ar1)
        if (INTERIOR) return Bc[size_t(i + shift) * step]; // every row the CTA touches (+-1) lies inside the image
        if (shift < 0 && i + R - 1 < 0) return 0;
        if (shift > 0 && i + R + 1 > rows + 2 * R - 1) return 0;
        const int r = clampi(clampi(i + shift, 0, rows - 1), ar0, ar1 - 1);
        return Bc[size_t(r) * step];
    };
    const int j0 = int(bid.y) * rows_per_cta;
    const int j1 = min(g.J, j0 + rows_per_cta);
    const bool ncc = g.cost != STEREO_COST_SSD;
    // vertical sum of squares of the row above the first one
    int v = 0;
    {
        const int y = g.base_y + j0 - 1;
#pragma unroll
        for (int i = -R; i <= R; ++i) { const int b = load(y + i); v += b * b; }
    }
    const int q = x + delta;                               // RQ column
    const bool rq_ok = t < TS && q < g.rq_pitch && RQ != nullptr;   // q >= 0 by construction; fused partner jobs have no RQ
    // horizontal stage: thread -> (row hr of the group, 4 centres from column c4)
    const int hr = t >> 6, c4 = (t & 63) * 4;
    const int q2 = x0 + c4;
    const bool h_ok = c4 < TS && q2 >= 0 && q2 < g.e2_pitch;
    const bool all_valid = (q2 - job.eoff >= job.cmin) && (q2 + 3 - job.eoff <= job.cmax);
    const uint32_t bias = key_bias(R);
    // entering rows (ye+R, ye+1+R) and leaving rows (ye-R-1, ye-R) of the two step-row pairs of a 4-row group
    const uint8_t* pn = nullptr; const uint8_t* po = nullptr;
    if (INTERIOR) {
        pn = Bc + size_t(g.base_y + j0 + R + shift) * step;
        po = Bc + size_t(g.base_y + j0 - R - 1 + shift) * step;
    }
    // two 4-row groups of source bytes are in flight ahead of the one being summed (16 bytes per thread: with one group
    // the stage moved 1.4 TB/s - 32 warps x 32 lanes x 8 bytes per SM against ~0.8 us of latency)
    int nb[8], nb2[8];
    auto fetch = [&](int jj, int (&dst)[8]) {
        if (INTERIOR) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                dst[4 * h] = pn[0]; dst[4 * h + 2] = pn[step]; dst[4 * h + 1] = po[0]; dst[4 * h + 3] = po[step];
                pn += 2 * step; po += 2 * step;
            }
        } else {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int ye = g.base_y + jj + 2 * h;
                dst[4 * h] = load(ye + R); dst[4 * h + 1] = load(ye - R - 1); dst[4 * h + 2] = load(ye + 1 + R); dst[4 * h + 3] = load(ye - R);
            }
        }
    };
    // vertical stage of one 4-row group: RQ words, running sums of squares into the shared-memory tile
    auto vertical = [&](const int jj, const int buf, const int (&cb)[8]) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int b0 = cb[4 * h], b1 = cb[4 * h + 1], b2 = cb[4 * h + 2], b3 = cb[4 * h + 3];
            if (rq_ok) RQ[size_t((jj >> 1) + h) * g.rq_pitch + q] = uint32_t(b0) | (uint32_t(b1) << 8) | (uint32_t(b2) << 16) | (uint32_t(b3) << 24);
            v += b0 * b0 - b1 * b1;
            vs[buf][2 * h][t] = v;
            v += b2 * b2 - b3 * b3;
            vs[buf][2 * h + 1][t] = v;
        }
    };
    fetch(j0, nb);
    if (j0 + 4 < j1) fetch(j0 + 4, nb2);
    // two groups per trip, each with its own registers: a refill is issued as soon as its group has been summed and stays in
    // flight across two barriers and horizontal stages (moving a group between register sets would wait for its loads)
    for (int jj0 = j0; jj0 < j1; jj0 += 8) {
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int jj = jj0 + 4 * half, buf = half;
        if (jj >= j1) break;
        if (half == 0) { vertical(jj, buf, nb); if (jj + 8 < j1) fetch(jj + 8, nb); }
        else           { vertical(jj, buf, nb2); if (jj + 8 < j1) fetch(jj + 8, nb2); }
        __syncthreads();
        const int y = g.base_y + jj + hr;
        if (h_ok && y >= g.rb && y < g.re) {
            int w[NVEC * 4];
#pragma unroll
            for (int i = 0; i < NVEC; ++i) {
                const int4 u = *reinterpret_cast<const int4*>(&vs[buf][hr][c4 + 4 * i]);
                w[4 * i] = u.x; w[4 * i + 1] = u.y; w[4 * i + 2] = u.z; w[4 * i + 3] = u.w;
            }
            int er[4];
            er[0] = 0;
#pragma unroll
            for (int i = 0; i <= 2 * R; ++i) er[0] += w[i];
#pragma unroll
            for (int k = 1; k < 4; ++k) er[k] = er[k - 1] + w[2 * R + k] - w[k - 1];
            const size_t o = size_t(jj + hr) * g.e2_pitch + q2;
            if (!ncc) {
                int key[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int uc = q2 + k - job.eoff;
                    const bool valid = all_valid || (uc >= job.cmin && uc <= job.cmax);
                    key[k] = valid ? int(bias + (uint32_t(er[k]) << FKEY_BITS) + uint32_t(q2 + k)) : int(KEY_INVALID);
                }
                *reinterpret_cast<int4*>(job.E2 + o) = make_int4(key[0], key[1], key[2], key[3]);
            } else {
                float rs[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int uc = q2 + k - job.eoff;
                    const bool valid = all_valid || (uc >= job.cmin && uc <= job.cmax);
                    rs[k] = (valid && er[k] > 0) ? float(1.0 / sqrt(double(er[k]))) : 0.f;
                }
                *reinterpret_cast<float4*>(job.RS + o) = make_float4(rs[0], rs[1], rs[2], rs[3]);
            }
        }
      }
    }
}

template <int R>
__device__ __forceinline__ void prep_tgt_stage(const FastKernelParams& P, const uint3 bid, int rows_per_cta, int (&vs)[2][4][PT_VSTRIDE]) {
    const FastGeom& g = P.g;
    const FastJob& job = P.job[bid.z];
    // rows this CTA reads: base_y + j0 - R - 2 .. base_y + j1 + R (one more either side for the row-wrap columns)
    const int j0 = int(bid.y) * rows_per_cta, j1 = min(g.J, j0 + rows_per_cta);
    const int ilo = g.base_y + j0 - R - 2, ihi = g.base_y + j1 + R;
    const bool interior = ilo >= max(0, g.ar0) && ihi <= min(g.rows, g.ar1) - 1;
    if (interior) prep_tgt_body<R, true>(g, job, bid, rows_per_cta, vs);
    else prep_tgt_body<R, false>(g, job, bid, rows_per_cta, vs);
}

// NCC: per strip (K pixels) and output row, the power of two just above sqrt(max EL) — the binade the
// fixed-point keys of that strip row live in.  V holds the vertical (2R+1)-sums of squares of the
// replicate-padded REFERENCE image: V column c = padded column c, so EL(x) = sum V[yy][x .. x+2R].
__global__ void __launch_bounds__(128) prep_scale_kernel(const __grid_constant__ FastKernelParams P, int vpitch) {
    const FastGeom& g = P.g;
    const FastJob& job = P.job[blockIdx.z];
    const int32_t* __restrict__ V = job.V;
    float* __restrict__ SC = job.SC;
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

// The same for launches that hold BOTH directions of every image pair (jobs i and i + npairs mirror each other): the
// reference image of a job is its partner's target image, whose window energies prep_tgt_kernel has just turned into
// the partner's RS rows (1/sqrt(E), 0 for E = 0 or an illegal centre).  max EL over a strip = 1 / min nonzero RS, so the
// separate vertical pass over the reference image (prep_v_kernel) is not needed.
__global__ void __launch_bounds__(128) prep_scale_pair_kernel(const __grid_constant__ FastKernelParams P, int npairs) {
    const FastGeom& g = P.g;
    const int z = blockIdx.z;
    const FastJob& job = P.job[z];
    const FastJob& partner = P.job[z < npairs ? z + npairs : z - npairs];
    float* __restrict__ SC = job.SC;
    const int strip = blockIdx.x * blockDim.x + threadIdx.x;
    const int yy = blockIdx.y;
    if (strip >= g.tilesX * g.spc) return;
    const int x0 = strip * g.K;
    float magic = 1.f;
    if (x0 < g.cols) {
        const float* rs = partner.RS + size_t(g.rb + yy - g.base_y) * g.e2_pitch + partner.eoff + x0;
        const int x1 = min(g.K, g.cols - x0);
        float rsmin = 0.f;                                   // min over the nonzero entries
        for (int x = 0; x < x1; ++x) { const float v = rs[x]; if (v > 0.f && (rsmin == 0.f || v < rsmin)) rsmin = v; }
        if (rsmin > 0.f) {
            // smallest power of two strictly above sqrt(elmax) * (1 + 2^-20); 1/rs is sqrt(E) to within 2^-23 relative
            // (float operands: 2^-10 of slack for the float32 running sums)
            const float bound = float((1.0 / double(rsmin)) * (1.0 + (g.opf ? 1.0 / 1024.0 : 1.0 / 1048576.0)));
            int e; frexpf(bound, &e);
            magic = ldexpf(1.f, e);
        }
    }
    if (g.opf) magic *= 3.f;        // C may be negative: keys live in [2 * 2^e, 4 * 2^e) (prep_scale_f_kernel)
    SC[size_t(strip) * g.nrows + yy] = magic;
}

// ---------------------------------------------------------------------------------------------------
// Float-operand path (general float32 images: the noise / contrast variants of main.cpp:140-153,191-193)
// ---------------------------------------------------------------------------------------------------
// bext() for float images: the value the reference reads at unpadded row i, extended column e (= unpadded column + 2R).
__device__ __forceinline__ float bextf(const uint8_t* __restrict__ B, size_t step, int rows, int cols, int R, int i, int e,
                                       int ar0, int ar1) {
    const int Wp = cols + 2 * R;
    const int c = e - R;                 // padded column, may be < 0 or >= Wp
    int src_row = i, src_col;
    if (c < 0) {                         // previous padded row, right padding
        if (i + R - 1 < 0) return 0.f;
        src_row = i - 1; src_col = cols - 1;
    } else if (c >= Wp) {                // next padded row, left padding
        if (i + R + 1 > rows + 2 * R - 1) return 0.f;
        src_row = i + 1; src_col = 0;
    } else {
        src_col = clampi(c - R, 0, cols - 1);
    }
    src_row = clampi(clampi(src_row, 0, rows - 1), ar0, ar1 - 1);
    return reinterpret_cast<const float*>(B + size_t(src_row) * step)[src_col];
}

// AF[jj][p]: rows of the replicate-padded reference image, array row jj <-> image row base_y - R - 1 + jj, p = padded
// column.  Operand row j of the hot kernel = (entering row AF[j + 2R + 1], leaving row AF[j]).  Fused pair launches: the
// columns past the padded row alias the next padded row exactly like the extended target image does (see prep_lp_kernel).
constexpr int FR_ROWS = 4;            // image rows per CTA of the two stages below
__device__ __forceinline__ void prep_af_body(const FastKernelParams& P, const uint3 bid, const int nrows_total) {
    const FastGeom& g = P.g;
    const FastJob& job = P.job[bid.z];
    const int p4 = (bid.x * 256 + threadIdx.x) * 4;
    if (p4 >= g.lp_pitch) return;
    const bool ext = g.npairs > 0 && p4 + 3 >= g.cols + 2 * g.R;
    const int jj0 = int(bid.y) * FR_ROWS, jj1 = min(nrows_total, jj0 + FR_ROWS);
#pragma unroll
    for (int jj = jj0; jj < jj1; ++jj) {
        const int i = g.base_y - g.R - 1 + jj;
        const float* row = reinterpret_cast<const float*>(job.A + size_t(clampi(clampi(i, 0, g.rows - 1), g.ar0, g.ar1 - 1)) * job.a_step);
        float v[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            v[t] = ext ? bextf(job.A, job.a_step, g.rows, g.cols, g.R, i, p4 + t + g.R, g.ar0, g.ar1)
                       : row[clampi(p4 + t - g.R, 0, g.cols - 1)];
        }
        *reinterpret_cast<float4*>(reinterpret_cast<float*>(job.LP) + size_t(jj) * g.lp_pitch + p4) = make_float4(v[0], v[1], v[2], v[3]);
    }
}

// BF[jj][q]: rows of the extended target image, column q <-> extended column e = q - qoff.
__device__ __forceinline__ void prep_bf_body(const FastKernelParams& P, const uint3 bid, const int nrows_total) {
    const FastGeom& g = P.g;
    const FastJob& job = P.job[bid.z];
    const int q4 = (bid.x * 256 + threadIdx.x) * 4;
    if (q4 >= g.rq_pitch || !job.RQ) return;
    const int jj0 = int(bid.y) * FR_ROWS, jj1 = min(nrows_total, jj0 + FR_ROWS);
#pragma unroll
    for (int jj = jj0; jj < jj1; ++jj) {
        const int i = g.base_y - g.R - 1 + jj;
        float v[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) v[t] = bextf(job.B, job.b_step, g.rows, g.cols, g.R, i, q4 + t - job.qoff, g.ar0, g.ar1);
        *reinterpret_cast<float4*>(reinterpret_cast<float*>(job.RQ) + size_t(jj) * g.rq_pitch + q4) = make_float4(v[0], v[1], v[2], v[3]);
    }
}

// Position / energy rows of the float path.  SSD: E2[j][q2] = q2 for a legal centre, KEY_INVALID otherwise (the key is
// 128 * SSD + position, there is no energy term).  NCC: RS[j][q2] = 1 / sqrt(window energy of the extended target image),
// summed in double like OpenCV's integral images (0 for an illegal centre or an empty window).
// (SSD only; 4 positions per thread, 256 threads.)
__device__ __forceinline__ void prep_e2f_body(const FastKernelParams& P, const uint3 bid) {
    const FastGeom& g = P.g;
    const FastJob& job = P.job[bid.z];
    const int q2 = (bid.x * 256 + threadIdx.x) * 4;
    const int j = bid.y;
    if (q2 >= g.e2_pitch) return;                          // (e2_pitch is a multiple of 4)
    int key[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int uc = q2 + k - job.eoff;
        key[k] = (uc >= job.cmin && uc <= job.cmax) ? q2 + k : int(KEY_INVALID);
    }
    *reinterpret_cast<int4*>(job.E2 + size_t(j) * g.e2_pitch + q2) = make_int4(key[0], key[1], key[2], key[3]);
}

// NCC, float path: RS[j][q2] = 1 / sqrt(window energy of the extended target image at centre q2), separably: a block owns
// RSF_ROWS consecutive operand rows and 256 - 2R centres; every thread keeps the sum of squares of ONE extended column over
// the 2R+1 window rows in double (like OpenCV's double integral images; squares of floats are exact in double) and slides it
// down the block's rows, the centres then add 2R+1 neighbouring column sums from shared memory.
constexpr int RSF_ROWS = 8;
__device__ __forceinline__ void prep_rsf_body(const FastKernelParams& P, const uint3 bid) {
    const FastGeom& g = P.g;
    const FastJob& job = P.job[bid.z];
    __shared__ double vs[RSF_ROWS][256];
    const int R = g.R, ts = 256 - 2 * R;
    const int t = threadIdx.x;
    const int j0 = bid.y * RSF_ROWS;
    const int q20 = bid.x * ts;
    const int e = q20 - job.eoff + R + t;                  // extended column of this thread
    auto sq = [&](int i) -> double { const double b = bextf(job.B, job.b_step, g.rows, g.cols, R, i, e, g.ar0, g.ar1); return b * b; };
    double v = 0;
    {
        const int y = g.base_y + j0;
        for (int wy = -R; wy <= R; ++wy) v += sq(y + wy);
    }
#pragma unroll 1
    for (int r = 0; r < RSF_ROWS; ++r) {
        const int y = g.base_y + j0 + r;
        if (r > 0) v += sq(y + R) - sq(y - R - 1);
        vs[r][t] = v;
    }
    __syncthreads();
    const int q2 = q20 + t;
    if (t >= ts || q2 >= g.e2_pitch) return;
    const int uc = q2 - job.eoff;
    const bool col_ok = uc >= job.cmin && uc <= job.cmax;
#pragma unroll 1
    for (int r = 0; r < RSF_ROWS; ++r) {
        const int j = j0 + r, y = g.base_y + j;
        if (j >= g.J) break;
        float rs = 0.f;
        if (col_ok && y >= g.rb && y < g.re) {
            double er = 0;
            for (int tt = 0; tt <= 2 * R; ++tt) er += vs[r][t + tt];
            rs = er > 0 ? float(1.0 / sqrt(er)) : 0.f;
        }
        job.RS[size_t(j) * g.e2_pitch + q2] = rs;
    }
}

// Float operands.  Stage 0: position rows (SSD) or 1/sqrt(energy) rows (NCC); 1: target rows; 2: reference rows; 3: prep_fill.
__global__ void __launch_bounds__(256) prep_f32_kernel(const __grid_constant__ FastKernelParams P, const __grid_constant__ PrepStages s) {
    if (blockIdx.x < s.first[1]) {
        if (P.g.cost == STEREO_COST_SSD) prep_e2f_body(P, prep_stage_block(s, 0));
        else prep_rsf_body(P, prep_stage_block(s, 0));
    } else if (blockIdx.x < s.first[2]) prep_bf_body(P, prep_stage_block(s, 1), s.rows_per_cta);
    else if (blockIdx.x < s.first[3]) prep_af_body(P, prep_stage_block(s, 2), s.rows_per_cta);
    else prep_fill_body(P, prep_stage_block(s, 3), s);
}

// NCC, float path: window energies of the reference image per pixel (replicate padding, double sums) ...
__global__ void __launch_bounds__(128) prep_elf_kernel(const __grid_constant__ FastKernelParams P, int vpitch) {
    const FastGeom& g = P.g;
    const FastJob& job = P.job[blockIdx.z];
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int yy = blockIdx.y;
    if (x >= g.cols) return;
    const int y = g.rb + yy;
    double el = 0;
    for (int wy = -g.R; wy <= g.R; ++wy) {
        const float* row = reinterpret_cast<const float*>(job.A + size_t(clampi(clampi(y + wy, 0, g.rows - 1), g.ar0, g.ar1 - 1)) * job.a_step);
        for (int wx = -g.R; wx <= g.R; ++wx) { const double a = row[clampi(x + wx, 0, g.cols - 1)]; el += a * a; }
    }
    reinterpret_cast<float*>(job.V)[size_t(yy) * vpitch + x] = float(el);
}
// ... and per strip row the key scale: 3 * 2^e with 2^e just above sqrt(max EL) - |C * rs| <= sqrt(EL) (Cauchy-Schwarz; 2^-10
// of slack for the float32 running sums), so C * rs + 3 * 2^e lies in the binade [2 * 2^e, 4 * 2^e) whatever the sign of C.
__global__ void __launch_bounds__(128) prep_scale_f_kernel(const __grid_constant__ FastKernelParams P, int vpitch) {
    const FastGeom& g = P.g;
    const FastJob& job = P.job[blockIdx.z];
    const int strip = blockIdx.x * blockDim.x + threadIdx.x;
    const int yy = blockIdx.y;
    if (strip >= g.tilesX * g.spc) return;
    const int x0 = strip * g.K;
    float magic = 3.f;
    if (x0 < g.cols) {
        const float* v = reinterpret_cast<const float*>(job.V) + size_t(yy) * vpitch + x0;
        const int x1 = min(g.K, g.cols - x0);
        float elmax = 0.f;
        for (int x = 0; x < x1; ++x) elmax = fmaxf(elmax, v[x]);
        const float bound = float(sqrt(double(elmax)) * (1.0 + 1.0 / 1024.0));
        int e; frexpf(bound, &e);                                       // bound = f * 2^e, f in [0.5, 1)
        magic = elmax > 0.f ? 3.f * ldexpf(1.f, e) : 3.f;
    }
    job.SC[size_t(strip) * g.nrows + yy] = magic;
}

// ---------------------------------------------------------------------------------------------------
// Merge: winning key per group -> disparity (+ cost), in the caller's layout
// ---------------------------------------------------------------------------------------------------
// Stores 4 consecutive disparities of one row (vector store when the caller's layout allows it).
__device__ __forceinline__ void store_disp4(void* disp_out, size_t disp_step, int elem, int yy, int x, int cols, const int (&d)[4]) {
    char* drow = reinterpret_cast<char*>(disp_out) + size_t(yy) * disp_step;
    if (x + 3 < cols) {
        if (elem == 2 && ((reinterpret_cast<uintptr_t>(drow) + 2 * size_t(x)) & 7) == 0) {
            const uint32_t lo = (uint32_t(d[0]) & 0xFFFFu) | (uint32_t(d[1]) << 16), hi = (uint32_t(d[2]) & 0xFFFFu) | (uint32_t(d[3]) << 16);
            *reinterpret_cast<uint2*>(drow + 2 * size_t(x)) = make_uint2(lo, hi);
            return;
        }
        if (elem == 1 && ((reinterpret_cast<uintptr_t>(drow) + size_t(x)) & 3) == 0) {
            *reinterpret_cast<uint32_t*>(drow + x) = (uint32_t(d[0]) & 0xFFu) | ((uint32_t(d[1]) & 0xFFu) << 8) | ((uint32_t(d[2]) & 0xFFu) << 16) | (uint32_t(d[3]) << 24);
            return;
        }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (x + k >= cols) break;
        if (elem == 1) reinterpret_cast<int8_t*>(drow)[x + k] = int8_t(uint8_t(uint32_t(d[k]) & 0xFFu));
        else if (elem == 2) reinterpret_cast<int16_t*>(drow)[x + k] = int16_t(d[k]);
        else reinterpret_cast<int32_t*>(drow)[x + k] = d[k];
    }
}

// A thread takes MG_SUB blocks of 4 pixels, 512 columns apart: the 16-byte loads of the partial keys of all of them (first
// MG_PRE groups) are issued before the first is used - 64 bytes per thread in flight; one block per thread left the launch
// at 3.6 TB/s (188 us for four 4K pairs) - then vector stores of the disparities.
#ifndef SB_MG_SUB
#define SB_MG_SUB 2
#endif
constexpr int MG_SUB = SB_MG_SUB, MG_PRE = 2, MG_THREADS = 128, MG_COLS = MG_THREADS * 4 * MG_SUB;
__device__ __forceinline__ void merge_ssd_block(const FastKernelParams& P, const FastJob& job, const int yy, const int x4, const int4 (&pre)[MG_PRE]) {
    const FastGeom& g = P.g;
    const int32_t* __restrict__ PART = job.PART;
    const uint8_t* __restrict__ A = job.A; const size_t a_step = job.a_step;
    void* best_out = job.best; const size_t best_step = job.best_step;
    // A key is BIAS + 128 * cost + position, ordered by (cost, position) among the < 128 positions of ONE group only.  Relative
    // to its group's first position (nk = key - qlo) every group's keys read 128 * cost' + rel with rel < 128, so the winner
    // over the groups - smallest cost, then the lowest group (its candidates come first), then the smallest position - is
    // one shift and one compare per further group, and position / cost are decoded once per pixel instead of once per group
    // (the launch was as much bound by these integer ops as by its bytes: 160 us for four 4K pairs whose bytes need 100).
    uint32_t bnk[4], bq[4];          // best relative key and its group's first position
    const uint32_t thresh = g.opf ? KEY_INVALID : key_invalid_threshold(g.R), bias = key_bias(g.R);
    const uint32_t q0 = uint32_t(x4 + job.dlo0 + job.eoff);                  // position of group 0's first candidate for pixel x4
    auto take = [&](const int grp, const int4 kv) {
        const uint32_t keys[4] = {uint32_t(kv.x), uint32_t(kv.y), uint32_t(kv.z), uint32_t(kv.w)};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t qlo = q0 + uint32_t(k + g.dg * grp);
            const uint32_t nk = keys[k] >= thresh ? KEY_INVALID : keys[k] - qlo;     // no legal candidate in this group: loses
            if (grp == 0 || (nk >> FKEY_BITS) < (bnk[k] >> FKEY_BITS)) { bnk[k] = nk; bq[k] = qlo; }
        }
    };
#pragma unroll
    for (int i = 0; i < MG_PRE; ++i) if (i < g.G) take(i, pre[i]);
    for (int grp = MG_PRE; grp < g.G; ++grp) take(grp, *reinterpret_cast<const int4*>(PART + (size_t(grp) * g.nrows + yy) * g.wpart + x4));
    int bestc[4], bestd[4];
    bool found[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        found[k] = bnk[k] != KEY_INVALID;
        const uint32_t rel = bnk[k] & uint32_t(FGROUP - 1);
        // packed: ER - 2C (exact: multiple of 128);  float operands: the SSD itself (key = 128 * SSD + position)
        bestc[k] = found[k] ? (g.opf ? int((bnk[k] - rel) >> FKEY_BITS) : (int(bnk[k] - rel - bias) >> FKEY_BITS)) : INT_MAX;
        bestd[k] = found[k] ? int(bq[k] + rel) - job.eoff - (x4 + k) : 0;
    }
    store_disp4(job.disp, job.disp_step, job.elem, yy, x4, g.cols, bestd);
    // EL(x), the window energy of the reference image (replicate padding), is only needed to report
    // the cost and to honour the 99999999 threshold (DisparitySSD.cpp:37); with (2R+1)^2 * 255^2 < 99999999
    // (R <= 19) the threshold can never bind, so the energy is computed only when the caller asked for costs.
    if (best_out) {
        const int y = g.rb + yy;
        for (int k = 0; k < 4 && x4 + k < g.cols; ++k) {
            const int x = x4 + k;
            int cost = 99999999;
            if (found[k] && g.opf) {
                cost = bestc[k] < 99999999 ? bestc[k] : 99999999;
            } else if (found[k]) {
                int el = 0;
                for (int wy = -g.R; wy <= g.R; ++wy) {
                    const uint8_t* row = A + size_t(clampi(clampi(y + wy, 0, g.rows - 1), g.ar0, g.ar1 - 1)) * a_step;
                    for (int wx = -g.R; wx <= g.R; ++wx) { const int v = row[clampi(x + wx, 0, g.cols - 1)]; el += v * v; }
                }
                cost = bestc[k] + el;
            }
            reinterpret_cast<int32_t*>(reinterpret_cast<char*>(best_out) + size_t(yy) * best_step)[x] = cost;
        }
    }
}

__global__ void __launch_bounds__(MG_THREADS) fast_merge_ssd_kernel(const __grid_constant__ FastKernelParams P) {
    const FastGeom& g = P.g;
    const FastJob& job = P.job[blockIdx.z];
    const int yy = blockIdx.y;
    int4 pre[MG_SUB][MG_PRE];
#pragma unroll
    for (int sblk = 0; sblk < MG_SUB; ++sblk) {
        const int x4 = blockIdx.x * MG_COLS + sblk * (MG_THREADS * 4) + threadIdx.x * 4;
#pragma unroll
        for (int i = 0; i < MG_PRE; ++i)
            if (x4 < g.cols && i < g.G) pre[sblk][i] = *reinterpret_cast<const int4*>(job.PART + (size_t(i) * g.nrows + yy) * g.wpart + x4);
    }
#pragma unroll
    for (int sblk = 0; sblk < MG_SUB; ++sblk) {
        const int x4 = blockIdx.x * MG_COLS + sblk * (MG_THREADS * 4) + threadIdx.x * 4;
        if (x4 < g.cols) merge_ssd_block(P, job, yy, x4, pre[sblk]);
    }
}

// NCC: winning key per group -> first maximum over the groups -> disparity with the reference's
// alignment rule (DisparityNCorr.cpp:67) and, on request, the winning score recomputed exactly with
// TM_CCORR_NORMED's arithmetic (float32 numerator, double energies; see ncorr_exact_kernel).
__device__ __forceinline__ void merge_ncc_block(const FastKernelParams& P, const FastJob& job, const int yy, const int x4, const int4 (&pre)[MG_PRE]) {
    const FastGeom& g = P.g;
    const int32_t* __restrict__ PART = job.PART;
    const uint8_t* __restrict__ A = job.A; const size_t a_step = job.a_step;
    const uint8_t* __restrict__ B = job.B; const size_t b_step = job.b_step;
    void* best_out = job.best; const size_t best_step = job.best_step;
    // first maximum over the groups: a later group wins only with strictly more value bits (same pixel -> same key scale ->
    // comparable); the winner's candidate index is decoded once per pixel
    uint32_t bkey[4]; int bgrp[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) { bkey[k] = NCC_KEY_NONE; bgrp[k] = 0; }
    auto take = [&](const int grp, const int4 kv) {
        const uint32_t keys[4] = {uint32_t(kv.x), uint32_t(kv.y), uint32_t(kv.z), uint32_t(kv.w)};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t key = keys[k];
            // (NCC_KEY_NONE = 0 never replaces anything; a legal key replaces NONE even when its value bits are 0)
            if (key != NCC_KEY_NONE && (bkey[k] == NCC_KEY_NONE || (key >> NCC_KEY_SHIFT) > (bkey[k] >> NCC_KEY_SHIFT))) { bkey[k] = key; bgrp[k] = grp; }
        }
    };
#pragma unroll
    for (int i = 0; i < MG_PRE; ++i) if (i < g.G) take(i, pre[i]);
    for (int grp = MG_PRE; grp < g.G; ++grp) take(grp, *reinterpret_cast<const int4*>(PART + (size_t(grp) * g.nrows + yy) * g.wpart + x4));
    int bestd[4];
#pragma unroll
    for (int k = 0; k < 4; ++k)
        bestd[k] = bkey[k] == NCC_KEY_NONE ? 0 : job.dlo0 + g.dg * bgrp[k] + (FGROUP - 1 - int((bkey[k] >> 2) & (FGROUP - 1)));
    const bool right_aligned = (job.dmin <= 0 && job.dmax <= 0);
    int disp[4], centre[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int x = x4 + k;
        const int startc = max(0, x + job.dmin), endc = min(g.cols - 1, x + job.dmax);
        centre[k] = x + bestd[k];                                       // winning window centre (unpadded column)
        // Illegal positions (centre outside the image) carry the score-0 key of their position (RS = 0).  One of
        // them winning means every legal candidate scored exactly 0 (a non-zero C*rs never quantises to 0:
        // rs >= 1/sqrt(Emax) > magic * 2^-24), and the first maximum of an all-zero result row is its first entry.
        if (centre[k] < startc || centre[k] > endc) centre[k] = startc;
        disp[k] = (centre[k] - startc) - (right_aligned ? endc - startc : 0);
    }
    store_disp4(job.disp, job.disp_step, job.elem, yy, x4, g.cols, disp);
    if (best_out) {
        const int y = g.rb + yy;
        for (int k = 0; k < 4 && x4 + k < g.cols; ++k) {
            const int x = x4 + k;
            double num, wnd, eld;
            if (g.opf) {       // float images: double sums (products of floats are exact in double), float32 numerator
                double acc = 0, e1 = 0, e2 = 0;
                for (int wy = -g.R; wy <= g.R; ++wy) {
                    const int wr = clampi(clampi(y + wy, 0, g.rows - 1), g.ar0, g.ar1 - 1);
                    const float* arow = reinterpret_cast<const float*>(A + size_t(wr) * a_step);
                    const float* brow = reinterpret_cast<const float*>(B + size_t(wr) * b_step);
                    for (int wx = -g.R; wx <= g.R; ++wx) {
                        const double l = arow[clampi(x + wx, 0, g.cols - 1)], r = brow[clampi(centre[k] + wx, 0, g.cols - 1)];
                        acc += l * r; e1 += l * l; e2 += r * r;
                    }
                }
                num = double(float(acc)); wnd = e2; eld = e1;
            } else {
                int c = 0, el = 0, er = 0;
                for (int wy = -g.R; wy <= g.R; ++wy) {
                    const int wr = clampi(clampi(y + wy, 0, g.rows - 1), g.ar0, g.ar1 - 1);
                    const uint8_t* arow = A + size_t(wr) * a_step;
                    const uint8_t* brow = B + size_t(wr) * b_step;
                    for (int wx = -g.R; wx <= g.R; ++wx) {
                        const int l = arow[clampi(x + wx, 0, g.cols - 1)], r = brow[clampi(centre[k] + wx, 0, g.cols - 1)];
                        c += l * r; el += l * l; er += r * r;
                    }
                }
                num = double(float(c)); wnd = double(er); eld = double(el);
            }
            const double lim = fmin(0.5, 10 * double(FLT_EPSILON) * wnd);
            const double t = (wnd <= lim) ? 0 : sqrt(wnd) * sqrt(eld);
            if (fabs(num) < t) num /= t;
            else if (fabs(num) < t * 1.125) num = num > 0 ? 1 : -1;
            else num = 0;
            reinterpret_cast<float*>(reinterpret_cast<char*>(best_out) + size_t(yy) * best_step)[x] = float(num);
        }
    }
}

__global__ void __launch_bounds__(MG_THREADS) fast_merge_ncc_kernel(const __grid_constant__ FastKernelParams P) {
    const FastGeom& g = P.g;
    const FastJob& job = P.job[blockIdx.z];
    const int yy = blockIdx.y;
    int4 pre[MG_SUB][MG_PRE];
#pragma unroll
    for (int sblk = 0; sblk < MG_SUB; ++sblk) {
        const int x4 = blockIdx.x * MG_COLS + sblk * (MG_THREADS * 4) + threadIdx.x * 4;
#pragma unroll
        for (int i = 0; i < MG_PRE; ++i)
            if (x4 < g.cols && i < g.G) pre[sblk][i] = *reinterpret_cast<const int4*>(job.PART + (size_t(i) * g.nrows + yy) * g.wpart + x4);
    }
#pragma unroll
    for (int sblk = 0; sblk < MG_SUB; ++sblk) {
        const int x4 = blockIdx.x * MG_COLS + sblk * (MG_THREADS * 4) + threadIdx.x * 4;
        if (x4 < g.cols) merge_ncc_block(P, job, yy, x4, pre[sblk]);
    }
}

// ---------------------------------------------------------------------------------------------------
// Fused pair launches: the partner's candidates centred in the right padding, evaluated directly
// ---------------------------------------------------------------------------------------------------
// The fused hot kernel gives the right-referenced pixel x' every candidate centred on an image column.  The reference
// also searches the right padding (DisparitySSD.cpp:39-40: the clamp is colsP-1): centres pos in
// [cols, min(x' + range, cols-1+R)], whose windows run past the end of the padded row and alias the next row through
// the flat index (bext).  Normally the strips simply extend R columns (LP rows built with bext); when those R columns
// would cost a whole extra tile (narrow images: 1280 columns = 4 tiles of 320, +4 columns = a fifth), the <= R
// candidates of the last `range` pixels of every row are evaluated here instead and merged into the same partial-key
// map (RED.MIN) with the key the hot kernel would have produced: BIAS + 128*(ER - 2C) + pos.
//
// A CTA of 256 threads takes 2^k pixels (32..256, the smallest that covers `range`) of 256 / 2^k consecutive rows.  The 3R
// extended target columns of a window row - R image columns, R times the replicated last column, R times the aliased first
// column of the next padded row (zero past the last padded row) - depend on the row alone: the CTA packs every candidate's
// W-byte window of every window row it touches into words in shared memory once (and sums their squares: ER).  The
// reference-image patch under the CTA's windows is staged in shared memory by the same round of loads; a thread packs its
// pixel's window rows from there and takes 4 IDP.4A per (window row, candidate) against broadcast LDS.128 reads.
// (The first version read every byte of both windows from global memory per candidate: 45 us for 511 rows x 95 pixels
// against 50 us for the whole cost volume.)
constexpr int FB_THREADS = 256;
constexpr int FB_MAXROWS = FB_THREADS / 32;
static inline int fused_border_ppr_log2(int range) { int k = 5; while (k < 8 && (1 << k) < range) ++k; return k; }
__host__ __device__ constexpr int fb_patch_bytes(int R) {          // largest patch over k = 5..8: (W + 256/2^k - 1) rows x (2^k + 2R + 1)
    int best = 0;
    for (int k = 5; k <= 8; ++k) { const int v = (2 * R + 1 + (256 >> k) - 1) * ((1 << k) + 2 * R + 1); if (v > best) best = v; }
    return best;
}
template <int R>
__device__ __forceinline__ void fused_border_stage(const FastKernelParams& P, const uint3 bid, const int ppr_log2) {
    constexpr int W = 2 * R + 1, NB = 3 * R > 0 ? 3 * R : 1, NCAND = R > 0 ? R : 1, NWORD = (W + 3) / 4;
    constexpr int TR = W + FB_MAXROWS - 1;                     // window rows a CTA touches at most
    __shared__ uint8_t bcol[TR][NB + 1];
    __shared__ __align__(16) uint32_t bw[TR][NCAND][4];
    __shared__ int er[FB_MAXROWS][NCAND];
    __shared__ uint8_t apatch[fb_patch_bytes(R)];
    const FastGeom& g = P.g;
    const FastJob& job = P.job[g.npairs + bid.z];               // the right-referenced direction: A = right image, B = left image
    const int range = job.dmax;
    const int ppr = 1 << ppr_log2, rpc = FB_THREADS >> ppr_log2; // pixels per row and output rows of this CTA
    const int tr = W + rpc - 1;                                  // window rows it touches
    const int yy0 = int(bid.y) * rpc, y0 = g.rb + yy0;
    const int xr = g.cols - 1 - (int(bid.x) << ppr_log2);        // its rightmost pixel; pixel il is x' = xr - il
    const int pw = ppr + 2 * R, pitch = pw + 1, xlo = xr - ppr + 1 - R;   // patch column pc <-> image column xlo + pc
    const uint8_t* __restrict__ A = job.A; const uint8_t* __restrict__ B = job.B;
    for (int t = threadIdx.x; t < tr * NB; t += FB_THREADS) {
        const int wr = t / NB, k = t % NB;
        const int ra = clampi(clampi(y0 + wr - R, 0, g.rows - 1), g.ar0, g.ar1 - 1);
        const uint8_t* brow = B + size_t(ra) * job.b_step;
        bcol[wr][k] = k < R ? brow[max(g.cols - R + k, 0)] : k < 2 * R ? brow[g.cols - 1]
                    : uint8_t(bext(B, job.b_step, g.rows, g.cols, R, y0 + wr - R, g.cols + 4 * R, g.ar0, g.ar1));   // padded column cols + 3R >= Wp
    }
    for (int pr = threadIdx.x >> 5; pr < tr; pr += FB_THREADS / 32) {
        const int ra = clampi(clampi(y0 + pr - R, 0, g.rows - 1), g.ar0, g.ar1 - 1);
        const uint8_t* arow = A + size_t(ra) * job.a_step;
        for (int pc = threadIdx.x & 31; pc < pw; pc += 32) apatch[pr * pitch + pc] = arow[clampi(xlo + pc, 0, g.cols - 1)];
    }
    __syncthreads();
    for (int t = threadIdx.x; t < tr * NCAND * 4; t += FB_THREADS) {
        const int wr = t / (NCAND * 4), c = (t / 4) % NCAND, q = t % 4;
        uint32_t word = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) { const int i = 4 * q + k; if (i < W) word |= uint32_t(bcol[wr][c + i]) << (8 * k); }
        bw[wr][c][q] = word;
    }
    __syncthreads();
    if (threadIdx.x < rpc * NCAND) {
        const int r = threadIdx.x / NCAND, c = threadIdx.x % NCAND;
        uint32_t e = 0;
        for (int wr = 0; wr < W; ++wr)
#pragma unroll
            for (int q = 0; q < NWORD; ++q) e = __dp4a(bw[r + wr][c][q], bw[r + wr][c][q], e);
        er[r][c] = int(e);
    }
    __syncthreads();
    const int r = threadIdx.x >> ppr_log2, il = threadIdx.x & (ppr - 1);
    const int idx = (int(bid.x) << ppr_log2) + il;
    const int xp = xr - il;                                     // the pixel x'
    const int yy = yy0 + r;
    if (yy >= g.nrows || idx >= range || xp < 0) return;
    const int ncand = min(range - idx, R);                      // centres cols .. min(x' + range, cols-1+R)
    uint32_t acc[NCAND];
#pragma unroll
    for (int c = 0; c < NCAND; ++c) acc[c] = 0;
    const uint8_t* ap = apatch + r * pitch + (ppr - 1 - il);    // window element i of window row wr: ap[wr * pitch + i]
#pragma unroll 3
    for (int wr = 0; wr < W; ++wr, ap += pitch) {
        uint32_t aw[NWORD];
#pragma unroll
        for (int q = 0; q < NWORD; ++q) {
            uint32_t word = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) { const int i = 4 * q + k; if (i < W) word |= uint32_t(ap[i]) << (8 * k); }
            aw[q] = word;
        }
#pragma unroll
        for (int c = 0; c < NCAND; ++c) {
            const uint4 b4 = *reinterpret_cast<const uint4*>(bw[r + wr][c]);
            const uint32_t bq[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int q = 0; q < NWORD; ++q) acc[c] = __dp4a(aw[q], bq[q], acc[c]);
        }
    }
    uint32_t* __restrict__ PART2 = reinterpret_cast<uint32_t*>(job.PART);
#pragma unroll
    for (int c = 0; c < NCAND; ++c) {
        if (c >= ncand) break;
        const int pos = g.cols + c;
        const uint32_t key = key_bias(R) + ((uint32_t(er[r][c]) - 2u * acc[c]) << FKEY_BITS) + uint32_t(pos);
        atomicMin(PART2 + (size_t((pos - xp) / g.dg) * g.nrows + yy) * g.wpart + xp, key);
    }
}

// u8 operands.  Stage 0: target rows, energies / position keys (prep_tgt); 1: the partner's candidates in the right padding
// (fused_border, fused pairs only - after a memset of the partners' partial keys); 2: reference rows (prep_lp); 3: fused pairs
// without a border stage: the partners' partial keys start from "no candidate" (prep_fill).
template <int R>
__global__ void __launch_bounds__(256, R >= 6 ? 3 : 4) prep_u8_kernel(const __grid_constant__ FastKernelParams P, const __grid_constant__ PrepStages s) {
    static_assert(PT_THREADS == 256 && FB_THREADS == 256, "the stages share one CTA shape");
    __shared__ __align__(16) int vs[2][4][PT_VSTRIDE];
    if (blockIdx.x < s.first[1]) prep_tgt_stage<R>(P, prep_stage_block(s, 0), s.rows_per_cta, vs);
    else if (blockIdx.x < s.first[2]) { if constexpr (R > 0) fused_border_stage<R>(P, prep_stage_block(s, 1), s.border_ppr_log2); }
    else if (blockIdx.x < s.first[3]) prep_lp_body(P, prep_stage_block(s, 2));
    else prep_fill_body(P, prep_stage_block(s, 3), s);
}
typedef void (*prep_u8_fn)(const FastKernelParams, const PrepStages);
static inline prep_u8_fn prep_u8_pick(int R) {
    switch (R) {
        case 0: return prep_u8_kernel<0>; case 1: return prep_u8_kernel<1>; case 2: return prep_u8_kernel<2>;
        case 3: return prep_u8_kernel<3>; case 4: return prep_u8_kernel<4>; case 5: return prep_u8_kernel<5>;
        case 6: return prep_u8_kernel<6>; case 7: return prep_u8_kernel<7>;
    }
    return nullptr;
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

// Problems that may share one launch sequence: same images shape, band, window, cost and candidate count.
static inline bool fast_batchable(const Problem& a, const Problem& b) {
    return a.cost == b.cost && a.rows == b.rows && a.cols == b.cols && a.R == b.R && a.row_begin == b.row_begin &&
           a.row_end == b.row_end && a.avail_begin == b.avail_begin && a.avail_end == b.avail_end &&
           (a.dmax - a.dmin) == (b.dmax - b.dmin) && a.ref.type == a.tgt.type && b.ref.type == b.tgt.type && a.ref.type == b.ref.type;
}

// Strips per warp: 2 for searches of at most 64 candidates (a warp then covers 2 x 24 pixels x 64 disparities
// instead of leaving half its lanes idle).  STEREO_FAST_HS=1 forces the single-strip kernels (debug knob).
static inline int fast_pick_hs(int D) {
    static const int forced = [] { const char* e = getenv("STEREO_FAST_HS"); return e ? atoi(e) : 0; }();
    if (forced == 1) return 1;
    return D <= 64 ? 2 : 1;
}

// A left-referenced and a right-referenced problem of the SAME image pair whose maps can come out of one cost
// volume (fused pair launch): SSD, any supported window, mirrored ranges [-r, 0] / [0, r].  The walked direction's
// disparity groups are aligned to the TOP of its range (FastJob::dlo0), so the partner's groups are those groups
// reversed whatever the candidate count; the candidates below -r of the lowest group are masked in both maps.
static inline bool fast_pair_fusable(const Problem& a, const Problem& b) {
    static const bool off = [] { const char* e = getenv("STEREO_FUSE_PAIRS"); return e && atoi(e) == 0; }();
    if (off || !fast_batchable(a, b)) return false;
    const int range = -a.dmin;
    static const bool ncc_off = [] { const char* e = getenv("STEREO_FUSE_NCC"); return e && atoi(e) == 0; }();
    const bool cost_ok = a.cost == STEREO_COST_SSD || (a.cost == STEREO_COST_NCORR && a.ref.type == PixType::U8 && !ncc_off);
    return cost_ok && a.R <= FMAXR && a.dmax == 0 && range > 0 && b.dmin == 0 && b.dmax == range &&
           a.ref.ptr == b.tgt.ptr && a.tgt.ptr == b.ref.ptr && a.ref.step == b.tgt.step && a.tgt.step == b.ref.step;
}

// Jobs i and i + n/2 of a launch are the two directions of one image pair (any cost): left-referenced problems first.
static inline bool fast_launch_is_pairs(const Problem* ps, int n) {
    if (n < 2 || (n & 1)) return false;
    const int np = n / 2;
    for (int k = 0; k < np; ++k) {
        const Problem& a = ps[k]; const Problem& b = ps[np + k];
        if (!(a.ref.ptr == b.tgt.ptr && a.tgt.ptr == b.ref.ptr && a.ref.step == b.tgt.step && a.tgt.step == b.ref.step &&
              a.dmax == 0 && b.dmin == 0 && b.dmax == -a.dmin)) return false;
    }
    return true;
}

static inline size_t fast_stage_bytes(const FastGeom& g) {
    if (g.opf) return (size_t(2 * FRPS) * g.lpw + size_t(2 * FRPS) * g.rqw + size_t(FRPS) * g.e2w + size_t(FRPS) * g.elw) * 4;
    return (size_t(FRPS) * g.lpw + size_t(FRPS / 2) * g.rqw + size_t(FRPS) * g.e2w + size_t(FRPS) * g.elw) * 4;
}

// Geometry of a launch over `n` batchable problems; fills the per-job offsets of `jobs` (pointers are the
// caller's business).
// `fused_pairs` > 0: a fused pair launch over n = 2*fused_pairs problems, left-referenced directions first; the hot
// kernel walks the first half only and produces the partners' partial keys on the way.
static inline void fast_geometry_nw(const stereo_ctx* ctx, const Problem* ps, int n, FastGeom& g, FastJob* jobs, int fused_pairs, int nw) {
    const Problem& p = ps[0];
    g.rows = p.rows; g.cols = p.cols; g.R = p.R; g.cost = p.cost;
    g.D = p.dmax - p.dmin + 1;
    g.rb = p.row_begin; g.re = p.row_end; g.nrows = g.re - g.rb;
    g.ar0 = p.avail_begin; g.ar1 = p.avail_end;
    g.njobs = n;
    g.npairs = fused_pairs;
    g.elw = 0;
    g.border = 0;
    g.nw = nw;
    g.opf = p.ref.type == PixType::F32 ? 1 : 0;
    g.hs = g.opf ? 1 : fast_pick_hs(g.D);
    const int w = 2 * p.R + 1;
    for (;;) {
        g.K = g.opf ? fast_kf(p.R) : ((fused_pairs > 0 && p.cost == STEREO_COST_NCORR) ? FK_FUSED_NCC : fast_k(p.R, fused_pairs > 0, g.hs));
        g.dg = FGROUP / g.hs;
        g.G = (g.D + g.dg - 1) / g.dg;
        g.gc = (g.hs == 1 && g.G % 2 == 0) ? 2 : 1;
        g.spc = g.nw * g.hs / g.gc;
        const int tile_px = g.spc * g.K;
        g.lpw = round_up(tile_px + 2 * p.R, 4);
        g.rqw = round_up(tile_px + 2 * p.R + g.dg * g.gc + FM, 4);
        g.e2w = round_up(tile_px + g.dg * g.gc + FM, 4);
        g.elw = fused_pairs > 0 ? tile_px : 0;
        g.nst = int((FSMEM_BUDGET - 2 * FNST_MAX * 8 - 16 - FWARPS * 32 - (fused_pairs > 0 ? FWARPS * 256 : 0)) / fast_stage_bytes(g));
        if (g.nst > FNST_MAX) g.nst = FNST_MAX;
        if (g.nst >= 4 || g.hs == 1) break;
        g.hs = 1;                                  // tile rows too wide for a useful pipeline: single-strip kernels
    }
    const int tile_px = g.spc * g.K;
    // fused: the strips also cover the R columns of right padding the partner direction may centre a window on -
    // unless that costs a whole tile of a narrow image (>= 1/8 more work); then fused_border_kernel takes them
    g.nstrips = (p.cols + g.K - 1) / g.K;
    if (fused_pairs > 0 && p.R > 0 && p.cost == STEREO_COST_SSD) {      // (NCC candidates never centre in the padding)
        const int ext_strips = (p.cols + p.R + g.K - 1) / g.K;
        const int t0 = (g.nstrips + g.spc - 1) / g.spc, t1 = (ext_strips + g.spc - 1) / g.spc;
        if (t1 > t0 && t0 <= 8 && !g.opf) g.border = 1;      // (the border kernel is 8-bit only)
        else g.nstrips = ext_strips;
    }
    g.tilesX = (g.nstrips + g.spc - 1) / g.spc;
    g.gblocks = g.G / g.gc;
    g.base_y = floor_div(g.rb - w, FRPS) * FRPS;
    g.J = round_up(g.re - g.base_y, FRPS);
    g.wpart = g.tilesX * tile_px;
    g.lp_pitch = round_up((g.tilesX - 1) * tile_px + g.lpw, 64);
    const int last_p0 = (g.tilesX - 1) * tile_px;
    const int gmax = g.dg * (g.gblocks - 1) * g.gc;
    g.rq_pitch = 0; g.e2_pitch = 0;
    for (int i = 0; i < n; ++i) {
        FastJob& jb = jobs[i];
        const Problem& q = ps[i];
        jb.dmin = q.dmin; jb.dmax = q.dmax;
        // group layout: from dmin upwards; the walked directions of a fused launch from dmax downwards (whole groups)
        jb.dlo0 = (fused_pairs > 0 && i < fused_pairs) ? q.dmax - (g.G * g.dg - 1) : q.dmin;
        if (q.cost == STEREO_COST_SSD) { jb.cmin = -q.R; jb.cmax = q.cols - 1 + q.R; }
        else { jb.cmin = 0; jb.cmax = q.cols - 1; }
        // fused partner: its energy rows feed the diagonal minima; candidates left of the image do not exist for it
        // (and, when fused_border_kernel takes the right padding, none beyond the last image column either)
        if (fused_pairs > 0 && i >= fused_pairs) { jb.cmin = 0; if (g.border) jb.cmax = q.cols - 1; }
        // RQ column q = e + qoff with e = x0 + dl + R + (c+m); first index must be >= 0 and 4-aligned
        int qo = -(jb.dlo0 + q.R); if (qo < 0) qo = 0;
        while (((jb.dlo0 + q.R + qo) & 3) != 0) ++qo;
        jb.qoff = qo;
        int eo = -jb.dlo0; if (eo < 0) eo = 0;
        while (((jb.dlo0 + eo) & 3) != 0) ++eo;
        jb.eoff = eo;
        const int rqp = round_up(last_p0 + jb.dlo0 + gmax + q.R + jb.qoff + g.rqw, 64);
        const int e2p = round_up(last_p0 + jb.dlo0 + gmax + jb.eoff + g.e2w, 64);
        if (rqp > g.rq_pitch) g.rq_pitch = rqp;
        if (e2p > g.e2_pitch) g.e2_pitch = e2p;
    }
    // grid: one CTA per SM, two ways to cut the (tile, row) space:
    //   linear   - split evenly into sm_count shares; a share that crosses a tile boundary pays the (2R+1)-row warm-up twice
    //   segments - (fewer tiles than SMs) every tile is cut into the same number of row segments, one CTA each: no share
    //              straddles, but ntiles * spt CTAs may leave SMs idle
    // whichever has the smaller modelled critical path (a warm-up row costs about half a regular row).
    const long long ntiles = (long long)(fused_pairs > 0 ? fused_pairs : n) * g.tilesX * g.gblocks;
    const long long lin_total = ntiles * g.nrows;
    const long long lin_L = (lin_total + ctx->sm_count - 1) / ctx->sm_count;
    const double t_lin = double(lin_L) + w * 0.5 * (lin_L < g.nrows ? 2 : 1 + (lin_L + g.nrows - 1) / g.nrows);
    int spt = 0, seg_L = 0;
    double t_seg = 1e30;
    if (ntiles < ctx->sm_count) {
        spt = int(ctx->sm_count / ntiles);                     // segments per tile
        seg_L = (g.nrows + spt - 1) / spt;
        if (seg_L < FRPS) seg_L = FRPS;                        // at least one pipeline stage of rows per CTA
        spt = (g.nrows + seg_L - 1) / seg_L;
        t_seg = double(seg_L) + w * 0.5;
    }
    g.sched = 0; g.ntiles = int(ntiles); g.nbands = 1;
    // Many tiles (large images): row-band-major items, strided over the CTAs (FastGeom::sched) when some band count up to 16
    // fills the waves to within 2 % of the linear split - the price of the L2 locality is one more warm-up per band.
    static const int sched_env = [] { const char* e = getenv("STEREO_FAST_SCHED"); return e ? atoi(e) : -1; }();
    int best_nb = 0; double t_band = 1e30;
    if (ntiles >= ctx->sm_count && sched_env != 0) {
        for (int nb = 1; nb <= 16; ++nb) {
            const int Lb = (g.nrows + nb - 1) / nb;
            if (Lb < 8 * w) break;
            const long long items = ntiles * ((g.nrows + Lb - 1) / Lb);
            const long long waves = (items + ctx->sm_count - 1) / ctx->sm_count;
            const double t = double(waves) * (Lb + w * 0.5);
            if (t < t_band) { t_band = t; best_nb = nb; }
        }
    }
    if (best_nb > 0 && (t_band <= 1.02 * t_lin || sched_env == 1)) {
        const int Lb = (g.nrows + best_nb - 1) / best_nb;
        g.sched = 1;
        g.L = Lb;
        g.nbands = (g.nrows + Lb - 1) / Lb;
        g.nrl = g.nrows;
        g.total = ntiles * g.nrows;
        const long long items = ntiles * g.nbands;
        g.ctas = int(items < ctx->sm_count ? items : ctx->sm_count);
    } else if (t_seg < t_lin) {
        g.L = seg_L;
        g.nrl = spt * seg_L;
        g.total = ntiles * g.nrl;
        g.ctas = int(ntiles * spt);
    } else {
        g.nrl = g.nrows;
        g.total = lin_total;
        g.L = lin_L;
        g.ctas = int((g.total + g.L - 1) / g.L);
    }
}

static inline void fast_geometry(const stereo_ctx* ctx, const Problem* ps, int n, FastGeom& g, FastJob* jobs, int fused_pairs = 0) {
    fast_geometry_nw(ctx, ps, n, g, jobs, fused_pairs, FWARPS);
}

// Fused launches whose blocks are all MODE 1 (whole disparity groups, masks free through the keys: R <= 5) take the
// kernel compiled without the other row flavours.
static inline bool fast_fused_all_mode1(const FastGeom& g) { return g.R <= FFREE_MASK_R && g.D % g.dg == 0; }

static inline size_t fast_smem_bytes(const FastGeom& g) {
    return size_t(g.nst) * fast_stage_bytes(g) + 2 * FNST_MAX * 8 + 16 + FWARPS * 32 + (g.elw ? FWARPS * 256 : 0);      // stages, barriers, producer caches, tails
}

static inline int fast_vpitch(const FastGeom& g) { return round_up(g.e2_pitch + 2 * g.R + PE_COLS, 64); }

// Scratch of ONE job of a launch over problems shaped like `p` (the larger of the plain and the fused-pair geometry:
// fused launches add R columns of strips and may use another strip width).
static inline size_t fast_scratch_bytes(stereo_ctx* ctx, const Problem& p) {
    size_t best = 0;
    for (int fused = 0; fused <= 1; ++fused) {
        if (fused && p.cost != STEREO_COST_SSD && p.ref.type != PixType::U8) break;
      {
        FastGeom g; FastJob jb{};
        fast_geometry(ctx, &p, 1, g, &jb, fused);
        // pitches may grow by one alignment step when jobs with other range signs join the launch
        g.rq_pitch += 64; g.e2_pitch += 64;
        size_t b = 0;
        auto add = [&](size_t bytes) { b += ((bytes + 255) & ~size_t(255)); };
        if (g.opf) { add(size_t(g.J + 2 * g.R + 1) * g.lp_pitch * 4); add(size_t(g.J + 2 * g.R + 1) * g.rq_pitch * 4); }
        else { add(size_t(g.J) * g.lp_pitch * 4); add(size_t(g.J / 2) * g.rq_pitch * 4); }
        add(size_t(g.J) * g.e2_pitch * 4);
        add((size_t(g.G) * g.nrows * g.wpart + 2 * fused_part_pad_words(g.G)) * 4);
        add(size_t(g.nrows) * fast_vpitch(g) * 4);
        if (p.cost == STEREO_COST_NCORR) { add(size_t(g.J) * g.e2_pitch * 4); add(size_t(g.tilesX) * g.spc * g.nrows * 4); }
        if (b > best) best = b;
      }
    }
    return best + 4096;
}

static inline int fast_ctx_init(stereo_ctx*) {
    for (int cost = 0; cost <= 1; ++cost)
        for (int hs = 1; hs <= 2; ++hs)
            for (int R = 0; R <= FMAXR; ++R) {
                fast_kernel_fn fn = fast_pick(cost, R, hs);
                if (!fn) { set_error("hot kernel (cost %d, R %d, hs %d) missing from the build", cost, R, hs); return STEREO_ERR_UNSUPPORTED; }
                cudaError_t e = cudaFuncSetAttribute(reinterpret_cast<const void*>(fn),
                                                     cudaFuncAttributeMaxDynamicSharedMemorySize, FSMEM_BUDGET);
                if (e != cudaSuccess) { set_error("cudaFuncSetAttribute(smem): %s", cudaGetErrorString(e)); return STEREO_ERR_CUDA; }
            }
    for (int R = 0; R <= FMAXR; ++R)
      for (int hs = 1; hs <= 2; ++hs)
        for (int gen = (R <= FFREE_MASK_R ? 0 : 1); gen <= 1; ++gen) {
        fast_kernel_fn fn = fast_pick_fused(R, hs, gen);
        if (!fn) { set_error("fused pair kernel (R %d, hs %d, gen %d) missing from the build", R, hs, gen); return STEREO_ERR_UNSUPPORTED; }
        cudaError_t e = cudaFuncSetAttribute(reinterpret_cast<const void*>(fn), cudaFuncAttributeMaxDynamicSharedMemorySize, FSMEM_BUDGET);
        if (e != cudaSuccess) { set_error("cudaFuncSetAttribute(smem): %s", cudaGetErrorString(e)); return STEREO_ERR_CUDA; }
    }
    for (int R = 0; R <= FMAXR; ++R)
        for (int hs = 1; hs <= 2; ++hs) {
            fast_kernel_fn fn = fast_pick_fused_ncc(R, hs);
            if (!fn) { set_error("fused NCC pair kernel (R %d, hs %d) missing from the build", R, hs); return STEREO_ERR_UNSUPPORTED; }
            cudaError_t e = cudaFuncSetAttribute(reinterpret_cast<const void*>(fn), cudaFuncAttributeMaxDynamicSharedMemorySize, FSMEM_BUDGET);
            if (e != cudaSuccess) { set_error("cudaFuncSetAttribute(smem): %s", cudaGetErrorString(e)); return STEREO_ERR_CUDA; }
        }
    for (int R = 0; R <= FMAXR; ++R)
        for (int kind = 0; kind < 3; ++kind) {
            fast_kernel_fn fn = fast_pick_opf(R, kind);
            if (!fn) { set_error("float-operand kernel (R %d, kind %d) missing from the build", R, kind); return STEREO_ERR_UNSUPPORTED; }
            cudaError_t e = cudaFuncSetAttribute(reinterpret_cast<const void*>(fn), cudaFuncAttributeMaxDynamicSharedMemorySize, FSMEM_BUDGET);
            if (e != cudaSuccess) { set_error("cudaFuncSetAttribute(smem): %s", cudaGetErrorString(e)); return STEREO_ERR_CUDA; }
        }
    return STEREO_OK;
}

// One launch sequence (prep x4-6, hot kernel, merge) over `n` <= FMAXJOBS batchable problems.
// `fused_pairs` > 0: ps holds fused_pairs left-referenced problems followed by their right-referenced partners
// (fast_pair_fusable pairwise); ONE hot launch over the left-referenced directions yields both maps of every pair.
static inline int run_fast_batch(stereo_ctx* ctx, const Problem* ps, int n, cudaStream_t st, int fused_pairs = 0) {
    if (n < 1 || n > FMAXJOBS || (fused_pairs && n != 2 * fused_pairs)) { set_error("bad job count (internal)"); return STEREO_ERR_INVALID_ARG; }
    FastKernelParams kp{};
    FastGeom& g = kp.g;
    fast_geometry(ctx, ps, n, g, kp.job, fused_pairs);
    const bool ncc = ps[0].cost == STEREO_COST_NCORR;
    const bool opf = g.opf != 0;
    const int vpitch = fast_vpitch(g);
    const int w = 2 * g.R + 1;
    const size_t lp_rows = opf ? size_t(g.J + w) : size_t(g.J), rq_rows = opf ? size_t(g.J + w) : size_t(g.J / 2);
    const size_t part_words = size_t(g.G) * g.nrows * g.wpart / 4;                       // (wpart is a multiple of 4)
    const bool fill_by_memset = (fused_pairs && g.border && g.R > 0) || part_words > 0xFFFFFFFFull - FILL_WORDS_PER_CTA;
    for (int i = 0; i < n; ++i) {
        FastJob& jb = kp.job[i];
        const Problem& p = ps[i];
        jb.A = static_cast<const uint8_t*>(p.ref.ptr); jb.a_step = p.ref.step;
        jb.B = static_cast<const uint8_t*>(p.tgt.ptr); jb.b_step = p.tgt.step;
        jb.disp = p.disp.ptr; jb.disp_step = p.disp.step; jb.elem = p.disp.elem;
        jb.best = p.best.ptr; jb.best_step = p.best.step;
        const bool partner = fused_pairs > 0 && i >= fused_pairs;      // only its energy rows and partial keys exist
        if (partner) {
            jb.E2 = static_cast<int32_t*>(ctx->arena.take(size_t(g.J) * g.e2_pitch * 4));
            jb.RS = reinterpret_cast<float*>(jb.E2);                  // NCC: the same rows hold 1/sqrt(energy)
            const size_t pad = fused_part_pad_words(g.G);             // (a multiple of 64 words: the map stays 256-byte aligned)
            jb.PART = static_cast<int32_t*>(ctx->arena.take((size_t(g.G) * g.nrows * g.wpart + 2 * pad) * 4));
            if (!jb.E2 || !jb.PART) { set_error("scratch arena too small (internal)"); return STEREO_ERR_ALLOC; }
            jb.PART += pad;
            // every strip merges into the partner's partial keys with RED.MIN (NCC: RED.MAX): start from "no candidate" -
            // a stage of the prep launch, or (the border stage merges into them inside that launch) a memset in front of it
            if (fill_by_memset) SB_CUDA(cudaMemsetAsync(jb.PART, ncc ? 0x00 : 0xFF, size_t(g.G) * g.nrows * g.wpart * 4, st));
            continue;
        }
        jb.LP = static_cast<int32_t*>(ctx->arena.take(lp_rows * g.lp_pitch * 4));
        jb.RQ = static_cast<uint32_t*>(ctx->arena.take(rq_rows * g.rq_pitch * 4));
        jb.E2 = static_cast<int32_t*>(ctx->arena.take(size_t(g.J) * g.e2_pitch * 4));
        jb.PART = static_cast<int32_t*>(ctx->arena.take(size_t(g.G) * g.nrows * g.wpart * 4));
        jb.V = static_cast<int32_t*>(ctx->arena.take(size_t(g.nrows) * vpitch * 4));
        if (ncc) {
            jb.RS = static_cast<float*>(ctx->arena.take(size_t(g.J) * g.e2_pitch * 4));
            jb.SC = static_cast<float*>(ctx->arena.take(size_t(g.tilesX) * g.spc * g.nrows * 4));
        }
        if (!jb.LP || !jb.RQ || !jb.E2 || !jb.PART || !jb.V || (ncc && (!jb.RS || !jb.SC))) {
            set_error("scratch arena too small (internal)"); return STEREO_ERR_ALLOC;
        }
    }
    const unsigned nz = unsigned(n), nwalk = fused_pairs ? unsigned(fused_pairs) : nz;
    static const bool legacy_prep = [] { const char* e = getenv("STEREO_PREP_LEGACY"); return e && atoi(e) != 0; }();
    fast_kernel_fn fn = nullptr;
    PrepStages stg{};
    if (opf) {
        // float operand rows straight from the images (padding, row-wrap aliasing), position / energy rows: one launch
        if (!ncc) prep_stage_set(stg, 0, unsigned(div_round_up(g.e2_pitch / 4, 256)), g.J, nz);
        else prep_stage_set(stg, 0, unsigned(div_round_up(g.e2_pitch, 256 - 2 * g.R)), unsigned(div_round_up(g.J, RSF_ROWS)), nz);
        stg.rows_per_cta = int(lp_rows);                 // (== rq_rows)
        prep_stage_set(stg, 1, unsigned(div_round_up(g.rq_pitch / 4, 256)), unsigned(div_round_up(rq_rows, size_t(FR_ROWS))), nwalk);
        prep_stage_set(stg, 2, unsigned(div_round_up(g.lp_pitch / 4, 256)), unsigned(div_round_up(lp_rows, size_t(FR_ROWS))), nwalk);
        if (fused_pairs && !fill_by_memset) {
            stg.fill_words = unsigned(part_words); stg.fill_value = ncc ? 0 : -1;
            prep_stage_set(stg, 3, unsigned(div_round_up((long long)part_words, FILL_WORDS_PER_CTA)), 1, unsigned(fused_pairs));
        }
        ctx->last_launches += prep_launch(prep_f32_kernel, kp, stg, st);
        if (ncc && fast_launch_is_pairs(ps, n)) {
            // both directions of every pair are in the launch: the reference-image energies are the partner's RS rows
            prep_scale_pair_kernel<<<dim3(div_round_up(g.tilesX * g.spc, 128), g.nrows, nz), 128, 0, st>>>(kp, n / 2);
            ctx->last_launches += 1;
        } else if (ncc) {
            prep_elf_kernel<<<dim3(div_round_up(g.cols, 128), g.nrows, nz), 128, 0, st>>>(kp, vpitch);
            prep_scale_f_kernel<<<dim3(div_round_up(g.tilesX * g.spc, 128), g.nrows, nz), 128, 0, st>>>(kp, vpitch);
            ctx->last_launches += 2;
        }
        fn = fast_pick_opf(g.R, ncc ? OPF_NCC : (fused_pairs ? OPF_SSD_FUSED : OPF_SSD));
    } else {
        // reference rows, target rows + energies and (fused pairs with a narrow last tile) the partner's candidates centred in
        // the right padding: one launch.  (STEREO_PREP_LEGACY=1, unfused launches only: the target stage as three passes.)
        const bool three_pass = legacy_prep && !fused_pairs;
        if (!three_pass) {
            int delta_max = 0;
            for (int i = 0; i < n; ++i) { const int d = kp.job[i].qoff - kp.job[i].eoff + g.R; if (d > delta_max) delta_max = d; }
            const int span = g.rq_pitch > g.e2_pitch + delta_max ? g.rq_pitch : g.e2_pitch + delta_max;
            const int tiles = int(div_round_up(span, pt_ts(g.R)));
            int rpc = PT_ROWS;                          // fewer rows per CTA while the grid would leave SMs idle
            while (rpc > 16 && (long long)tiles * div_round_up(g.J, rpc) * n < 6LL * ctx->sm_count) rpc /= 2;
            stg.rows_per_cta = rpc;
            prep_stage_set(stg, 0, unsigned(tiles), unsigned(div_round_up(g.J, rpc)), nz);
        }
        if (fused_pairs && g.border && g.R > 0) {
            const int range = -ps[0].dmin, k = fused_border_ppr_log2(range);
            stg.border_ppr_log2 = k;
            prep_stage_set(stg, 1, unsigned(div_round_up(range, 1 << k)), unsigned(div_round_up(g.nrows, FB_THREADS >> k)), unsigned(fused_pairs));
        }
        prep_stage_set(stg, 2, unsigned(div_round_up(g.lp_pitch / 4, 256)), unsigned(div_round_up(g.J, LP_ROWS)), nwalk);
        prep_u8_fn pf = prep_u8_pick(g.R);
        if (!pf) { set_error("no prep kernel for R=%d (internal)", g.R); return STEREO_ERR_UNSUPPORTED; }
        if (fused_pairs && !fill_by_memset) {
            stg.fill_words = unsigned(part_words); stg.fill_value = ncc ? 0 : -1;
            prep_stage_set(stg, 3, unsigned(div_round_up((long long)part_words, FILL_WORDS_PER_CTA)), 1, unsigned(fused_pairs));
        }
        ctx->last_launches += prep_launch(pf, kp, stg, st);
        if (three_pass) {   // RQ rows, vertical sums in HBM, horizontal sums
            prep_rq_kernel<<<dim3(div_round_up(g.rq_pitch / 4, 256), g.J / 2, nz), 256, 0, st>>>(kp);
            prep_v_kernel<<<dim3(div_round_up(vpitch, 128), div_round_up(g.nrows, PV_ROWS), nz), 128, 0, st>>>(kp, vpitch, 0);
            const size_t pe_smem = size_t(PE_ROWS) * (PE_COLS + 2 * g.R) * sizeof(int);
            prep_e2_kernel<<<dim3(div_round_up(g.e2_pitch, PE_COLS), div_round_up(g.nrows, PE_ROWS), nz), PE_COLS, pe_smem, st>>>(kp, vpitch);
            ctx->last_launches += 3;
        }
        if (ncc && fused_pairs) {
            // fused NCC pairs scale every key per pixel, from the other direction's RS rows inside the hot kernel
        } else if (ncc && (!legacy_prep) && fast_launch_is_pairs(ps, n)) {
            // both directions of every pair are in the launch: the reference-image energies are the partner's RS rows
            prep_scale_pair_kernel<<<dim3(div_round_up(g.tilesX * g.spc, 128), g.nrows, nz), 128, 0, st>>>(kp, n / 2);
            ctx->last_launches += 1;
        } else if (ncc) {   // window energies of the reference image -> per strip-row key binade (V is free again after prep_e2)
            prep_v_kernel<<<dim3(div_round_up(vpitch, 128), div_round_up(g.nrows, PV_ROWS), nz), 128, 0, st>>>(kp, vpitch, 1);
            prep_scale_kernel<<<dim3(div_round_up(g.tilesX * g.spc, 128), g.nrows, nz), 128, 0, st>>>(kp, vpitch);
            ctx->last_launches += 2;
        }
        fn = fused_pairs ? (ncc ? fast_pick_fused_ncc(g.R, g.hs) : fast_pick_fused(g.R, g.hs, fast_fused_all_mode1(g) ? 0 : 1)) : fast_pick(ps[0].cost, g.R, g.hs);
    }
    if (!fn) { set_error("no hot kernel for R=%d hs=%d opf=%d (internal)", g.R, g.hs, g.opf); return STEREO_ERR_UNSUPPORTED; }
    if (fused_pairs) ctx->fused_pairs_done += fused_pairs;     // (the partners' memsets are not counted as kernel launches)
    const int hot = ctx->hot_used < stereo_ctx::HOT_EVENTS ? ctx->hot_used : -1;
    if (hot >= 0) cudaEventRecord(ctx->hot0[hot], st);
    fn<<<g.ctas, g.nw * 32, fast_smem_bytes(g), st>>>(kp);
    if (hot >= 0) { cudaEventRecord(ctx->hot1[hot], st); ctx->hot_used++; ctx->hot_jobs += n; }
    ctx->hot_total++;
    if (ncc) fast_merge_ncc_kernel<<<dim3(div_round_up(g.cols, MG_COLS), g.nrows, nz), MG_THREADS, 0, st>>>(kp);
    else     fast_merge_ssd_kernel<<<dim3(div_round_up(g.cols, MG_COLS), g.nrows, nz), MG_THREADS, 0, st>>>(kp);
    ctx->last_launches += 2;
    SB_CUDA(cudaGetLastError());
    return STEREO_OK;
}

static inline int run_fast(stereo_ctx* ctx, const Problem& p, cudaStream_t st) { return run_fast_batch(ctx, &p, 1, st); }

} // namespace sb
