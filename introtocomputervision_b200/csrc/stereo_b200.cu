// stereo_b200 — C ABI implementation (context, validation, path selection, staging).
// Kernels: exact.cuh (general float path), fast.cuh (packed u8 path).  sm_100a only.
#include "common.cuh"
#include "exact.cuh"
#include "fast.cuh"
#include "compat.cuh"
#include "imgproc.cuh"
#include "host_pack.hpp"

#include <atomic>
#include <chrono>
#include <cstdarg>
#include <thread>
#include <new>
#include <vector>

namespace sb {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int Arena::reserve(size_t bytes) {
    if (bytes <= cap) return STEREO_OK;
    release();
    size_t want = bytes + (bytes >> 3) + (1u << 20);
    cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&base), want);
    if (e != cudaSuccess) {
        base = nullptr; cap = 0;
        set_error("cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
        (void)cudaGetLastError();
        return STEREO_ERR_ALLOC;
    }
    cap = want;
    return STEREO_OK;
}

void Arena::release() {
    if (base) cudaFree(base);
    base = nullptr; cap = 0; used = 0;
}

static size_t align256(size_t v) { return (v + 255) & ~size_t(255); }
static int quiesce(stereo_ctx* ctx, cudaStream_t st);

// ---- validation ---------------------------------------------------------------------------------

static int validate(const Problem& p, size_t ref_min_step, size_t tgt_min_step) {
    if (!p.ref.ptr || !p.tgt.ptr || !p.disp.ptr) { set_error("null image/output pointer"); return STEREO_ERR_INVALID_ARG; }
    if (p.rows <= 0 || p.cols <= 0) { set_error("rows/cols must be positive (got %d x %d)", p.rows, p.cols); return STEREO_ERR_INVALID_ARG; }
    if (p.rows > 32768 || p.cols > 32768) { set_error("image larger than 32768 in a dimension"); return STEREO_ERR_UNSUPPORTED; }
    if (p.R < 0 || p.R > 64) { set_error("window_rad must be in [0, 64] (got %d)", p.R); return STEREO_ERR_INVALID_ARG; }
    if (p.cost != STEREO_COST_SSD && p.cost != STEREO_COST_NCORR) { set_error("unknown cost %d", p.cost); return STEREO_ERR_INVALID_ARG; }
    if (p.disp.elem != 1 && p.disp.elem != 2 && p.disp.elem != 4) { set_error("disp_elem_bytes must be 1, 2 or 4"); return STEREO_ERR_INVALID_ARG; }
    if (p.ref.step < ref_min_step || p.tgt.step < tgt_min_step) { set_error("image step smaller than a row"); return STEREO_ERR_INVALID_ARG; }
    if (p.disp.step < size_t(p.cols) * p.disp.elem) { set_error("disp_step smaller than a row"); return STEREO_ERR_INVALID_ARG; }
    if (p.best.ptr && p.best.step < size_t(p.cols) * 4) { set_error("best_step smaller than a row"); return STEREO_ERR_INVALID_ARG; }
    if (p.dmin > p.dmax) { set_error("min_disp (%d) > max_disp (%d)", p.dmin, p.dmax); return STEREO_ERR_INVALID_RANGE; }
    if (p.dmin < -32768 || p.dmax > 32768) { set_error("disparity range outside [-32768, 32768]"); return STEREO_ERR_INVALID_RANGE; }
    if (p.row_begin < 0 || p.row_end > p.rows || p.row_begin >= p.row_end) { set_error("bad row band [%d, %d)", p.row_begin, p.row_end); return STEREO_ERR_INVALID_ARG; }
    if (p.avail_begin < 0 || p.avail_end > p.rows || p.avail_begin > p.row_begin || p.avail_end < p.row_end) {
        set_error("available rows [%d, %d) do not cover the band [%d, %d)", p.avail_begin, p.avail_end, p.row_begin, p.row_end);
        return STEREO_ERR_INVALID_ARG;
    }
    if (p.cost == STEREO_COST_NCORR) {
        // Every pixel needs >= 1 candidate window; the reference would throw inside
        // cv::Mat::operator()(Rect) otherwise (DisparityNCorr.cpp:50-53).
        const int Wp = p.cols + 2 * p.R, w = 2 * p.R + 1;
        // ncand(x) is piecewise linear in x: checking both ends and the clamp knees suffices,
        // checking every column is cheap enough and obviously right.
        for (int x = p.R; x < Wp - p.R; ++x) {
            long s = long(x) + p.dmin - p.R; if (s < 0) s = 0;
            long e = long(x) + p.dmax + 1 + p.R; if (e > Wp) e = Wp;
            if (e - s - w + 1 <= 0) {
                set_error("NCC range [%d, %d] leaves column %d without a candidate", p.dmin, p.dmax, x - p.R);
                return STEREO_ERR_INVALID_RANGE;
            }
        }
    }
    return STEREO_OK;
}

// ---- exact path ---------------------------------------------------------------------------------

static size_t exact_scratch_bytes(const Problem& p) {
    const size_t Hp = p.rows + 2 * p.R, Wp = p.cols + 2 * p.R, guard = size_t(p.R) + 64;
    const size_t img = align256((Hp * Wp + 2 * guard) * sizeof(float));
    const size_t band = size_t(p.row_end - p.row_begin) * p.cols;
    return 2 * img + 2 * align256(band * 4) + 4096;
}

static int run_exact(stereo_ctx* ctx, const Problem& p, cudaStream_t st) {
    const int Hp = p.rows + 2 * p.R, Wp = p.cols + 2 * p.R;
    const size_t guard = size_t(p.R) + 64;
    const size_t img_elems = size_t(Hp) * Wp + 2 * guard;
    float* lbuf = static_cast<float*>(ctx->arena.take(img_elems * sizeof(float)));
    float* rbuf = static_cast<float*>(ctx->arena.take(img_elems * sizeof(float)));
    const int band_rows = p.row_end - p.row_begin;
    int32_t* d_disp = static_cast<int32_t*>(ctx->arena.take(size_t(band_rows) * p.cols * 4));
    uint32_t* d_best = static_cast<uint32_t*>(ctx->arena.take(size_t(band_rows) * p.cols * 4));
    if (!lbuf || !rbuf || !d_disp || !d_best) { set_error("scratch arena too small (internal)"); return STEREO_ERR_ALLOC; }
    // zero guards (the tgt guard stands in for the reference's out-of-allocation reads)
    SB_CUDA(cudaMemsetAsync(rbuf, 0, guard * sizeof(float), st));
    SB_CUDA(cudaMemsetAsync(rbuf + guard + size_t(Hp) * Wp, 0, guard * sizeof(float), st));
    float* Lp = lbuf + guard;
    float* Rp = rbuf + guard;
    dim3 pb(32, 8), pg(div_round_up(Wp, 32), div_round_up(Hp, 8));
    if (p.ref.type == PixType::F32)
        pad_replicate_kernel<float><<<pg, pb, 0, st>>>(static_cast<const float*>(p.ref.ptr), p.ref.step, p.rows, p.cols, p.R, Lp, Hp, Wp, p.avail_begin, p.avail_end);
    else
        pad_replicate_kernel<uint8_t><<<pg, pb, 0, st>>>(static_cast<const uint8_t*>(p.ref.ptr), p.ref.step, p.rows, p.cols, p.R, Lp, Hp, Wp, p.avail_begin, p.avail_end);
    if (p.tgt.type == PixType::F32)
        pad_replicate_kernel<float><<<pg, pb, 0, st>>>(static_cast<const float*>(p.tgt.ptr), p.tgt.step, p.rows, p.cols, p.R, Rp, Hp, Wp, p.avail_begin, p.avail_end);
    else
        pad_replicate_kernel<uint8_t><<<pg, pb, 0, st>>>(static_cast<const uint8_t*>(p.tgt.ptr), p.tgt.step, p.rows, p.cols, p.R, Rp, Hp, Wp, p.avail_begin, p.avail_end);
    ctx->last_launches += 2;
    dim3 kb(128), kg(div_round_up(p.cols, 128), band_rows);
    if (p.cost == STEREO_COST_SSD)
        ssd_exact_kernel<<<kg, kb, 0, st>>>(Lp, Rp, p.rows, p.cols, p.R, p.dmin, p.dmax, p.row_begin, d_disp,
                                            reinterpret_cast<int32_t*>(d_best), 0, p.cols);
    else
        ncorr_exact_kernel<<<kg, kb, 0, st>>>(Lp, Rp, p.rows, p.cols, p.R, p.dmin, p.dmax, p.row_begin, d_disp,
                                              reinterpret_cast<float*>(d_best));
    store_output_kernel<<<kg, kb, 0, st>>>(d_disp, d_best, band_rows, p.cols, p.disp.ptr, p.disp.step, p.disp.elem,
                                           p.best.ptr, p.best.step);
    ctx->last_launches += 2;
    SB_CUDA(cudaGetLastError());
    return STEREO_OK;
}

// ---- path selection -------------------------------------------------------------------------------

static bool fused_layout(const stereo_ctx* ctx, const Problem* ps, int n);

// Float images: which kernel family applies (decided from the device-side classification, see classify_convert_kernel).
enum { CLS_U8 = 0, CLS_F32_FAST = 1, CLS_EXACT = 2 };

static Problem as_u8(Problem p) { p.ref.type = PixType::U8; p.tgt.type = PixType::U8; return p; }

// Scratch for a float problem whose kernel family is only known after the classification: the larger of the three
// families' needs plus the u8 copies of both images.  `dirs` = directions that share one launch sequence.
static size_t f32_scratch_bytes(stereo_ctx* ctx, const Problem& p, int dirs) {
    size_t need = exact_scratch_bytes(p);
    const size_t f = size_t(dirs) * fast_scratch_bytes(ctx, p), u = size_t(dirs) * fast_scratch_bytes(ctx, as_u8(p));
    if (f > need) need = f;
    if (u > need) need = u;
    return need + 2 * align256(size_t(p.rows) * align256(p.cols)) + 4096;
}

// Converts both float images to u8 copies while classifying them, reads the verdict back (one 16-byte copy and a stream
// synchronize: the kernel family is a host decision) and says which family serves this image pair.
static int classify_images(stereo_ctx* ctx, const Problem& p, cudaStream_t st, uint8_t* a8, uint8_t* b8, size_t pitch, int* cls) {
    SB_CUDA(cudaMemsetAsync(ctx->d_flag, 0, 4 * sizeof(int), st));
    dim3 cb(32, 8), cg(div_round_up(p.cols, 32), div_round_up(p.rows, 8), 2);
    classify_convert_kernel<<<cg, cb, 0, st>>>(static_cast<const float*>(p.ref.ptr), p.ref.step, a8, static_cast<const float*>(p.tgt.ptr), p.tgt.step, b8,
                                               p.rows, p.cols, pitch, ctx->d_flag);
    ctx->last_launches += 1;
    SB_CUDA(cudaMemcpyAsync(ctx->h_flag, ctx->d_flag, 4 * sizeof(int), cudaMemcpyDeviceToHost, st));
    SB_CUDA(cudaEventRecord(ctx->ev_gap0, st));
    SB_CUDA(cudaStreamSynchronize(st));
    SB_CUDA(cudaEventRecord(ctx->ev_gap1, st));          // (the kernels chosen below follow this mark)
    ctx->gap_recorded = true;
    const unsigned* hf = reinterpret_cast<const unsigned*>(ctx->h_flag);
    if (hf[0] == 0 && ctx->force_path != STEREO_PATH_FAST_F32) { *cls = CLS_U8; return STEREO_OK; }
    *cls = CLS_EXACT;
    if (hf[0] & 2u) return STEREO_OK;                                  // a pixel that is not finite
    if (ctx->force_path == STEREO_PATH_FAST_U8) return STEREO_OK;      // (reported as "not applicable" by the caller)
    // value range: [0, 255] for the blocks without a non-8-bit pixel, the tracked extremes for the others
    double lo = 0.0, hi = 255.0;
    if (hf[1]) { const double v = -double(f32_from_ordered(hf[1])); if (v < lo) lo = v; }
    if (hf[2]) { const double v = double(f32_from_ordered(hf[2])); if (v > hi) hi = v; }
    const double range = (hi - lo) * (1.0 + 1e-6), wnd = double(2 * p.R + 1) * (2 * p.R + 1);
    bool ok;
    if (p.cost == STEREO_COST_SSD)
        // per element: round((l - r)^2) + 0.5 < 2^23 (mantissa trick of ssd_q_bits); per window: 128 * SSD + position < 2^32 - 1
        ok = range * range + 1.0 < 8388607.0 && wnd * (range * range + 1.0) < 33554432.0 - 8193.0;
    else
        ok = -lo < 1.0e6 && hi < 1.0e6;                                // float32 sums of products stay far from overflow
    if (ok) *cls = CLS_F32_FAST;
    return STEREO_OK;
}

// Enqueues one direction.
static int run_problem(stereo_ctx* ctx, Problem p, cudaStream_t st, bool reset_arena = true) {
    const bool is_f32 = p.ref.type == PixType::F32;
    const size_t min_step = size_t(p.cols) * (is_f32 ? 4 : 1);
    int rc = validate(p, min_step, min_step);
    if (rc != STEREO_OK) return rc;
    if (p.ref.type != p.tgt.type) { set_error("mixed pixel types (internal)"); return STEREO_ERR_INVALID_ARG; }

    const bool fast_ok = fast_supported(p) && ctx->force_path != STEREO_PATH_EXACT_F32;
    size_t need = exact_scratch_bytes(p);
    if (fast_ok) need = is_f32 ? f32_scratch_bytes(ctx, p, 1) : (need > fast_scratch_bytes(ctx, p) ? need : fast_scratch_bytes(ctx, p));
    if (need > ctx->arena.cap) {
        rc = quiesce(ctx, st);
        if (rc != STEREO_OK) return rc;
        rc = ctx->arena.reserve(need);
        if (rc != STEREO_OK) return rc;
    }
    if (reset_arena) ctx->arena.reset();

    int cls = is_f32 ? CLS_EXACT : CLS_U8;
    if (fast_ok && is_f32) {
        const size_t pitch = align256(p.cols);
        uint8_t* a8 = static_cast<uint8_t*>(ctx->arena.take(size_t(p.rows) * pitch));
        uint8_t* b8 = static_cast<uint8_t*>(ctx->arena.take(size_t(p.rows) * pitch));
        if (!a8 || !b8) { set_error("scratch arena too small (internal)"); return STEREO_ERR_ALLOC; }
        rc = classify_images(ctx, p, st, a8, b8, pitch, &cls);
        if (rc != STEREO_OK) return rc;
        if (cls == CLS_U8) { p.ref = ImageView{a8, pitch, PixType::U8}; p.tgt = ImageView{b8, pitch, PixType::U8}; }
    }
    if (fast_ok && cls != CLS_EXACT) {
        ctx->last_path = cls == CLS_U8 ? STEREO_PATH_FAST_U8 : STEREO_PATH_FAST_F32;
        return run_fast(ctx, p, st);
    }
    if (ctx->force_path == STEREO_PATH_FAST_U8 || ctx->force_path == STEREO_PATH_FAST_F32) {
        set_error("fast path forced but not applicable (pixel values or window/range outside what the running-sum kernels cover)");
        return STEREO_ERR_UNSUPPORTED;
    }
    ctx->last_path = STEREO_PATH_EXACT_F32;
    return run_exact(ctx, p, st);
}

// Both directions of ONE pair of float images (ps[0] left-referenced, ps[1] right-referenced, the same two device images):
// classified once, then both maps from one launch sequence - one cost volume where the pair is fusable.
static int run_pair_f32(stereo_ctx* ctx, Problem* ps, cudaStream_t st) {
    const Problem& p = ps[0];
    const size_t min_step = size_t(p.cols) * 4;
    int rc = STEREO_OK;
    for (int d = 0; d < 2 && rc == STEREO_OK; ++d) rc = validate(ps[d], min_step, min_step);
    if (rc != STEREO_OK) return rc;
    const size_t need = f32_scratch_bytes(ctx, p, 2);
    if (need > ctx->arena.cap) {
        rc = quiesce(ctx, st);
        if (rc != STEREO_OK) return rc;
        rc = ctx->arena.reserve(need);
        if (rc != STEREO_OK) return rc;
    }
    ctx->arena.reset();
    const size_t pitch = align256(p.cols);
    uint8_t* a8 = static_cast<uint8_t*>(ctx->arena.take(size_t(p.rows) * pitch));
    uint8_t* b8 = static_cast<uint8_t*>(ctx->arena.take(size_t(p.rows) * pitch));
    if (!a8 || !b8) { set_error("scratch arena too small (internal)"); return STEREO_ERR_ALLOC; }
    int cls = CLS_EXACT;
    rc = classify_images(ctx, p, st, a8, b8, pitch, &cls);
    if (rc != STEREO_OK) return rc;
    if (cls == CLS_EXACT) {
        if (ctx->force_path == STEREO_PATH_FAST_U8 || ctx->force_path == STEREO_PATH_FAST_F32) {
            set_error("fast path forced but not applicable (pixel values outside what the running-sum kernels cover)");
            return STEREO_ERR_UNSUPPORTED;
        }
        ctx->last_path = STEREO_PATH_EXACT_F32;
        for (int d = 0; d < 2; ++d) {
            ctx->arena.reset();
            rc = run_exact(ctx, ps[d], st);
            if (rc != STEREO_OK) return rc;
        }
        return STEREO_OK;
    }
    if (cls == CLS_U8) {
        // ps[0]: ref = left, tgt = right; ps[1] the other way round
        ps[0].ref = ImageView{a8, pitch, PixType::U8}; ps[0].tgt = ImageView{b8, pitch, PixType::U8};
        ps[1].ref = ImageView{b8, pitch, PixType::U8}; ps[1].tgt = ImageView{a8, pitch, PixType::U8};
    }
    ctx->last_path = cls == CLS_U8 ? STEREO_PATH_FAST_U8 : STEREO_PATH_FAST_F32;
    return run_fast_batch(ctx, ps, 2, st, fused_layout(ctx, ps, 2) ? 1 : 0);
}

// Enqueues `n` u8 problems that all take the packed path, several directions per launch sequence where they
// are batchable (same shape, band, window, cost and candidate count).  `batch_budget` bounds the scratch one
// launch sequence may hold.
// True when ps = n/2 left-referenced problems followed by their right-referenced partners, each pair fusable.
static bool fused_layout(const stereo_ctx* ctx, const Problem* ps, int n) {
    if (!ctx->fuse_pairs || n < 2 || (n & 1) || n > FMAXJOBS) return false;
    const int np = n / 2;
    for (int k = 0; k < np; ++k)
        if (!fast_pair_fusable(ps[k], ps[np + k]) || !fast_batchable(ps[0], ps[k])) return false;
    return true;
}

// One launch sequence over batchable problems, fused when they are whole pairs.
static int run_fast_auto(stereo_ctx* ctx, const Problem* ps, int n, cudaStream_t st) {
    return run_fast_batch(ctx, ps, n, st, fused_layout(ctx, ps, n) ? n / 2 : 0);
}

static int run_fast_jobs(stereo_ctx* ctx, const Problem* ps, int n, cudaStream_t st) {
    const size_t budget = size_t(3) << 30;
    if (fused_layout(ctx, ps, n)) {          // (n <= FMAXJOBS: one launch sequence)
        const size_t need = size_t(n) * fast_scratch_bytes(ctx, ps[0]);
        if (need > ctx->arena.cap) {
            int rc = quiesce(ctx, st);
            if (rc != STEREO_OK) return rc;
            rc = ctx->arena.reserve(need);
            if (rc != STEREO_OK) return rc;
        }
        ctx->arena.reset();
        ctx->last_path = STEREO_PATH_FAST_U8;
        return run_fast_batch(ctx, ps, n, st, n / 2);
    }
    int i = 0;
    while (i < n) {
        const size_t per_job = fast_scratch_bytes(ctx, ps[i]);
        int m = 1;
        while (i + m < n && m < FMAXJOBS && fast_batchable(ps[i], ps[i + m]) && size_t(m + 1) * per_job <= budget) ++m;
        const size_t need = size_t(m) * per_job;
        if (need > ctx->arena.cap) {
            int rc = quiesce(ctx, st);
            if (rc != STEREO_OK) return rc;
            rc = ctx->arena.reserve(need);
            if (rc != STEREO_OK) return rc;
        }
        ctx->arena.reset();
        ctx->last_path = STEREO_PATH_FAST_U8;
        int rc = run_fast_batch(ctx, ps + i, m, st);
        if (rc != STEREO_OK) return rc;
        i += m;
    }
    return STEREO_OK;
}

// Every call of a context shares one scratch arena (and the classification flag), so calls are ordered: a call
// enqueued on another stream than its predecessor first waits for the predecessor's end-of-call event.
static void begin_call(stereo_ctx* ctx, cudaStream_t st) {
    if (ctx->have_prev_call && ctx->prev_stream != st) cudaStreamWaitEvent(st, ctx->ev1, 0);
    ctx->prev_stream = st;
    ctx->last_launches = 0;
    ctx->last_ms = -1.f;
    ctx->last_path = STEREO_PATH_NONE;
    ctx->hot_used = 0;
    ctx->hot_total = 0;
    ctx->hot_jobs = 0;
    ctx->fused_pairs_done = 0;
    ctx->gap_recorded = false;
    cudaEventRecord(ctx->ev0, st);
}
static void end_call(stereo_ctx* ctx, cudaStream_t st) {
    cudaEventRecord(ctx->ev1, st);
    ctx->timing_pending = true;
    ctx->have_prev_call = true;
}

// Before the scratch arena is reallocated: nothing enqueued by earlier calls may still use it.
static int quiesce(stereo_ctx* ctx, cudaStream_t st) {
    SB_CUDA(cudaStreamSynchronize(st));
    SB_CUDA(cudaStreamSynchronize(ctx->stream));
    if (ctx->have_prev_call && ctx->prev_stream != st && ctx->prev_stream != ctx->stream) SB_CUDA(cudaEventSynchronize(ctx->ev1));
    return STEREO_OK;
}

static int check_ctx(stereo_ctx* ctx) {
    if (!ctx) { set_error("null context"); return STEREO_ERR_INVALID_ARG; }
    cudaError_t e = cudaSetDevice(ctx->device);
    if (e != cudaSuccess) { set_error("cudaSetDevice(%d): %s", ctx->device, cudaGetErrorString(e)); return STEREO_ERR_CUDA; }
    return STEREO_OK;
}

static int ensure_pinned(stereo_ctx* ctx, size_t bytes) {
    if (bytes <= ctx->pinned_cap) return STEREO_OK;
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    ctx->pinned = nullptr; ctx->pinned_cap = 0;
    cudaError_t e = cudaHostAlloc(&ctx->pinned, bytes, cudaHostAllocDefault);
    if (e != cudaSuccess) { set_error("cudaHostAlloc(%zu): %s", bytes, cudaGetErrorString(e)); (void)cudaGetLastError(); return STEREO_ERR_ALLOC; }
    ctx->pinned_cap = bytes;
    return STEREO_OK;
}

// ---- pipelined host entry points --------------------------------------------------------------------
// The reference wrapper uploads, computes and downloads strictly one after the other on pageable
// memory (DisparitySSD.cu:171-206).  Here a host call is cut into (pair, row band) work items that flow
// through three streams — H2D of band b+1, compute of band b (both directions) and D2H of band b-1 run
// concurrently — so a call costs about max(upload, compute, download) instead of their sum.  Bands are
// the same R(+1)-row-halo bands the multi-GPU sharding uses: bit-identical results.
// f32 images are converted to u8 band by band while the "is it really 8-bit" flag accumulates; it is
// read once at the end, and a non-8-bit image makes the caller redo the call on the exact path.

struct HostDir {          // one direction of a host pair job
    bool swap;            // false: ref = left, tgt = right; true: ref = right, tgt = left
    int dmin, dmax;
    void* const* out;     // per pair: host output pointer
};
struct HostPairIn { const void* left; size_t left_step; const void* right; size_t right_step; };

enum { PIPE_NOT_APPLICABLE = 1, PIPE_NOT_8BIT = 2 };

// STEREO_PIPE_TRACE=1: the pipeline's events keep timestamps and every pipelined call prints, per work item and band, when
// its upload, compute and download finished (ms after the call's first enqueue) - a diagnostic for the e2e numbers.
static bool pipe_trace() {
    static const bool on = getenv("STEREO_PIPE_TRACE") != nullptr;
    return on;
}

static int ensure_pipe(stereo_ctx* ctx, int events) {
    if (!ctx->s_in) SB_CUDA(cudaStreamCreateWithFlags(&ctx->s_in, cudaStreamNonBlocking));
    if (!ctx->s_out) SB_CUDA(cudaStreamCreateWithFlags(&ctx->s_out, cudaStreamNonBlocking));
    if (events > ctx->pipe_ev_cap) {
        const int cap = events + 64;
        cudaEvent_t* ev = new (std::nothrow) cudaEvent_t[cap]();
        if (!ev) { set_error("out of host memory"); return STEREO_ERR_ALLOC; }
        for (int i = 0; i < ctx->pipe_ev_cap; ++i) ev[i] = ctx->pipe_ev[i];
        for (int i = ctx->pipe_ev_cap; i < cap; ++i) {
            cudaError_t e = cudaEventCreateWithFlags(&ev[i], pipe_trace() ? cudaEventDefault : cudaEventDisableTiming);
            if (e != cudaSuccess) {
                for (int k = ctx->pipe_ev_cap; k < i; ++k) cudaEventDestroy(ev[k]);
                delete[] ev;
                set_error("cudaEventCreate: %s", cudaGetErrorString(e));
                return STEREO_ERR_CUDA;
            }
        }
        delete[] ctx->pipe_ev;
        ctx->pipe_ev = ev;
        ctx->pipe_ev_cap = cap;
    }
    return STEREO_OK;
}

static int pipe_bands(const stereo_ctx* ctx, int n_pairs, int rows) {
    if (ctx->pipe_bands > 0) return ctx->pipe_bands < rows ? ctx->pipe_bands : rows;
    if (rows < 512) return 1;                           // too small to be worth cutting
    if (n_pairs >= 4) return rows >= 1024 ? 2 : 1;      // enough pairs in flight; two bands shorten the fill and the drain (tools/e2e_sweep.py)
    int nb = (rows + 255) / 512;                        // ~512-row bands: the (2R+1)-row warm-up stays < 3 %
    return nb < 1 ? 1 : (nb > 8 ? 8 : nb);
}

// Pairs per work item of the pipeline.  Small images (BASELINE config 5: 1280x720, 64 disparities) ride several
// pairs per item so that one launch sequence carries several pairs, as the device batch entry point does (which takes
// up to FMAXJOBS / 2; items stay at PIPE_CPMAX so that uploads, kernels and downloads of different items overlap); large
// images stay one pair (or one band of a pair) per item.
constexpr int PIPE_CPMAX = 4;
static_assert(2 * PIPE_CPMAX <= FMAXJOBS, "a pipeline item is one launch sequence");
static int pipe_chunk_pairs(int n_pairs, int nb, int rows, int cols) {
    if (nb > 1) return 1;
    const long long px = (long long)rows * cols;
    long long cp = ((4ll << 20) + px - 1) / px;           // ~4 Mpix of reference image per item
    if (cp > PIPE_CPMAX) cp = PIPE_CPMAX;
    if (cp > n_pairs / 3) cp = n_pairs / 3;               // keep at least three items in flight
    return cp < 1 ? 1 : int(cp);
}


// Large images (4K): the first work item of a call starts with a short band and the last one ends with a short band, so
// that the part of a call nothing overlaps - the first upload and the last download - is an eighth of an image, and the
// items in between go as whole images (every launch sequence costs ~65 us of ramp-up, warm-up rows and tail whatever
// its size: tools/e2e_sweep.py, STEREO_PIPE_TRACE).
static bool pipe_ramped(const stereo_ctx* ctx, int rows, int cols) {
    return ctx->pipe_bands == 0 && rows >= 1024 && (long long)rows * cols >= (4ll << 20);
}

// Row-band boundaries of work item w of a pipelined call: ascending, from 0 to rows.
static void pipe_item_bounds(const stereo_ctx* ctx, int n_pairs, int n_items, int w, int rows, int cols, std::vector<int>& b) {
    b.clear();
    b.push_back(0);
    if (pipe_ramped(ctx, rows, cols)) {
        const int e = rows / 8;
        if (w == 0) { b.push_back(e); b.push_back(3 * e); }
        if (w == n_items - 1) { if (rows - 3 * e > b.back()) b.push_back(rows - 3 * e); b.push_back(rows - e); }
        b.push_back(rows);
        return;
    }
    const int nb = pipe_bands(ctx, n_pairs, rows);
    const int band_rows = (rows + nb - 1) / nb;
    for (int r = band_rows; r < rows; r += band_rows) b.push_back(r);
    b.push_back(rows);
}

static int pairs_host_pipelined_body(stereo_ctx* ctx, int cost, PixType type, int n_pairs, const HostPairIn* in,
                                     const HostDir* dirs, int n_dirs, int rows, int cols, int R, size_t disp_step, int elem);

// Failures inside the pipeline leave asynchronous copies from / to the CALLER's host buffers queued on the three
// streams: drain them before the error reaches the caller, who may free or reuse those buffers.
static int pairs_host_pipelined(stereo_ctx* ctx, int cost, PixType type, int n_pairs, const HostPairIn* in,
                                const HostDir* dirs, int n_dirs, int rows, int cols, int R, size_t disp_step, int elem) {
    const int rc = pairs_host_pipelined_body(ctx, cost, type, n_pairs, in, dirs, n_dirs, rows, cols, R, disp_step, elem);
    if (rc < STEREO_OK) {
        if (ctx->s_in) cudaStreamSynchronize(ctx->s_in);
        if (ctx->stream) cudaStreamSynchronize(ctx->stream);
        if (ctx->s_out) cudaStreamSynchronize(ctx->s_out);
        (void)cudaGetLastError();
    }
    return rc;
}

// Host-side packing of CV_32FC1 images (host_pack.cpp): on when the context has enough host threads to convert faster
// than the link would carry the floats (stereo_ctx_set_host_threads; automatic: min(16, cores / LOCAL_WORLD_SIZE) threads,
// packing from HOST_PACK_MIN_THREADS up).
constexpr int HOST_PACK_MIN_THREADS = 8;
static bool host_pack_enabled(stereo_ctx* ctx) {
    if (ctx->host_threads == 0) ctx->host_threads = default_host_threads();
    if (ctx->host_threads < 0) return false;                  // switched off
    if (ctx->host_threads < HOST_PACK_MIN_THREADS && !ctx->host_pack_forced) return false;
    if (!ctx->pool || ctx->pool->threads() != ctx->host_threads) {
        delete ctx->pool;
        ctx->pool = new (std::nothrow) HostPool(ctx->host_threads);
    }
    return ctx->pool != nullptr;
}

// A few pixels of a host float image: false as soon as one is not an integer in 0..255.  Noisy / contrast-scaled images
// (main.cpp:140-153,191-193) fail on the first samples, so they skip the optimistic 8-bit pipeline instead of running
// it, finding the flag set and being computed a second time by the float kernels.
static bool host_sample_is_8bit(const void* img, size_t step, int rows, int cols) {
    const int ny = rows < 8 ? rows : 8, nx = cols < 16 ? cols : 16;
    for (int iy = 0; iy < ny; ++iy) {
        const float* row = reinterpret_cast<const float*>(static_cast<const char*>(img) + size_t((long long)iy * (rows - 1) / (ny > 1 ? ny - 1 : 1)) * step);
        for (int ix = 0; ix < nx; ++ix) {
            const float v = row[(long long)ix * (cols - 1) / (nx > 1 ? nx - 1 : 1)];
            if (!(v >= 0.f && v <= 255.f && v == float(int(v)))) return false;
        }
    }
    return true;
}

static int pairs_host_pipelined_body(stereo_ctx* ctx, int cost, PixType type, int n_pairs, const HostPairIn* in,
                                     const HostDir* dirs, int n_dirs, int rows, int cols, int R, size_t disp_step, int elem) {
    if (ctx->force_path == STEREO_PATH_EXACT_F32 || ctx->force_path == STEREO_PATH_FAST_F32) return PIPE_NOT_APPLICABLE;
    const auto t_call = std::chrono::steady_clock::now();       // (trace only)
    double t_pack = 0.0, t_ring = 0.0;
    auto since = [](std::chrono::steady_clock::time_point t0) { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); };
    if (type == PixType::F32)
        for (int i = 0; i < n_pairs; ++i)
            if (!host_sample_is_8bit(in[i].left, in[i].left_step, rows, cols) || !host_sample_is_8bit(in[i].right, in[i].right_step, rows, cols))
                return PIPE_NOT_8BIT;
    // CV_32FC1 host images with enough host threads: converted to u8 on the HOST (checking that every pixel is 8-bit), so the
    // link carries 1 byte per pixel; otherwise the float rows are uploaded and converted / classified on the device.
    // (Sending a share of the rows as floats next to the host conversion was measured and dropped: the copy engine's reads
    // and the converting threads compete for the same host memory bandwidth - 4K x 4 pairs: 4.88 ms with none, 5.01 ms with a
    // quarter, 5.57 ms with half of the rows as floats.)
    const bool pack = type == PixType::F32 && host_pack_enabled(ctx);
    const bool any_f32 = type == PixType::F32 && !pack, any_pack = pack;
    const size_t f_pitch = align256(size_t(cols) * 4), u8_pitch = align256(cols), d_pitch = align256(size_t(cols) * elem);
    // validate each direction on a full-image problem (pointers only need to be non-null here)
    Problem full{};
    full.cost = cost; full.rows = rows; full.cols = cols; full.row_begin = 0; full.row_end = rows;
    full.avail_begin = 0; full.avail_end = rows; full.R = R;
    full.ref = ImageView{in, u8_pitch, PixType::U8}; full.tgt = full.ref;
    full.disp = OutView{ctx, d_pitch, elem}; full.best = OutView{nullptr, 0, 4};
    const int nb_uniform = pipe_bands(ctx, n_pairs, rows);
    const int cp = pipe_chunk_pairs(n_pairs, nb_uniform, rows, cols);
    const int n_items = (n_pairs + cp - 1) / cp;
    // row bands of every work item (pipe_item_bounds): item w owns bounds[boff[w]] .. bounds[boff[w + 1] - 1]
    std::vector<int> bounds, boff(n_items + 1, 0), tmpb;
    for (int w = 0; w < n_items; ++w) {
        pipe_item_bounds(ctx, n_pairs, n_items, w, rows, cols, tmpb);
        bounds.insert(bounds.end(), tmpb.begin(), tmpb.end());
        boff[w + 1] = int(bounds.size());
    }
    auto nbw = [&](int w) { return boff[w + 1] - boff[w] - 1; };                           // bands of item w
    std::vector<int> uoff(n_items + 1, 0);                                                 // units (item, band) before item w
    for (int w = 0; w < n_items; ++w) uoff[w + 1] = uoff[w] + nbw(w);
    const int n_units = uoff[n_items];
    size_t scratch = 0;
    int max_up = 0;                                       // most rows one unit uploads
    for (int d = 0; d < n_dirs; ++d) {
        full.dmin = dirs[d].dmin; full.dmax = dirs[d].dmax;
        int rc = validate(full, u8_pitch, u8_pitch);
        if (rc != STEREO_OK) return rc;
        if (!fast_supported(full)) return PIPE_NOT_APPLICABLE;
        // every band: the operand-row count depends on where the band starts relative to the FRPS-row stages
        for (int w = 0; w < n_items; ++w) {
            int up = 0;
            for (int b = 0; b < nbw(w); ++b) {
                Problem band = full; band.row_begin = bounds[boff[w] + b]; band.row_end = bounds[boff[w] + b + 1];
                const size_t need = size_t(n_dirs) * cp * fast_scratch_bytes(ctx, band);
                scratch = need > scratch ? need : scratch;
                const int up_hi = (b == nbw(w) - 1) ? rows : ((band.row_end + R + 16 < rows) ? band.row_end + R + 16 : rows);
                max_up = up_hi - up > max_up ? up_hi - up : max_up;
                up = up_hi;
            }
        }
    }
    constexpr int NSTG = 4;                               // pinned staging ring of the host-packed uploads
    int rc = ensure_pipe(ctx, 3 * n_units + 4 + NSTG);
    if (rc != STEREO_OK) return rc;
    cudaStream_t s_in = ctx->s_in, s_cmp = ctx->stream, s_out = ctx->s_out;
    // uploads go chunk by chunk (~a quarter of a large image): the first copy of a unit leaves while the host threads are
    // still converting the rest of it
    const int chunk_rows = cp > 1 || rows < 1024 ? max_up : ((rows + 3) / 4 < 128 ? 128 : (rows + 3) / 4);
    const int stg_rows = chunk_rows < max_up ? chunk_rows : max_up;
    const size_t stg_pitch = (size_t(cols) + 63) & ~size_t(63);
    const size_t stg_slot = 2 * size_t(cp) * stg_rows * stg_pitch;
    if (any_pack) { rc = ensure_pinned(ctx, NSTG * stg_slot); if (rc != STEREO_OK) return rc; }

    const int S = n_items < 3 ? n_items : 3;            // device slots (ring), one work item each
    const size_t pair_bytes = 2 * align256(u8_pitch * rows) + (any_f32 ? 2 * align256(f_pitch * rows) : 0) + size_t(n_dirs) * align256(d_pitch * rows);
    const size_t slot_bytes = size_t(cp) * pair_bytes;
    if (size_t(S) * slot_bytes + 1024 > ctx->io.cap || scratch > ctx->arena.cap) {
        SB_CUDA(cudaStreamSynchronize(s_in)); SB_CUDA(cudaStreamSynchronize(s_cmp)); SB_CUDA(cudaStreamSynchronize(s_out));
        if (size_t(S) * slot_bytes + 1024 > ctx->io.cap) { rc = ctx->io.reserve(size_t(S) * slot_bytes + 1024); if (rc != STEREO_OK) return rc; }
        if (scratch > ctx->arena.cap) { rc = ctx->arena.reserve(scratch); if (rc != STEREO_OK) return rc; }
    }
    ctx->io.reset();
    struct Slot { char* lf; char* rf; uint8_t* l8; uint8_t* r8; char* out[2]; };      // lf / rf: float rows awaiting conversion
    constexpr int CPMAX = PIPE_CPMAX;
    Slot slot[3][CPMAX] = {};
    for (int k = 0; k < S; ++k)
        for (int c = 0; c < cp; ++c) {
            Slot& sl = slot[k][c];
            sl.l8 = static_cast<uint8_t*>(ctx->io.take(u8_pitch * rows));
            sl.r8 = static_cast<uint8_t*>(ctx->io.take(u8_pitch * rows));
            if (any_f32) {
                sl.lf = static_cast<char*>(ctx->io.take(f_pitch * rows));
                sl.rf = static_cast<char*>(ctx->io.take(f_pitch * rows));
            }
            for (int d = 0; d < n_dirs; ++d) sl.out[d] = static_cast<char*>(ctx->io.take(d_pitch * rows));
            if (!sl.l8 || !sl.r8 || (any_f32 && (!sl.lf || !sl.rf)) || !sl.out[n_dirs - 1]) { set_error("io arena too small (internal)"); return STEREO_ERR_ALLOC; }
        }
    auto ev = [&](int item, int band, int kind) { return ctx->pipe_ev[3 * (uoff[item] + band) + kind + 1]; };   // kind: 0 in, 1 cmp, 2 out

    begin_call(ctx, s_cmp);
    // copy streams start after whatever the context's stream was doing with these buffers
    SB_CUDA(cudaEventRecord(ctx->pipe_ev[0], s_cmp));
    SB_CUDA(cudaStreamWaitEvent(s_in, ctx->pipe_ev[0], 0));
    SB_CUDA(cudaStreamWaitEvent(s_out, ctx->pipe_ev[0], 0));
    if (any_f32) SB_CUDA(cudaMemsetAsync(ctx->d_flag, 0, 4 * sizeof(int), s_cmp));
    ctx->last_path = STEREO_PATH_FAST_U8;
    auto stg_ev = [&](int k) { return ctx->pipe_ev[3 * n_units + 1 + k]; };
    struct FloatRows { int r0, n; };                      // float row ranges of the unit being uploaded (converted on the device)
    std::vector<FloatRows> frows;
    // host conversion runs one upload chunk ahead of the enqueueing thread: the chunks in the order the loops below meet them
    struct PackChunk { int w, r0, n; };
    std::vector<PackChunk> pchunks;
    if (any_pack)
        for (int w = 0; w < n_items; ++w) {
            int up = 0;
            for (int b = 0; b < nbw(w); ++b) {
                const int re = bounds[boff[w] + b + 1];
                const int up_hi = (b == nbw(w) - 1) ? rows : ((re + R + 16 < rows) ? re + R + 16 : rows);
                for (int r0 = up; r0 < up_hi; r0 += chunk_rows) pchunks.push_back(PackChunk{w, r0, up_hi - r0 < chunk_rows ? up_hi - r0 : chunk_rows});
                if (up_hi > up) up = up_hi;
            }
        }
    // STEREO_PACK_AHEAD=0 (diagnostic): convert each chunk only when the loop reaches it, as before the look-ahead
    static const bool pack_ahead = [] { const char* e = getenv("STEREO_PACK_AHEAD"); return !(e && atoi(e) == 0); }();
    PackAsync pa;
    struct PoolDrain { HostPool* p; ~PoolDrain() { if (p) p->end(); } } pool_drain{any_pack ? ctx->pool : nullptr};   // no conversion outlives `pa`
    int pk = 0, pk_begun = 0;
    auto pack_begin = [&](int idx) -> int {
        const PackChunk& ch = pchunks[idx];
        const int k = idx % NSTG;
        const auto t_r = std::chrono::steady_clock::now();
        if (idx >= NSTG) SB_CUDA(cudaEventSynchronize(stg_ev(k)));      // the upload that last read this staging slot
        t_ring += since(t_r);
        uint8_t* stg = static_cast<uint8_t*>(ctx->pinned) + size_t(k) * stg_slot;
        const int i0 = ch.w * cp, np = (n_pairs - i0 < cp) ? n_pairs - i0 : cp;
        PackJob jobs[2 * PIPE_CPMAX];
        for (int c = 0; c < np; ++c) {
            const HostPairIn& hp = in[i0 + c];
            jobs[2 * c] = PackJob{reinterpret_cast<const float*>(static_cast<const char*>(hp.left) + size_t(ch.r0) * hp.left_step), hp.left_step,
                                  stg + size_t(2 * c) * stg_rows * stg_pitch, stg_pitch, ch.n, cols};
            jobs[2 * c + 1] = PackJob{reinterpret_cast<const float*>(static_cast<const char*>(hp.right) + size_t(ch.r0) * hp.right_step), hp.right_step,
                                      stg + size_t(2 * c + 1) * stg_rows * stg_pitch, stg_pitch, ch.n, cols};
        }
        pa.begin(*ctx->pool, jobs, 2 * np);
        pk_begun = idx + 1;
        return STEREO_OK;
    };

    for (int w = 0; w < n_items; ++w) {
        const Slot* sl = slot[w % S];
        const int i0 = w * cp, np = (n_pairs - i0 < cp) ? n_pairs - i0 : cp;     // pairs i0 .. i0+np-1 ride this item
        int uploaded = 0;
        for (int b = 0; b < nbw(w); ++b) {
            const int rb = bounds[boff[w] + b], re = bounds[boff[w] + b + 1];
            // ---- upload the rows this band adds: window halo R, the +1 row of the SSD flat-index wrap, and the
            //      operand rows the FRPS-row pipeline stages round up to
            const int up_hi = (b == nbw(w) - 1) ? rows : ((re + R + 16 < rows) ? re + R + 16 : rows);
            if (b == 0 && w >= S) SB_CUDA(cudaStreamWaitEvent(s_in, ev(w - S, nbw(w - S) - 1, 1), 0));   // slot inputs free again
            frows.clear();
            for (int r0 = uploaded; r0 < up_hi; r0 += chunk_rows) {
                const int n = up_hi - r0 < chunk_rows ? up_hi - r0 : chunk_rows;
                if (type == PixType::U8) {
                    for (int c = 0; c < np; ++c) {
                        const HostPairIn& hp = in[i0 + c];
                        SB_CUDA(cudaMemcpy2DAsync(sl[c].l8 + size_t(r0) * u8_pitch, u8_pitch, static_cast<const char*>(hp.left) + size_t(r0) * hp.left_step,
                                                  hp.left_step, cols, n, cudaMemcpyHostToDevice, s_in));
                        SB_CUDA(cudaMemcpy2DAsync(sl[c].r8 + size_t(r0) * u8_pitch, u8_pitch, static_cast<const char*>(hp.right) + size_t(r0) * hp.right_step,
                                                  hp.right_step, cols, n, cudaMemcpyHostToDevice, s_in));
                    }
                    continue;
                }
                // float rows as they are (converted by classify_convert_kernel in front of the unit's kernels), or through the
                // pinned staging ring as 8-bit pixels
                const int nf = any_pack ? 0 : n;
                if (nf > 0) {
                    for (int c = 0; c < np; ++c) {
                        const HostPairIn& hp = in[i0 + c];
                        SB_CUDA(cudaMemcpy2DAsync(sl[c].lf + size_t(r0) * f_pitch, f_pitch, static_cast<const char*>(hp.left) + size_t(r0) * hp.left_step,
                                                  hp.left_step, size_t(cols) * 4, nf, cudaMemcpyHostToDevice, s_in));
                        SB_CUDA(cudaMemcpy2DAsync(sl[c].rf + size_t(r0) * f_pitch, f_pitch, static_cast<const char*>(hp.right) + size_t(r0) * hp.right_step,
                                                  hp.right_step, size_t(cols) * 4, nf, cudaMemcpyHostToDevice, s_in));
                    }
                    frows.push_back(FloatRows{r0, nf});
                }
                const int p0 = r0 + nf, npk = n - nf;
                if (npk > 0) {
                    // this chunk's conversion was started while the previous chunk's copies and launches were being enqueued
                    // (the very first one starts here); the next chunk's starts before this one's copies are enqueued
                    const int idx = pk++, k = idx % NSTG;
                    if (idx >= int(pchunks.size()) || pchunks[idx].w != w || pchunks[idx].r0 != p0 || pchunks[idx].n != npk) {
                        set_error("upload chunk list out of step (internal)"); return STEREO_ERR_INVALID_ARG;
                    }
                    if (pk_begun <= idx) { rc = pack_begin(idx); if (rc != STEREO_OK) return rc; }
                    const auto t_p = std::chrono::steady_clock::now();
                    const bool all8 = pa.end(*ctx->pool);
                    t_pack += since(t_p);
                    if (!all8) {
                        // a pixel that is not 8-bit: give the call to the float kernels (nothing of it has reached the caller's
                        // maps that will not be overwritten)
                        SB_CUDA(cudaStreamSynchronize(s_in)); SB_CUDA(cudaStreamSynchronize(s_cmp)); SB_CUDA(cudaStreamSynchronize(s_out));
                        return PIPE_NOT_8BIT;
                    }
                    if (pack_ahead && idx + 1 < int(pchunks.size())) { rc = pack_begin(idx + 1); if (rc != STEREO_OK) return rc; }
                    const uint8_t* stg = static_cast<const uint8_t*>(ctx->pinned) + size_t(k) * stg_slot;
                    for (int c = 0; c < np; ++c) {
                        SB_CUDA(cudaMemcpy2DAsync(sl[c].l8 + size_t(p0) * u8_pitch, u8_pitch, stg + size_t(2 * c) * stg_rows * stg_pitch, stg_pitch, cols, npk,
                                                  cudaMemcpyHostToDevice, s_in));
                        SB_CUDA(cudaMemcpy2DAsync(sl[c].r8 + size_t(p0) * u8_pitch, u8_pitch, stg + size_t(2 * c + 1) * stg_rows * stg_pitch, stg_pitch, cols, npk,
                                                  cudaMemcpyHostToDevice, s_in));
                    }
                    SB_CUDA(cudaEventRecord(stg_ev(k), s_in));
                }
            }
            if (up_hi > uploaded) uploaded = up_hi;
            SB_CUDA(cudaEventRecord(ev(w, b, 0), s_in));
            // ---- compute
            SB_CUDA(cudaStreamWaitEvent(s_cmp, ev(w, b, 0), 0));
            if (b == 0 && w >= S) SB_CUDA(cudaStreamWaitEvent(s_cmp, ev(w - S, nbw(w - S) - 1, 2), 0));  // slot outputs downloaded
            for (const FloatRows& fr : frows) {
                dim3 cb(32, 8), cg(div_round_up(cols, 32), div_round_up(fr.n, 8), 2);
                for (int c = 0; c < np; ++c) {
                    classify_convert_kernel<<<cg, cb, 0, s_cmp>>>(reinterpret_cast<const float*>(sl[c].lf + size_t(fr.r0) * f_pitch), f_pitch,
                                                                 sl[c].l8 + size_t(fr.r0) * u8_pitch,
                                                                 reinterpret_cast<const float*>(sl[c].rf + size_t(fr.r0) * f_pitch), f_pitch,
                                                                 sl[c].r8 + size_t(fr.r0) * u8_pitch, fr.n, cols, u8_pitch, ctx->d_flag);
                    ctx->last_launches += 1;
                }
            }
            {
                // direction-major: jobs with the same range sign are neighbours (they share offsets and pitches)
                Problem pd[FMAXJOBS];
                for (int d = 0; d < n_dirs; ++d)
                    for (int c = 0; c < np; ++c) {
                        Problem& p = pd[d * np + c];
                        p = full;
                        p.row_begin = rb; p.row_end = re;
                        p.dmin = dirs[d].dmin; p.dmax = dirs[d].dmax;
                        p.ref = ImageView{dirs[d].swap ? sl[c].r8 : sl[c].l8, u8_pitch, PixType::U8};
                        p.tgt = ImageView{dirs[d].swap ? sl[c].l8 : sl[c].r8, u8_pitch, PixType::U8};
                        p.disp = OutView{sl[c].out[d] + size_t(rb) * d_pitch, d_pitch, elem};
                    }
                const int nj = n_dirs * np;
                bool together = true;
                for (int k = 1; k < nj; ++k) together = together && fast_batchable(pd[0], pd[k]);
                if (together) {
                    ctx->arena.reset();
                    rc = run_fast_auto(ctx, pd, nj, s_cmp);
                } else {
                    for (int k = 0; k < nj && rc == STEREO_OK; ++k) { ctx->arena.reset(); rc = run_fast(ctx, pd[k], s_cmp); }
                }
                if (rc != STEREO_OK) return rc;
            }
            SB_CUDA(cudaEventRecord(ev(w, b, 1), s_cmp));
            // ---- download
            SB_CUDA(cudaStreamWaitEvent(s_out, ev(w, b, 1), 0));
            for (int c = 0; c < np; ++c)
                for (int d = 0; d < n_dirs; ++d)
                    SB_CUDA(cudaMemcpy2DAsync(static_cast<char*>(dirs[d].out[i0 + c]) + size_t(rb) * disp_step, disp_step, sl[c].out[d] + size_t(rb) * d_pitch, d_pitch,
                                              size_t(cols) * elem, re - rb, cudaMemcpyDeviceToHost, s_out));
            SB_CUDA(cudaEventRecord(ev(w, b, 2), s_out));
        }
    }
    if (any_f32) SB_CUDA(cudaMemcpyAsync(ctx->h_flag, ctx->d_flag, sizeof(int), cudaMemcpyDeviceToHost, s_cmp));
    end_call(ctx, s_cmp);
    const double t_enq = since(t_call);
    SB_CUDA(cudaStreamSynchronize(s_cmp));
    SB_CUDA(cudaStreamSynchronize(s_out));
    SB_CUDA(cudaStreamSynchronize(s_in));
    if (pipe_trace()) {
        fprintf(stderr, "[pipe] host: everything enqueued after %.3f ms (converting %.3f, waiting for a staging slot %.3f), streams idle after %.3f ms\n",
                t_enq, t_pack, t_ring, since(t_call));
        fprintf(stderr, "[pipe] %d pairs %dx%d, %d items, %d units, %s\n", n_pairs, rows, cols, n_items, n_units,
                any_pack ? "host pack" : (type == PixType::F32 ? "float upload" : "u8"));
        for (int w = 0; w < n_items; ++w)
            for (int b = 0; b < nbw(w); ++b) {
                float t[3] = {-1.f, -1.f, -1.f};
                for (int k = 0; k < 3; ++k) cudaEventElapsedTime(&t[k], ctx->pipe_ev[0], ev(w, b, k));
                fprintf(stderr, "[pipe]   item %d rows %d..%d: in %.3f  compute %.3f  out %.3f ms\n", w, bounds[boff[w] + b], bounds[boff[w] + b + 1], t[0], t[1], t[2]);
            }
        (void)cudaGetLastError();
    }
    if (any_f32 && *ctx->h_flag != 0) return PIPE_NOT_8BIT;
    return STEREO_OK;
}

// Host-buffer single direction: upload (2D async copies), compute, download, synchronize — the
// shape of cuda::disparitySSD (DisparitySSD.cu:171-206) without per-call allocation.
static int host_single(stereo_ctx* ctx, int cost, PixType type, const void* ref, size_t ref_step, const void* tgt,
                       size_t tgt_step, int rows, int cols, int R, int dmin, int dmax, void* disp_out,
                       size_t disp_step, int elem, void* best_out, size_t best_step) {
    int rc = check_ctx(ctx);
    if (rc != STEREO_OK) return rc;
    if (!ref || !tgt || !disp_out) { set_error("null pointer"); return STEREO_ERR_INVALID_ARG; }
    if (rows <= 0 || cols <= 0 || rows > 32768 || cols > 32768) { set_error("bad image size %d x %d", rows, cols); return STEREO_ERR_INVALID_ARG; }
    if (elem != 1 && elem != 2 && elem != 4) { set_error("disp_elem_bytes must be 1, 2 or 4"); return STEREO_ERR_INVALID_ARG; }
    const size_t px = type == PixType::F32 ? 4 : 1;
    if (ref_step < cols * px || tgt_step < cols * px || disp_step < size_t(cols) * elem || (best_out && best_step < size_t(cols) * 4)) {
        set_error("a step is smaller than its row"); return STEREO_ERR_INVALID_ARG;
    }
    if (!best_out) {      // pipelined band-by-band; falls through to the sequential path when the packed kernels do not apply
        const HostPairIn in{ref, ref_step, tgt, tgt_step};
        void* outs[1] = {disp_out};
        const HostDir dir{false, dmin, dmax, outs};
        rc = pairs_host_pipelined(ctx, cost, type, 1, &in, &dir, 1, rows, cols, R, disp_step, elem);
        if (rc <= STEREO_OK) return rc;
    }
    cudaStream_t st = ctx->stream;
    const size_t in_pitch = align256(cols * px), d_pitch = align256(size_t(cols) * elem), b_pitch = align256(size_t(cols) * 4);
    const size_t need = 2 * in_pitch * rows + d_pitch * rows + (best_out ? b_pitch * rows : 0) + 1024;
    if (need > ctx->io.cap) {
        SB_CUDA(cudaStreamSynchronize(st));
        rc = ctx->io.reserve(need);
        if (rc != STEREO_OK) return rc;
    }
    ctx->io.reset();
    char* d_ref = static_cast<char*>(ctx->io.take(in_pitch * rows));
    char* d_tgt = static_cast<char*>(ctx->io.take(in_pitch * rows));
    char* d_disp = static_cast<char*>(ctx->io.take(d_pitch * rows));
    char* d_best = best_out ? static_cast<char*>(ctx->io.take(b_pitch * rows)) : nullptr;
    SB_CUDA(cudaMemcpy2DAsync(d_ref, in_pitch, ref, ref_step, cols * px, rows, cudaMemcpyHostToDevice, st));
    SB_CUDA(cudaMemcpy2DAsync(d_tgt, in_pitch, tgt, tgt_step, cols * px, rows, cudaMemcpyHostToDevice, st));
    Problem p{};
    p.cost = cost;
    p.ref = ImageView{d_ref, in_pitch, type};
    p.tgt = ImageView{d_tgt, in_pitch, type};
    p.rows = rows; p.cols = cols; p.row_begin = 0; p.row_end = rows; p.avail_begin = 0; p.avail_end = rows;
    p.R = R; p.dmin = dmin; p.dmax = dmax;
    p.disp = OutView{d_disp, d_pitch, elem};
    p.best = OutView{d_best, b_pitch, 4};
    begin_call(ctx, st);
    rc = run_problem(ctx, p, st);
    end_call(ctx, st);
    if (rc != STEREO_OK) return rc;
    SB_CUDA(cudaMemcpy2DAsync(disp_out, disp_step, d_disp, d_pitch, size_t(cols) * elem, rows, cudaMemcpyDeviceToHost, st));
    if (best_out) SB_CUDA(cudaMemcpy2DAsync(best_out, best_step, d_best, b_pitch, size_t(cols) * 4, rows, cudaMemcpyDeviceToHost, st));
    SB_CUDA(cudaStreamSynchronize(st));
    return STEREO_OK;
}

static int device_single(stereo_ctx* ctx, int cost, PixType type, const void* ref, size_t ref_step, const void* tgt,
                         size_t tgt_step, int rows, int cols, int row_begin, int row_end, int R, int dmin, int dmax,
                         void* disp_out, size_t disp_step, int elem, void* best_out, size_t best_step, void* stream,
                         int avail_begin = 0, int avail_end = -1) {
    int rc = check_ctx(ctx);
    if (rc != STEREO_OK) return rc;
    cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : ctx->stream;
    Problem p{};
    p.cost = cost;
    p.ref = ImageView{ref, ref_step, type};
    p.tgt = ImageView{tgt, tgt_step, type};
    p.rows = rows; p.cols = cols; p.row_begin = row_begin; p.row_end = row_end;
    p.avail_begin = avail_begin; p.avail_end = avail_end < 0 ? rows : avail_end;
    p.R = R; p.dmin = dmin; p.dmax = dmax;
    p.disp = OutView{disp_out, disp_step, elem};
    p.best = OutView{best_out, best_step, 4};
    begin_call(ctx, st);
    rc = run_problem(ctx, p, st);
    end_call(ctx, st);
    return rc;
}

} // namespace sb

using namespace sb;

// =================================================================================================
// C ABI
// =================================================================================================

extern "C" {

int stereo_abi_version(void) { return STEREO_B200_ABI_VERSION; }

const char* stereo_last_error(void) { return g_err; }

const char* stereo_status_string(int status) {
    switch (status) {
    case STEREO_OK: return "ok";
    case STEREO_ERR_INVALID_ARG: return "invalid argument";
    case STEREO_ERR_INVALID_RANGE: return "invalid disparity range";
    case STEREO_ERR_NO_DEVICE: return "no sm_100 CUDA device";
    case STEREO_ERR_CUDA: return "CUDA error";
    case STEREO_ERR_ALLOC: return "allocation failed";
    case STEREO_ERR_UNSUPPORTED: return "unsupported parameters";
    default: return "unknown status";
    }
}

int stereo_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { (void)cudaGetLastError(); return 0; }
    int ok = 0;
    for (int i = 0; i < n; ++i) {
        int major = 0;
        if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, i) == cudaSuccess && major == 10) ++ok;
    }
    return ok;
}

int stereo_ctx_create(int device, stereo_ctx** ctx_out) {
    if (!ctx_out) { set_error("ctx_out is null"); return STEREO_ERR_INVALID_ARG; }
    *ctx_out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) {
        (void)cudaGetLastError();
        set_error("no CUDA device available (this library has no CPU fallback)");
        return STEREO_ERR_NO_DEVICE;
    }
    if (device < 0 || device >= n) { set_error("device %d out of range [0, %d)", device, n); return STEREO_ERR_INVALID_ARG; }
    int major = 0, sms = 0;
    SB_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
    SB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    if (major != 10) { set_error("device %d is sm_%d0; this library is built for sm_100a only", device, major); return STEREO_ERR_NO_DEVICE; }
    SB_CUDA(cudaSetDevice(device));
    stereo_ctx* c = new (std::nothrow) stereo_ctx();
    if (!c) { set_error("out of host memory"); return STEREO_ERR_ALLOC; }
    c->device = device;
    c->sm_count = sms;
    cudaError_t e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreate(&c->ev0);
    if (e == cudaSuccess) e = cudaEventCreate(&c->ev1);
    if (e == cudaSuccess) e = cudaEventCreate(&c->ev_gap0);
    if (e == cudaSuccess) e = cudaEventCreate(&c->ev_gap1);
    for (int i = 0; i < stereo_ctx::HOT_EVENTS && e == cudaSuccess; ++i) {
        e = cudaEventCreate(&c->hot0[i]);
        if (e == cudaSuccess) e = cudaEventCreate(&c->hot1[i]);
    }
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&c->d_flag), 256);
    if (e == cudaSuccess) e = cudaHostAlloc(reinterpret_cast<void**>(&c->h_flag), 256, cudaHostAllocDefault);
    if (e != cudaSuccess) {
        set_error("context setup failed: %s", cudaGetErrorString(e));
        stereo_ctx_destroy(c);
        return STEREO_ERR_CUDA;
    }
    int rc = fast_ctx_init(c);
    if (rc != STEREO_OK) { stereo_ctx_destroy(c); return rc; }
    *ctx_out = c;
    return STEREO_OK;
}

void stereo_ctx_destroy(stereo_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) { cudaStreamSynchronize(ctx->stream); }
    ctx->arena.release();
    ctx->io.release();
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    delete ctx->pool;
    if (ctx->d_flag) cudaFree(ctx->d_flag);
    if (ctx->h_flag) cudaFreeHost(ctx->h_flag);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->ev_gap0) cudaEventDestroy(ctx->ev_gap0);
    if (ctx->ev_gap1) cudaEventDestroy(ctx->ev_gap1);
    for (int i = 0; i < stereo_ctx::HOT_EVENTS; ++i) {
        if (ctx->hot0[i]) cudaEventDestroy(ctx->hot0[i]);
        if (ctx->hot1[i]) cudaEventDestroy(ctx->hot1[i]);
    }
    for (int i = 0; i < ctx->pipe_ev_cap; ++i) cudaEventDestroy(ctx->pipe_ev[i]);
    delete[] ctx->pipe_ev;
    for (int i = 0; i < stereo_ctx::PEER_STREAMS; ++i) if (ctx->s_peer[i]) { cudaStreamSynchronize(ctx->s_peer[i]); cudaStreamDestroy(ctx->s_peer[i]); }
    if (ctx->peer_ready) cudaEventDestroy(ctx->peer_ready);
    for (int t = 0; t < stereo_ctx::PEER_TICKETS; ++t)
        for (int i = 0; i < stereo_ctx::PEER_STREAMS; ++i) if (ctx->peer_done[t][i]) cudaEventDestroy(ctx->peer_done[t][i]);
    if (ctx->s_in) cudaStreamDestroy(ctx->s_in);
    if (ctx->s_out) cudaStreamDestroy(ctx->s_out);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

int stereo_ctx_last_path(const stereo_ctx* ctx) { return ctx ? ctx->last_path : 0; }

float stereo_ctx_last_kernel_ms(const stereo_ctx* ctx) {
    if (!ctx) return -1.f;
    stereo_ctx* c = const_cast<stereo_ctx*>(ctx);
    if (c->timing_pending) {
        float ms = -1.f;
        if (cudaEventSynchronize(c->ev1) == cudaSuccess && cudaEventElapsedTime(&ms, c->ev0, c->ev1) == cudaSuccess) {
            float gap = 0.f;      // float device calls: the host round trip of the classification verdict is not kernel time
            if (c->gap_recorded && cudaEventElapsedTime(&gap, c->ev_gap0, c->ev_gap1) == cudaSuccess && gap > 0.f && gap < ms) ms -= gap;
            c->last_ms = ms;
        }
        else (void)cudaGetLastError();
        c->timing_pending = false;
    }
    return c->last_ms;
}

int stereo_ctx_last_launches(const stereo_ctx* ctx) { return ctx ? ctx->last_launches : 0; }

float stereo_ctx_last_hot_kernel_ms(const stereo_ctx* ctx, int* launches_measured) {
    if (launches_measured) *launches_measured = 0;
    if (!ctx || ctx->hot_used == 0) return -1.f;
    float total = 0.f;
    for (int i = 0; i < ctx->hot_used; ++i) {
        float ms = 0.f;
        if (cudaEventSynchronize(ctx->hot1[i]) != cudaSuccess || cudaEventElapsedTime(&ms, ctx->hot0[i], ctx->hot1[i]) != cudaSuccess) {
            (void)cudaGetLastError();
            return -1.f;
        }
        total += ms;
    }
    if (launches_measured) *launches_measured = ctx->hot_used;
    return total;
}

int stereo_ctx_last_hot_jobs(const stereo_ctx* ctx) { return ctx ? ctx->hot_jobs : 0; }

int stereo_ctx_force_path(stereo_ctx* ctx, int path) {
    if (!ctx || path < 0 || path > STEREO_PATH_FAST_F32) { set_error("bad force_path argument"); return STEREO_ERR_INVALID_ARG; }
    ctx->force_path = path;
    return STEREO_OK;
}

int stereo_ctx_set_pipe_bands(stereo_ctx* ctx, int bands) {
    if (!ctx || bands < 0 || bands > 4096) { set_error("bad pipe_bands argument"); return STEREO_ERR_INVALID_ARG; }
    ctx->pipe_bands = bands;
    return STEREO_OK;
}

// Pure host arithmetic (no device): how a host batch call is cut into pipeline items.
int stereo_host_pipeline_plan(int n_pairs, int rows, int cols, int bands_override, int* bands_per_pair, int* pairs_per_item) {
    if (n_pairs <= 0 || rows <= 0 || cols <= 0 || bands_override < 0 || !bands_per_pair || !pairs_per_item) {
        set_error("bad arguments"); return STEREO_ERR_INVALID_ARG;
    }
    stereo_ctx tmp;
    tmp.pipe_bands = bands_override;
    const int nb = pipe_bands(&tmp, n_pairs, rows);
    const int cp = pipe_chunk_pairs(n_pairs, nb, rows, cols);
    const int n_items = (n_pairs + cp - 1) / cp;
    std::vector<int> b;
    int most = 0;
    for (int w = 0; w < n_items; w = (w == 0 && n_items > 2) ? n_items - 1 : w + 1) {      // first and last item hold the most bands
        pipe_item_bounds(&tmp, n_pairs, n_items, w, rows, cols, b);
        most = int(b.size()) - 1 > most ? int(b.size()) - 1 : most;
    }
    *bands_per_pair = most;
    *pairs_per_item = cp;
    return STEREO_OK;
}

int stereo_host_pipeline_item_bands(int n_pairs, int rows, int cols, int bands_override, int item, int* bounds, int cap) {
    if (n_pairs <= 0 || rows <= 0 || cols <= 0 || bands_override < 0 || item < 0 || !bounds || cap < 2) {
        set_error("bad arguments"); return STEREO_ERR_INVALID_ARG;
    }
    stereo_ctx tmp;
    tmp.pipe_bands = bands_override;
    const int cp = pipe_chunk_pairs(n_pairs, pipe_bands(&tmp, n_pairs, rows), rows, cols);
    const int n_items = (n_pairs + cp - 1) / cp;
    if (item >= n_items) { set_error("item %d of %d", item, n_items); return STEREO_ERR_INVALID_ARG; }
    std::vector<int> b;
    pipe_item_bounds(&tmp, n_pairs, n_items, item, rows, cols, b);
    if (int(b.size()) > cap) { set_error("%d boundaries, room for %d", int(b.size()), cap); return STEREO_ERR_INVALID_ARG; }
    for (size_t i = 0; i < b.size(); ++i) bounds[i] = b[i];
    return int(b.size());
}

// Pure host arithmetic (no device): the geometry of the hot-kernel launch a batch of `n_pairs` pair problems would get.
int stereo_launch_plan(int cost, int float_operands, int n_pairs, int rows, int cols, int window_rad, int disparity_range, int fuse,
                       int sm_count, stereo_launch_plan_t* plan) {
    if (!plan || n_pairs < 1 || 2 * n_pairs > FMAXJOBS || rows <= 0 || cols <= 0 || window_rad < 0 || window_rad > FMAXR || disparity_range < 0 ||
        sm_count < 1 || (cost != STEREO_COST_SSD && cost != STEREO_COST_NCORR)) {
        set_error("bad launch plan arguments"); return STEREO_ERR_INVALID_ARG;
    }
    stereo_ctx tmp;
    tmp.sm_count = sm_count;
    Problem ps[FMAXJOBS];
    static const char dummy_l = 0, dummy_r = 0;
    for (int k = 0; k < n_pairs; ++k) {
        Problem two[2];
        const PixType t = float_operands ? PixType::F32 : PixType::U8;
        two[0] = Problem{};
        two[0].cost = cost; two[0].rows = rows; two[0].cols = cols; two[0].row_begin = 0; two[0].row_end = rows;
        two[0].avail_begin = 0; two[0].avail_end = rows; two[0].R = window_rad;
        two[1] = two[0];
        two[0].ref = ImageView{&dummy_l + k, size_t(cols), t}; two[0].tgt = ImageView{&dummy_r + k, size_t(cols), t};
        two[0].dmin = -disparity_range; two[0].dmax = 0;
        two[1].ref = two[0].tgt; two[1].tgt = two[0].ref; two[1].dmin = 0; two[1].dmax = disparity_range;
        ps[k] = two[0]; ps[n_pairs + k] = two[1];
    }
    if (!fast_supported(ps[0])) { set_error("outside what the running-sum kernels cover"); return STEREO_ERR_UNSUPPORTED; }
    tmp.fuse_pairs = fuse ? 1 : 0;
    const bool fused = fused_layout(&tmp, ps, 2 * n_pairs);
    FastKernelParams kp{};
    fast_geometry(&tmp, ps, 2 * n_pairs, kp.g, kp.job, fused ? n_pairs : 0);
    const FastGeom& g = kp.g;
    plan->fused = fused ? 1 : 0;
    plan->strip_px = g.K; plan->strips_per_warp = g.hs; plan->groups = g.G; plan->tile_px = g.spc * g.K; plan->tiles = g.ntiles;
    plan->ctas = g.ctas; plan->schedule = g.sched == 1 ? 2 : (g.nrl != g.nrows ? 1 : 0); plan->rows_per_item = int(g.L);
    plan->bands = g.sched == 1 ? g.nbands : 0; plan->stages = g.nst; plan->smem_bytes = int(fast_smem_bytes(g)); plan->border_kernel = g.border;
    return STEREO_OK;
}

int stereo_ctx_last_fused_pairs(const stereo_ctx* ctx) { return ctx ? ctx->fused_pairs_done : 0; }

// Pure host code (no device): the conversion the pipelined CV_32FC1 entry points run on their host threads.
int stereo_host_pack_f32_u8(const float* src, size_t src_step, uint8_t* dst, size_t dst_step, int rows, int cols, int threads, int* all_8bit) {
    if (!src || !dst || !all_8bit || rows <= 0 || cols <= 0 || src_step < size_t(cols) * 4 || dst_step < size_t(cols) || threads < 1 || threads > 256) {
        set_error("bad pack arguments"); return STEREO_ERR_INVALID_ARG;
    }
    HostPool pool(threads);
    *all_8bit = pack_f32_u8(pool, src, src_step, dst, dst_step, rows, cols) ? 1 : 0;
    return STEREO_OK;
}

int stereo_host_pool_selftest(int threads, int rounds) {
    if (threads < 1 || threads > 256 || rounds < 1) { set_error("bad selftest arguments"); return STEREO_ERR_INVALID_ARG; }
    HostPool pool(threads);
    std::vector<std::atomic<int>> hits(257);
    for (int r = 0; r < rounds; ++r) {
        const int n_tasks = 1 + (r * 37) % 257;
        for (int i = 0; i < n_tasks; ++i) hits[i].store(0, std::memory_order_relaxed);
        // odd rounds: the asynchronous form the host pipeline uses - the caller does something else between begin() and end()
        if (r & 1) {
            pool.begin(n_tasks, [&](int t) { hits[t].fetch_add(1, std::memory_order_relaxed); });
            volatile int spin = 0;
            for (int i = 0; i < (r % 7) * 300; ++i) spin = spin + i;
            pool.end();
        } else {
            pool.run(n_tasks, [&](int t) { hits[t].fetch_add(1, std::memory_order_relaxed); });
        }
        for (int i = 0; i < n_tasks; ++i)
            if (hits[i].load() != 1) { set_error("round %d: task %d of %d ran %d times", r, i, n_tasks, hits[i].load()); return STEREO_ERR_UNSUPPORTED; }
        if (r % 64 == 63) std::this_thread::sleep_for(std::chrono::microseconds(600));     // let the workers fall asleep now and then
    }
    return STEREO_OK;
}

int stereo_ctx_set_host_threads(stereo_ctx* ctx, int threads) {
    if (!ctx || threads < -1 || threads > 256) { set_error("bad host_threads argument"); return STEREO_ERR_INVALID_ARG; }
    ctx->host_threads = threads;
    ctx->host_pack_forced = threads > 0;
    return STEREO_OK;
}

int stereo_ctx_host_threads(const stereo_ctx* ctx) {
    if (!ctx) return 0;
    return ctx->host_threads == 0 ? default_host_threads() : ctx->host_threads;
}

int stereo_ctx_set_fuse_pairs(stereo_ctx* ctx, int on) {
    if (!ctx) { set_error("null context"); return STEREO_ERR_INVALID_ARG; }
    ctx->fuse_pairs = on ? 1 : 0;
    return STEREO_OK;
}

// ---- peer gather: results pushed into other ranks' buffers by the copy engines ---------------------
static int ensure_peer(stereo_ctx* ctx) {
    if (ctx->peer_ready) return STEREO_OK;
    for (int i = 0; i < stereo_ctx::PEER_STREAMS; ++i) SB_CUDA(cudaStreamCreateWithFlags(&ctx->s_peer[i], cudaStreamNonBlocking));
    for (int t = 0; t < stereo_ctx::PEER_TICKETS; ++t)
        for (int i = 0; i < stereo_ctx::PEER_STREAMS; ++i) SB_CUDA(cudaEventCreateWithFlags(&ctx->peer_done[t][i], cudaEventDisableTiming));
    SB_CUDA(cudaEventCreateWithFlags(&ctx->peer_ready, cudaEventDisableTiming));
    return STEREO_OK;
}

int stereo_peer_buffer_create(stereo_ctx* ctx, size_t bytes, void** dev_ptr, unsigned char* handle_out) {
    int rc = check_ctx(ctx);
    if (rc != STEREO_OK) return rc;
    if (!dev_ptr || !handle_out || bytes == 0) { set_error("bad peer buffer arguments"); return STEREO_ERR_INVALID_ARG; }
    static_assert(sizeof(cudaIpcMemHandle_t) == STEREO_IPC_HANDLE_BYTES, "IPC handle size");
    void* p = nullptr;
    if (cudaMalloc(&p, bytes) != cudaSuccess) { (void)cudaGetLastError(); set_error("peer buffer of %zu bytes: out of device memory", bytes); return STEREO_ERR_ALLOC; }
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) { cudaFree(p); set_error("cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e)); return STEREO_ERR_CUDA; }
    memcpy(handle_out, &h, sizeof(h));
    *dev_ptr = p;
    return STEREO_OK;
}

int stereo_peer_buffer_open(stereo_ctx* ctx, const unsigned char* handle, void** peer_ptr) {
    int rc = check_ctx(ctx);
    if (rc != STEREO_OK) return rc;
    if (!handle || !peer_ptr) { set_error("bad peer buffer arguments"); return STEREO_ERR_INVALID_ARG; }
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    SB_CUDA(cudaIpcOpenMemHandle(peer_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return STEREO_OK;
}

int stereo_peer_buffer_close(stereo_ctx* ctx, void* peer_ptr) {
    int rc = check_ctx(ctx);
    if (rc != STEREO_OK) return rc;
    if (peer_ptr) SB_CUDA(cudaIpcCloseMemHandle(peer_ptr));
    return STEREO_OK;
}

int stereo_peer_buffer_destroy(stereo_ctx* ctx, void* dev_ptr) {
    int rc = check_ctx(ctx);
    if (rc != STEREO_OK) return rc;
    if (dev_ptr) SB_CUDA(cudaFree(dev_ptr));
    return STEREO_OK;
}

int stereo_peer_push(stereo_ctx* ctx, void* const* dst_ptrs, int n_dst, size_t dst_offset, const void* src, size_t bytes,
                     void* after_stream) {
    int rc = check_ctx(ctx);
    if (rc != STEREO_OK) return rc;
    if (!dst_ptrs || n_dst < 0 || (!src && bytes)) { set_error("bad peer push arguments"); return STEREO_ERR_INVALID_ARG; }
    rc = ensure_peer(ctx);
    if (rc != STEREO_OK) return rc;
    cudaStream_t producer = after_stream ? static_cast<cudaStream_t>(after_stream) : ctx->stream;
    SB_CUDA(cudaEventRecord(ctx->peer_ready, producer));
    for (int i = 0; i < stereo_ctx::PEER_STREAMS; ++i) SB_CUDA(cudaStreamWaitEvent(ctx->s_peer[i], ctx->peer_ready, 0));
    for (int i = 0; i < n_dst; ++i) {
        if (!dst_ptrs[i]) continue;
        cudaStream_t s = ctx->s_peer[ctx->peer_next_stream];
        ctx->peer_next_stream = (ctx->peer_next_stream + 1) % stereo_ctx::PEER_STREAMS;
        SB_CUDA(cudaMemcpyAsync(static_cast<char*>(dst_ptrs[i]) + dst_offset, src, bytes, cudaMemcpyDeviceToDevice, s));
    }
    return STEREO_OK;
}

int stereo_peer_mark(stereo_ctx* ctx, int* ticket_out) {
    int rc = check_ctx(ctx);
    if (rc != STEREO_OK) return rc;
    if (!ticket_out) { set_error("ticket_out is null"); return STEREO_ERR_INVALID_ARG; }
    rc = ensure_peer(ctx);
    if (rc != STEREO_OK) return rc;
    const int t = ctx->peer_next_ticket;
    ctx->peer_next_ticket = (t + 1) % stereo_ctx::PEER_TICKETS;
    for (int i = 0; i < stereo_ctx::PEER_STREAMS; ++i) SB_CUDA(cudaEventRecord(ctx->peer_done[t][i], ctx->s_peer[i]));
    *ticket_out = t;
    return STEREO_OK;
}

int stereo_peer_wait(stereo_ctx* ctx, int ticket, void* cuda_stream) {
    int rc = check_ctx(ctx);
    if (rc != STEREO_OK) return rc;
    if (ticket < 0 || ticket >= stereo_ctx::PEER_TICKETS || !ctx->peer_ready) { set_error("bad peer ticket"); return STEREO_ERR_INVALID_ARG; }
    cudaStream_t s = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : ctx->stream;
    for (int i = 0; i < stereo_ctx::PEER_STREAMS; ++i) SB_CUDA(cudaStreamWaitEvent(s, ctx->peer_done[ticket][i], 0));
    return STEREO_OK;
}

int stereo_ctx_synchronize(stereo_ctx* ctx, void* cuda_stream) {
    int rc = check_ctx(ctx);
    if (rc != STEREO_OK) return rc;
    SB_CUDA(cudaStreamSynchronize(cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : ctx->stream));
    return STEREO_OK;
}

int stereo_disparity_f32_host(stereo_ctx* ctx, int cost, const float* ref, size_t ref_step, const float* tgt,
                              size_t tgt_step, int rows, int cols, int window_rad, int min_disp, int max_disp,
                              void* disp_out, size_t disp_step, int disp_elem_bytes, void* best_out, size_t best_step) {
    return host_single(ctx, cost, PixType::F32, ref, ref_step, tgt, tgt_step, rows, cols, window_rad, min_disp,
                       max_disp, disp_out, disp_step, disp_elem_bytes, best_out, best_step);
}

int stereo_disparity_u8_host(stereo_ctx* ctx, int cost, const uint8_t* ref, size_t ref_step, const uint8_t* tgt,
                             size_t tgt_step, int rows, int cols, int window_rad, int min_disp, int max_disp,
                             void* disp_out, size_t disp_step, int disp_elem_bytes, void* best_out, size_t best_step) {
    return host_single(ctx, cost, PixType::U8, ref, ref_step, tgt, tgt_step, rows, cols, window_rad, min_disp,
                       max_disp, disp_out, disp_step, disp_elem_bytes, best_out, best_step);
}

// ---- device images and preprocessing (imgproc.cuh) ---------------------------------------------------------------------------
int stereo_dev_alloc(stereo_ctx* ctx, size_t bytes, void** ptr) {
    int rc = check_ctx(ctx);
    if (rc != STEREO_OK) return rc;
    if (!ptr || bytes == 0) { set_error("bad allocation arguments"); return STEREO_ERR_INVALID_ARG; }
    if (cudaMalloc(ptr, bytes) != cudaSuccess) { (void)cudaGetLastError(); *ptr = nullptr; set_error("device allocation of %zu bytes failed", bytes); return STEREO_ERR_ALLOC; }
    return STEREO_OK;
}

int stereo_dev_free(stereo_ctx* ctx, void* ptr) {
    int rc = check_ctx(ctx);
    if (rc != STEREO_OK) return rc;
    if (ptr) { SB_CUDA(cudaStreamSynchronize(ctx->stream)); SB_CUDA(cudaFree(ptr)); }
    return STEREO_OK;
}

int stereo_dev_upload(stereo_ctx* ctx, void* dst, size_t dst_step, const void* src, size_t src_step, size_t row_bytes, int rows) {
    int rc = check_ctx(ctx);
    if (rc != STEREO_OK) return rc;
    if (!dst || !src || rows <= 0 || row_bytes == 0 || dst_step < row_bytes || src_step < row_bytes) { set_error("bad upload arguments"); return STEREO_ERR_INVALID_ARG; }
    SB_CUDA(cudaMemcpy2DAsync(dst, dst_step, src, src_step, row_bytes, rows, cudaMemcpyHostToDevice, ctx->stream));
    SB_CUDA(cudaStreamSynchronize(ctx->stream));          // the host buffer is the caller's again
    return STEREO_OK;
}

int stereo_dev_download(stereo_ctx* ctx, void* dst, size_t dst_step, const void* src, size_t src_step, size_t row_bytes, int rows) {
    int rc = check_ctx(ctx);
    if (rc != STEREO_OK) return rc;
    if (!dst || !src || rows <= 0 || row_bytes == 0 || dst_step < row_bytes || src_step < row_bytes) { set_error("bad download arguments"); return STEREO_ERR_INVALID_ARG; }
    SB_CUDA(cudaMemcpy2DAsync(dst, dst_step, src, src_step, row_bytes, rows, cudaMemcpyDeviceToHost, ctx->stream));
    SB_CUDA(cudaStreamSynchronize(ctx->stream));
    return STEREO_OK;
}

int stereo_image_gray_f32_device(stereo_ctx* ctx, const uint8_t* img, size_t step, int rows, int cols, int channels, int shift,
                                 float* out, size_t out_step, void* cuda_stream) {
    int rc = check_ctx(ctx);
    if (rc != STEREO_OK) return rc;
    if (!img || !out || rows <= 0 || cols <= 0 || (channels != 1 && channels != 3 && channels != 4) || (shift != 14 && shift != 15) ||
        step < size_t(cols) * channels || out_step < size_t(cols) * 4) { set_error("bad gray conversion arguments"); return STEREO_ERR_INVALID_ARG; }
    cudaStream_t st = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : ctx->stream;
    gray_f32_kernel<<<dim3(div_round_up(cols, 256), rows), 256, 0, st>>>(img, step, rows, cols, channels, shift, out, out_step);
    SB_CUDA(cudaGetLastError());
    return STEREO_OK;
}

int stereo_image_scale_add_f32_device(stereo_ctx* ctx, const float* a, size_t a_step, const float* add, size_t add_step, float scale,
                                      int rows, int cols, float* out, size_t out_step, void* cuda_stream) {
    int rc = check_ctx(ctx);
    if (rc != STEREO_OK) return rc;
    if (!a || !out || rows <= 0 || cols <= 0 || a_step < size_t(cols) * 4 || out_step < size_t(cols) * 4 || (add && add_step < size_t(cols) * 4)) {
        set_error("bad scale/add arguments"); return STEREO_ERR_INVALID_ARG;
    }
    cudaStream_t st = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : ctx->stream;
    scale_add_f32_kernel<<<dim3(div_round_up(cols, 256), rows), 256, 0, st>>>(a, a_step, add, add_step, scale, rows, cols, out, out_step);
    SB_CUDA(cudaGetLastError());
    return STEREO_OK;
}

// ---- reference-GPU-semantics mode (compat.cuh) ------------------------------------------------------------------------------
int stereo_disparity_refgpu_f32_device(stereo_ctx* ctx, int cost, const float* ref, size_t ref_step, const float* tgt, size_t tgt_step,
                                       int rows, int cols, int window_rad, int min_disp, int max_disp, int8_t* disp_out,
                                       size_t disp_step, float* best_out, size_t best_step, void* cuda_stream) {
    int rc = check_ctx(ctx);
    if (rc != STEREO_OK) return rc;
    if (!ref || !tgt || !disp_out) { set_error("null pointer"); return STEREO_ERR_INVALID_ARG; }
    if (rows <= 0 || cols <= 0 || rows > 32768 || cols > 32768) { set_error("bad image size %d x %d", rows, cols); return STEREO_ERR_INVALID_ARG; }
    if (cost != STEREO_COST_SSD && cost != STEREO_COST_NCORR) { set_error("unknown cost %d", cost); return STEREO_ERR_INVALID_ARG; }
    if (ref_step < size_t(cols) * 4 || tgt_step < size_t(cols) * 4 || disp_step < size_t(cols) || (best_out && best_step < size_t(cols) * 4)) {
        set_error("a step is smaller than its row"); return STEREO_ERR_INVALID_ARG;
    }
    if (window_rad < 0 || window_rad > 64) { set_error("window_rad must be in [0, 64]"); return STEREO_ERR_INVALID_ARG; }
    if (min_disp > max_disp) { set_error("min_disp (%d) > max_disp (%d)", min_disp, max_disp); return STEREO_ERR_INVALID_RANGE; }
    // the reference's loop variable is a `char` (DisparitySSD.cu:41,56): ranges beyond int8 never terminate there
    if (min_disp < -128 || max_disp > 126) { set_error("the reference's GPU kernels index disparities with a char: range must lie in [-128, 126]"); return STEREO_ERR_INVALID_RANGE; }
    cudaStream_t st = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : ctx->stream;
    begin_call(ctx, st);
    const dim3 grid(div_round_up(cols, CG_TILE), div_round_up(rows, CG_ROWS));
    const int threads = CG_TILE + 2 * window_rad;
    const size_t smem = size_t(3) * threads * sizeof(float);
    if (cost == STEREO_COST_SSD)
        refgpu_kernel<STEREO_COST_SSD><<<grid, threads, smem, st>>>(ref, ref_step, tgt, tgt_step, rows, cols, window_rad, min_disp, max_disp, disp_out, disp_step, best_out, best_step);
    else
        refgpu_kernel<STEREO_COST_NCORR><<<grid, threads, smem, st>>>(ref, ref_step, tgt, tgt_step, rows, cols, window_rad, min_disp, max_disp, disp_out, disp_step, best_out, best_step);
    ctx->last_launches += 1;
    ctx->last_path = STEREO_PATH_REFGPU;
    end_call(ctx, st);
    SB_CUDA(cudaGetLastError());
    return STEREO_OK;
}

int stereo_disparity_refgpu_f32_host(stereo_ctx* ctx, int cost, const float* ref, size_t ref_step, const float* tgt, size_t tgt_step,
                                     int rows, int cols, int window_rad, int min_disp, int max_disp, int8_t* disp_out,
                                     size_t disp_step, float* best_out, size_t best_step) {
    int rc = check_ctx(ctx);
    if (rc != STEREO_OK) return rc;
    if (!ref || !tgt || !disp_out || rows <= 0 || cols <= 0 || rows > 32768 || cols > 32768) { set_error("bad arguments"); return STEREO_ERR_INVALID_ARG; }
    if (ref_step < size_t(cols) * 4 || tgt_step < size_t(cols) * 4) { set_error("a step is smaller than its row"); return STEREO_ERR_INVALID_ARG; }
    cudaStream_t st = ctx->stream;
    const size_t in_pitch = align256(size_t(cols) * 4), d_pitch = align256(cols);
    const size_t need = 3 * in_pitch * rows + d_pitch * rows + 1024;
    if (need > ctx->io.cap) {
        SB_CUDA(cudaStreamSynchronize(st));
        rc = ctx->io.reserve(need);
        if (rc != STEREO_OK) return rc;
    }
    ctx->io.reset();
    char* d_ref = static_cast<char*>(ctx->io.take(in_pitch * rows));
    char* d_tgt = static_cast<char*>(ctx->io.take(in_pitch * rows));
    char* d_best = static_cast<char*>(ctx->io.take(in_pitch * rows));
    char* d_disp = static_cast<char*>(ctx->io.take(d_pitch * rows));
    SB_CUDA(cudaMemcpy2DAsync(d_ref, in_pitch, ref, ref_step, size_t(cols) * 4, rows, cudaMemcpyHostToDevice, st));
    SB_CUDA(cudaMemcpy2DAsync(d_tgt, in_pitch, tgt, tgt_step, size_t(cols) * 4, rows, cudaMemcpyHostToDevice, st));
    rc = stereo_disparity_refgpu_f32_device(ctx, cost, reinterpret_cast<const float*>(d_ref), in_pitch, reinterpret_cast<const float*>(d_tgt), in_pitch,
                                            rows, cols, window_rad, min_disp, max_disp, reinterpret_cast<int8_t*>(d_disp), d_pitch,
                                            best_out ? reinterpret_cast<float*>(d_best) : nullptr, in_pitch, st);
    if (rc == STEREO_OK) {
        cudaError_t e = cudaMemcpy2DAsync(disp_out, disp_step, d_disp, d_pitch, cols, rows, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess && best_out) e = cudaMemcpy2DAsync(best_out, best_step, d_best, in_pitch, size_t(cols) * 4, rows, cudaMemcpyDeviceToHost, st);
        if (e != cudaSuccess) { set_error("download failed: %s", cudaGetErrorString(e)); rc = STEREO_ERR_CUDA; }
    }
    cudaError_t e = cudaStreamSynchronize(st);
    if (rc == STEREO_OK && e != cudaSuccess) { set_error("compat call failed: %s", cudaGetErrorString(e)); rc = STEREO_ERR_CUDA; }
    return rc;
}

int stereo_disparity_f32_device(stereo_ctx* ctx, int cost, const float* ref, size_t ref_step, const float* tgt,
                                size_t tgt_step, int rows, int cols, int window_rad, int min_disp, int max_disp,
                                void* disp_out, size_t disp_step, int disp_elem_bytes, void* best_out,
                                size_t best_step, void* cuda_stream) {
    return device_single(ctx, cost, PixType::F32, ref, ref_step, tgt, tgt_step, rows, cols, 0, rows, window_rad,
                         min_disp, max_disp, disp_out, disp_step, disp_elem_bytes, best_out, best_step, cuda_stream);
}

int stereo_disparity_u8_device(stereo_ctx* ctx, int cost, const uint8_t* ref, size_t ref_step, const uint8_t* tgt,
                               size_t tgt_step, int rows, int cols, int window_rad, int min_disp, int max_disp,
                               void* disp_out, size_t disp_step, int disp_elem_bytes, void* best_out,
                               size_t best_step, void* cuda_stream) {
    return device_single(ctx, cost, PixType::U8, ref, ref_step, tgt, tgt_step, rows, cols, 0, rows, window_rad,
                         min_disp, max_disp, disp_out, disp_step, disp_elem_bytes, best_out, best_step, cuda_stream);
}

int stereo_disparity_band_u8_device(stereo_ctx* ctx, int cost, const uint8_t* ref, size_t ref_step,
                                    const uint8_t* tgt, size_t tgt_step, int rows, int cols, int row_begin,
                                    int row_end, int window_rad, int min_disp, int max_disp, void* disp_out,
                                    size_t disp_step, int disp_elem_bytes, void* cuda_stream) {
    return device_single(ctx, cost, PixType::U8, ref, ref_step, tgt, tgt_step, rows, cols, row_begin, row_end,
                         window_rad, min_disp, max_disp, disp_out, disp_step, disp_elem_bytes, nullptr, 0, cuda_stream);
}

int stereo_disparity_band_halo_u8_device(stereo_ctx* ctx, int cost, const uint8_t* ref_halo, size_t ref_step,
                                         const uint8_t* tgt_halo, size_t tgt_step, int rows, int cols, int row_begin,
                                         int row_end, int halo_begin, int halo_end, int window_rad, int min_disp,
                                         int max_disp, void* disp_out, size_t disp_step, int disp_elem_bytes,
                                         void* cuda_stream) {
    if (!ref_halo || !tgt_halo) { set_error("null image pointer"); return STEREO_ERR_INVALID_ARG; }
    if (halo_begin < 0 || halo_end > rows || halo_begin >= halo_end) { set_error("bad halo rows [%d, %d)", halo_begin, halo_end); return STEREO_ERR_INVALID_ARG; }
    // Full-image origins that are never dereferenced outside [halo_begin, halo_end): every row index is clamped
    // into the available range, which changes nothing when the halo covers what the band needs.
    const uint8_t* ref0 = ref_halo - size_t(halo_begin) * ref_step;
    const uint8_t* tgt0 = tgt_halo - size_t(halo_begin) * tgt_step;
    return device_single(ctx, cost, PixType::U8, ref0, ref_step, tgt0, tgt_step, rows, cols, row_begin, row_end,
                         window_rad, min_disp, max_disp, disp_out, disp_step, disp_elem_bytes, nullptr, 0, cuda_stream,
                         halo_begin, halo_end);
}

int stereo_band_halo_rows(int rows, int row_begin, int row_end, int window_rad, int* halo_begin, int* halo_end) {
    if (!halo_begin || !halo_end || rows <= 0 || row_begin < 0 || row_end > rows || row_begin >= row_end || window_rad < 0) {
        set_error("bad arguments"); return STEREO_ERR_INVALID_ARG;
    }
    // R rows of window above and below, +1 for the neighbouring padded row the reference's SSD reads
    // through its flat index; clamped to the image (replicate padding).
    int b = row_begin - window_rad - 1, e = row_end + window_rad + 1;
    *halo_begin = b < 0 ? 0 : b;
    *halo_end = e > rows ? rows : e;
    return STEREO_OK;
}

// ---- pairs: L->R over [-range, 0], then R->L with images swapped over [0, +range] (main.cpp:21-48) ----

static void pair_problems(Problem* p, int cost, PixType type, const void* left, size_t left_step, const void* right,
                          size_t right_step, int rows, int cols, int R, int range, void* disp_left, void* disp_right,
                          size_t disp_step, int elem) {
    p[0] = Problem{};
    p[0].cost = cost;
    p[0].rows = rows; p[0].cols = cols; p[0].row_begin = 0; p[0].row_end = rows; p[0].avail_begin = 0; p[0].avail_end = rows; p[0].R = R;
    p[0].best = OutView{nullptr, 0, 4};
    p[1] = p[0];
    p[0].ref = ImageView{left, left_step, type}; p[0].tgt = ImageView{right, right_step, type};
    p[0].dmin = -range; p[0].dmax = 0; p[0].disp = OutView{disp_left, disp_step, elem};
    p[1].ref = ImageView{right, right_step, type}; p[1].tgt = ImageView{left, left_step, type};
    p[1].dmin = 0; p[1].dmax = range; p[1].disp = OutView{disp_right, disp_step, elem};
}

// true when every problem can go straight to the packed kernels (u8 images, supported window / range)
static bool all_fast(const stereo_ctx* ctx, const Problem* ps, int n) {
    if (ctx->force_path == STEREO_PATH_EXACT_F32) return false;
    for (int i = 0; i < n; ++i)
        if (ps[i].ref.type != PixType::U8 || ps[i].tgt.type != PixType::U8 || !fast_supported(ps[i])) return false;
    return true;
}

static int pair_device(stereo_ctx* ctx, int cost, PixType type, const void* left, size_t left_step, const void* right,
                       size_t right_step, int rows, int cols, int R, int range, void* disp_left, void* disp_right,
                       size_t disp_step, int elem, cudaStream_t st) {
    if (range < 0) { set_error("disparity_range must be >= 0"); return STEREO_ERR_INVALID_RANGE; }
    Problem p[2];
    pair_problems(p, cost, type, left, left_step, right, right_step, rows, cols, R, range, disp_left, disp_right, disp_step, elem);
    if (all_fast(ctx, p, 2)) {       // both directions in one launch sequence
        for (int d = 0; d < 2; ++d) {
            int rc = validate(p[d], size_t(cols), size_t(cols));
            if (rc != STEREO_OK) return rc;
        }
        return run_fast_jobs(ctx, p, 2, st);
    }
    if (type == PixType::F32 && fast_supported(p[0]) && fast_supported(p[1]) && ctx->force_path != STEREO_PATH_EXACT_F32)
        return run_pair_f32(ctx, p, st);
    int rc = run_problem(ctx, p[0], st);
    if (rc != STEREO_OK) return rc;
    return run_problem(ctx, p[1], st);
}

static int pair_host(stereo_ctx* ctx, int cost, PixType type, const void* left, size_t left_step, const void* right,
                     size_t right_step, int rows, int cols, int R, int range, void* disp_left, void* disp_right,
                     size_t disp_step, int elem) {
    int rc = check_ctx(ctx);
    if (rc != STEREO_OK) return rc;
    if (!left || !right || !disp_left || !disp_right) { set_error("null pointer"); return STEREO_ERR_INVALID_ARG; }
    if (rows <= 0 || cols <= 0 || rows > 32768 || cols > 32768) { set_error("bad image size %d x %d", rows, cols); return STEREO_ERR_INVALID_ARG; }
    if (elem != 1 && elem != 2 && elem != 4) { set_error("disp_elem_bytes must be 1, 2 or 4"); return STEREO_ERR_INVALID_ARG; }
    const size_t px = type == PixType::F32 ? 4 : 1;
    if (left_step < cols * px || right_step < cols * px || disp_step < size_t(cols) * elem) { set_error("a step is smaller than its row"); return STEREO_ERR_INVALID_ARG; }
    if (range < 0) { set_error("disparity_range must be >= 0"); return STEREO_ERR_INVALID_RANGE; }
    {
        const HostPairIn in{left, left_step, right, right_step};
        void* outs_l[1] = {disp_left};
        void* outs_r[1] = {disp_right};
        const HostDir dirs[2] = {{false, -range, 0, outs_l}, {true, 0, range, outs_r}};
        rc = pairs_host_pipelined(ctx, cost, type, 1, &in, dirs, 2, rows, cols, R, disp_step, elem);
        if (rc <= STEREO_OK) return rc;
    }
    cudaStream_t st = ctx->stream;
    const size_t in_pitch = align256(cols * px), d_pitch = align256(size_t(cols) * elem);
    const size_t need = 2 * in_pitch * rows + 2 * d_pitch * rows + 1024;
    if (need > ctx->io.cap) {
        SB_CUDA(cudaStreamSynchronize(st));
        rc = ctx->io.reserve(need);
        if (rc != STEREO_OK) return rc;
    }
    ctx->io.reset();
    char* d_l = static_cast<char*>(ctx->io.take(in_pitch * rows));
    char* d_r = static_cast<char*>(ctx->io.take(in_pitch * rows));
    char* d_dl = static_cast<char*>(ctx->io.take(d_pitch * rows));
    char* d_dr = static_cast<char*>(ctx->io.take(d_pitch * rows));
    SB_CUDA(cudaMemcpy2DAsync(d_l, in_pitch, left, left_step, cols * px, rows, cudaMemcpyHostToDevice, st));
    SB_CUDA(cudaMemcpy2DAsync(d_r, in_pitch, right, right_step, cols * px, rows, cudaMemcpyHostToDevice, st));
    begin_call(ctx, st);
    rc = pair_device(ctx, cost, type, d_l, in_pitch, d_r, in_pitch, rows, cols, R, range, d_dl, d_dr, d_pitch, elem, st);
    end_call(ctx, st);
    if (rc != STEREO_OK) return rc;
    SB_CUDA(cudaMemcpy2DAsync(disp_left, disp_step, d_dl, d_pitch, size_t(cols) * elem, rows, cudaMemcpyDeviceToHost, st));
    SB_CUDA(cudaMemcpy2DAsync(disp_right, disp_step, d_dr, d_pitch, size_t(cols) * elem, rows, cudaMemcpyDeviceToHost, st));
    SB_CUDA(cudaStreamSynchronize(st));
    return STEREO_OK;
}

int stereo_disparity_pair_f32_host(stereo_ctx* ctx, int cost, const float* left, size_t left_step, const float* right,
                                   size_t right_step, int rows, int cols, int window_rad, int disparity_range,
                                   void* disp_left, void* disp_right, size_t disp_step, int disp_elem_bytes) {
    return pair_host(ctx, cost, PixType::F32, left, left_step, right, right_step, rows, cols, window_rad,
                     disparity_range, disp_left, disp_right, disp_step, disp_elem_bytes);
}

int stereo_disparity_pair_u8_host(stereo_ctx* ctx, int cost, const uint8_t* left, size_t left_step,
                                  const uint8_t* right, size_t right_step, int rows, int cols, int window_rad,
                                  int disparity_range, void* disp_left, void* disp_right, size_t disp_step,
                                  int disp_elem_bytes) {
    return pair_host(ctx, cost, PixType::U8, left, left_step, right, right_step, rows, cols, window_rad,
                     disparity_range, disp_left, disp_right, disp_step, disp_elem_bytes);
}

int stereo_disparity_pair_u8_device(stereo_ctx* ctx, int cost, const uint8_t* left, size_t left_step,
                                    const uint8_t* right, size_t right_step, int rows, int cols, int window_rad,
                                    int disparity_range, void* disp_left, void* disp_right, size_t disp_step,
                                    int disp_elem_bytes, void* cuda_stream) {
    int rc = check_ctx(ctx);
    if (rc != STEREO_OK) return rc;
    cudaStream_t st = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : ctx->stream;
    begin_call(ctx, st);
    rc = pair_device(ctx, cost, PixType::U8, left, left_step, right, right_step, rows, cols, window_rad,
                     disparity_range, disp_left, disp_right, disp_step, disp_elem_bytes, st);
    end_call(ctx, st);
    return rc;
}

int stereo_disparity_pair_f32_device(stereo_ctx* ctx, int cost, const float* left, size_t left_step, const float* right,
                                     size_t right_step, int rows, int cols, int window_rad, int disparity_range,
                                     void* disp_left, void* disp_right, size_t disp_step, int disp_elem_bytes,
                                     void* cuda_stream) {
    int rc = check_ctx(ctx);
    if (rc != STEREO_OK) return rc;
    cudaStream_t st = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : ctx->stream;
    begin_call(ctx, st);
    rc = pair_device(ctx, cost, PixType::F32, left, left_step, right, right_step, rows, cols, window_rad,
                     disparity_range, disp_left, disp_right, disp_step, disp_elem_bytes, st);
    end_call(ctx, st);
    return rc;
}

// Both maps of a row band from one launch sequence (one rank's share of a row-band sharded PAIR, BASELINE config 4).
int stereo_disparity_pair_band_halo_u8_device(stereo_ctx* ctx, int cost, const uint8_t* left_halo, size_t left_step,
                                              const uint8_t* right_halo, size_t right_step, int rows, int cols,
                                              int row_begin, int row_end, int halo_begin, int halo_end, int window_rad,
                                              int disparity_range, void* disp_left, void* disp_right, size_t disp_step,
                                              int disp_elem_bytes, void* cuda_stream) {
    int rc = check_ctx(ctx);
    if (rc != STEREO_OK) return rc;
    if (!left_halo || !right_halo || !disp_left || !disp_right) { set_error("null pointer"); return STEREO_ERR_INVALID_ARG; }
    if (halo_begin < 0 || halo_end > rows || halo_begin >= halo_end) { set_error("bad halo rows [%d, %d)", halo_begin, halo_end); return STEREO_ERR_INVALID_ARG; }
    if (disparity_range < 0) { set_error("disparity_range must be >= 0"); return STEREO_ERR_INVALID_RANGE; }
    cudaStream_t st = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : ctx->stream;
    const uint8_t* l0 = left_halo - size_t(halo_begin) * left_step;       // full-image origins, see above
    const uint8_t* r0 = right_halo - size_t(halo_begin) * right_step;
    Problem p[2];
    pair_problems(p, cost, PixType::U8, l0, left_step, r0, right_step, rows, cols, window_rad, disparity_range, disp_left, disp_right,
                  disp_step, disp_elem_bytes);
    for (int d = 0; d < 2; ++d) { p[d].row_begin = row_begin; p[d].row_end = row_end; p[d].avail_begin = halo_begin; p[d].avail_end = halo_end; }
    begin_call(ctx, st);
    if (all_fast(ctx, p, 2)) {
        for (int d = 0; d < 2 && rc == STEREO_OK; ++d) rc = validate(p[d], size_t(cols), size_t(cols));
        if (rc == STEREO_OK) rc = run_fast_jobs(ctx, p, 2, st);
    } else {
        rc = run_problem(ctx, p[0], st);
        if (rc == STEREO_OK) rc = run_problem(ctx, p[1], st);
    }
    end_call(ctx, st);
    return rc;
}

// ---- one row band of a pair from HOST images (a device's share of a row-band sharded pair) ------------------------------
// Uploads only the slab of rows the band needs (window halo + the row of the SSD flat-index wrap), computes both maps of
// the band in one launch sequence and downloads them into the caller's band rows.  CV_32FC1 images are converted to u8 on
// the host while the slab is staged; an image that is not 8-bit-valued makes the call return STEREO_ERR_UNSUPPORTED (the
// band entry points are 8-bit only; callers then take a whole-image call).
static int pair_band_host(stereo_ctx* ctx, int cost, PixType type, const void* left, size_t left_step, const void* right,
                          size_t right_step, int rows, int cols, int row_begin, int row_end, int R, int range,
                          void* disp_left, void* disp_right, size_t disp_step, int elem) {
    int rc = check_ctx(ctx);
    if (rc != STEREO_OK) return rc;
    if (!left || !right || !disp_left || !disp_right) { set_error("null pointer"); return STEREO_ERR_INVALID_ARG; }
    if (rows <= 0 || cols <= 0 || rows > 32768 || cols > 32768) { set_error("bad image size %d x %d", rows, cols); return STEREO_ERR_INVALID_ARG; }
    if (row_begin < 0 || row_end > rows || row_begin >= row_end) { set_error("bad row band [%d, %d)", row_begin, row_end); return STEREO_ERR_INVALID_ARG; }
    if (elem != 1 && elem != 2 && elem != 4) { set_error("disp_elem_bytes must be 1, 2 or 4"); return STEREO_ERR_INVALID_ARG; }
    const size_t px = type == PixType::F32 ? 4 : 1;
    if (left_step < cols * px || right_step < cols * px || disp_step < size_t(cols) * elem) { set_error("a step is smaller than its row"); return STEREO_ERR_INVALID_ARG; }
    if (R < 0 || range < 0) { set_error("window_rad and disparity_range must be >= 0"); return STEREO_ERR_INVALID_ARG; }
    int h0 = 0, h1 = 0;
    rc = stereo_band_halo_rows(rows, row_begin, row_end, R, &h0, &h1);
    if (rc != STEREO_OK) return rc;
    const int nh = h1 - h0, nb = row_end - row_begin;
    cudaStream_t st = ctx->stream;
    const size_t u8_pitch = align256(cols), d_pitch = align256(size_t(cols) * elem);
    const size_t need = 2 * u8_pitch * nh + 2 * d_pitch * nb + 1024;
    if (need > ctx->io.cap) {
        SB_CUDA(cudaStreamSynchronize(st));
        rc = ctx->io.reserve(need);
        if (rc != STEREO_OK) return rc;
    }
    ctx->io.reset();
    uint8_t* d_l = static_cast<uint8_t*>(ctx->io.take(u8_pitch * nh));
    uint8_t* d_r = static_cast<uint8_t*>(ctx->io.take(u8_pitch * nh));
    char* d_dl = static_cast<char*>(ctx->io.take(d_pitch * nb));
    char* d_dr = static_cast<char*>(ctx->io.take(d_pitch * nb));
    if (!d_l || !d_r || !d_dl || !d_dr) { set_error("io arena too small (internal)"); return STEREO_ERR_ALLOC; }
    const char* hl = static_cast<const char*>(left) + size_t(h0) * left_step;
    const char* hr = static_cast<const char*>(right) + size_t(h0) * right_step;
    if (type == PixType::F32) {
        const size_t stg_pitch = (size_t(cols) + 63) & ~size_t(63);
        SB_CUDA(cudaStreamSynchronize(st));                     // (the staging may still feed an earlier upload)
        rc = ensure_pinned(ctx, 2 * stg_pitch * nh);
        if (rc != STEREO_OK) return rc;
        if (ctx->host_threads == 0) ctx->host_threads = default_host_threads();
        const int nt = ctx->host_threads > 0 ? ctx->host_threads : 1;
        if (!ctx->pool || ctx->pool->threads() != nt) { delete ctx->pool; ctx->pool = new (std::nothrow) HostPool(nt); }
        if (!ctx->pool) { set_error("out of host memory"); return STEREO_ERR_ALLOC; }
        uint8_t* sl = static_cast<uint8_t*>(ctx->pinned);
        uint8_t* sr = sl + stg_pitch * nh;
        const bool ok = pack_f32_u8(*ctx->pool, reinterpret_cast<const float*>(hl), left_step, sl, stg_pitch, nh, cols) &&
                        pack_f32_u8(*ctx->pool, reinterpret_cast<const float*>(hr), right_step, sr, stg_pitch, nh, cols);
        if (!ok) { set_error("row-band entry points take 8-bit-valued images only"); return STEREO_ERR_UNSUPPORTED; }
        SB_CUDA(cudaMemcpy2DAsync(d_l, u8_pitch, sl, stg_pitch, cols, nh, cudaMemcpyHostToDevice, st));
        SB_CUDA(cudaMemcpy2DAsync(d_r, u8_pitch, sr, stg_pitch, cols, nh, cudaMemcpyHostToDevice, st));
    } else {
        SB_CUDA(cudaMemcpy2DAsync(d_l, u8_pitch, hl, left_step, cols, nh, cudaMemcpyHostToDevice, st));
        SB_CUDA(cudaMemcpy2DAsync(d_r, u8_pitch, hr, right_step, cols, nh, cudaMemcpyHostToDevice, st));
    }
    rc = stereo_disparity_pair_band_halo_u8_device(ctx, cost, d_l, u8_pitch, d_r, u8_pitch, rows, cols, row_begin, row_end, h0, h1, R, range,
                                                   d_dl, d_dr, d_pitch, elem, st);
    if (rc == STEREO_OK) {
        cudaError_t e = cudaMemcpy2DAsync(disp_left, disp_step, d_dl, d_pitch, size_t(cols) * elem, nb, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaMemcpy2DAsync(disp_right, disp_step, d_dr, d_pitch, size_t(cols) * elem, nb, cudaMemcpyDeviceToHost, st);
        if (e != cudaSuccess) { set_error("download failed: %s", cudaGetErrorString(e)); rc = STEREO_ERR_CUDA; }
    }
    cudaError_t e = cudaStreamSynchronize(st);                  // also on errors: nothing may still touch the caller's buffers
    if (rc == STEREO_OK && e != cudaSuccess) { set_error("band call failed: %s", cudaGetErrorString(e)); rc = STEREO_ERR_CUDA; }
    return rc;
}

int stereo_disparity_pair_band_u8_host(stereo_ctx* ctx, int cost, const uint8_t* left, size_t left_step, const uint8_t* right,
                                       size_t right_step, int rows, int cols, int row_begin, int row_end, int window_rad,
                                       int disparity_range, void* disp_left, void* disp_right, size_t disp_step, int disp_elem_bytes) {
    return pair_band_host(ctx, cost, PixType::U8, left, left_step, right, right_step, rows, cols, row_begin, row_end, window_rad,
                          disparity_range, disp_left, disp_right, disp_step, disp_elem_bytes);
}

int stereo_disparity_pair_band_f32_host(stereo_ctx* ctx, int cost, const float* left, size_t left_step, const float* right,
                                        size_t right_step, int rows, int cols, int row_begin, int row_end, int window_rad,
                                        int disparity_range, void* disp_left, void* disp_right, size_t disp_step, int disp_elem_bytes) {
    return pair_band_host(ctx, cost, PixType::F32, left, left_step, right, right_step, rows, cols, row_begin, row_end, window_rad,
                          disparity_range, disp_left, disp_right, disp_step, disp_elem_bytes);
}

int stereo_disparity_pair_batch_u8_device(stereo_ctx* ctx, int cost, int n_pairs, const uint8_t* left,
                                          const uint8_t* right, size_t img_step, size_t pair_stride, int rows,
                                          int cols, int window_rad, int disparity_range, void* disp_left,
                                          void* disp_right, size_t disp_step, size_t disp_pair_stride,
                                          int disp_elem_bytes, void* cuda_stream) {
    int rc = check_ctx(ctx);
    if (rc != STEREO_OK) return rc;
    if (n_pairs <= 0) { set_error("n_pairs must be positive"); return STEREO_ERR_INVALID_ARG; }
    if (!left || !right || !disp_left || !disp_right) { set_error("null pointer"); return STEREO_ERR_INVALID_ARG; }
    cudaStream_t st = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : ctx->stream;
    if (disparity_range < 0) { set_error("disparity_range must be >= 0"); return STEREO_ERR_INVALID_RANGE; }
    begin_call(ctx, st);
    // jobs of FMAXJOBS/2 pairs at a time: up to FMAXJOBS directions share one launch sequence
    constexpr int CHUNK = FMAXJOBS / 2;
    Problem ps[FMAXJOBS];
    for (int i0 = 0; i0 < n_pairs && rc == STEREO_OK; i0 += CHUNK) {
        const int np = (n_pairs - i0 < CHUNK) ? n_pairs - i0 : CHUNK;
        // left-referenced maps first, then right-referenced ones: jobs with the same range sign are neighbours
        for (int k = 0; k < np; ++k) {
            Problem two[2];
            const int i = i0 + k;
            pair_problems(two, cost, PixType::U8, left + size_t(i) * pair_stride, img_step, right + size_t(i) * pair_stride, img_step,
                          rows, cols, window_rad, disparity_range, static_cast<char*>(disp_left) + size_t(i) * disp_pair_stride,
                          static_cast<char*>(disp_right) + size_t(i) * disp_pair_stride, disp_step, disp_elem_bytes);
            ps[k] = two[0]; ps[np + k] = two[1];
        }
        if (all_fast(ctx, ps, 2 * np)) {
            for (int k = 0; k < 2 * np && rc == STEREO_OK; ++k) rc = validate(ps[k], size_t(cols), size_t(cols));
            if (rc == STEREO_OK) rc = run_fast_jobs(ctx, ps, 2 * np, st);
        } else {
            for (int k = 0; k < np && rc == STEREO_OK; ++k) {
                rc = run_problem(ctx, ps[k], st);
                if (rc == STEREO_OK) rc = run_problem(ctx, ps[np + k], st);
            }
        }
    }
    end_call(ctx, st);
    return rc;
}

// Host batches, both pixel types.  Strides are in bytes.
static int pair_batch_host(stereo_ctx* ctx, int cost, PixType type, int n_pairs, const void* left_v, const void* right_v,
                           size_t img_step, size_t pair_stride, int rows, int cols, int window_rad, int disparity_range,
                           void* disp_left, void* disp_right, size_t disp_step, size_t disp_pair_stride, int disp_elem_bytes) {
    int rc = check_ctx(ctx);
    if (rc != STEREO_OK) return rc;
    const char* left = static_cast<const char*>(left_v);
    const char* right = static_cast<const char*>(right_v);
    const size_t px = type == PixType::F32 ? 4 : 1;
    if (n_pairs <= 0) { set_error("n_pairs must be positive"); return STEREO_ERR_INVALID_ARG; }
    if (!left || !right || !disp_left || !disp_right) { set_error("null pointer"); return STEREO_ERR_INVALID_ARG; }
    if (rows <= 0 || cols <= 0 || rows > 32768 || cols > 32768 || img_step < size_t(cols) * px || disp_step < size_t(cols) * disp_elem_bytes) { set_error("bad size/step"); return STEREO_ERR_INVALID_ARG; }
    if (disp_elem_bytes != 1 && disp_elem_bytes != 2 && disp_elem_bytes != 4) { set_error("disp_elem_bytes must be 1, 2 or 4"); return STEREO_ERR_INVALID_ARG; }
    if (disparity_range < 0) { set_error("disparity_range must be >= 0"); return STEREO_ERR_INVALID_RANGE; }
    {   // pipelined: item i+1 uploads while item i computes and item i-1 downloads (3 device slots)
        std::vector<HostPairIn> in(n_pairs);
        std::vector<void*> outs_l(n_pairs), outs_r(n_pairs);
        for (int i = 0; i < n_pairs; ++i) {
            in[i] = HostPairIn{left + size_t(i) * pair_stride, img_step, right + size_t(i) * pair_stride, img_step};
            outs_l[i] = static_cast<char*>(disp_left) + size_t(i) * disp_pair_stride;
            outs_r[i] = static_cast<char*>(disp_right) + size_t(i) * disp_pair_stride;
        }
        const HostDir dirs[2] = {{false, -disparity_range, 0, outs_l.data()}, {true, 0, disparity_range, outs_r.data()}};
        rc = pairs_host_pipelined(ctx, cost, type, n_pairs, in.data(), dirs, 2, rows, cols, window_rad, disp_step, disp_elem_bytes);
        if (rc <= STEREO_OK) return rc;
    }
    // Sequential fallback (parameters the packed kernels do not cover, or float images that are not 8-bit):
    // pair by pair through the single-pair host path (which picks the exact kernels where needed).
    int launches = 0;
    for (int i = 0; i < n_pairs; ++i) {
        rc = pair_host(ctx, cost, type, left + size_t(i) * pair_stride, img_step, right + size_t(i) * pair_stride, img_step, rows, cols,
                       window_rad, disparity_range, static_cast<char*>(disp_left) + size_t(i) * disp_pair_stride,
                       static_cast<char*>(disp_right) + size_t(i) * disp_pair_stride, disp_step, disp_elem_bytes);
        if (rc != STEREO_OK) return rc;
        launches += ctx->last_launches;
    }
    ctx->last_launches = launches;
    return STEREO_OK;
}

int stereo_disparity_pair_batch_u8_host(stereo_ctx* ctx, int cost, int n_pairs, const uint8_t* left,
                                        const uint8_t* right, size_t img_step, size_t pair_stride, int rows, int cols,
                                        int window_rad, int disparity_range, void* disp_left, void* disp_right,
                                        size_t disp_step, size_t disp_pair_stride, int disp_elem_bytes) {
    return pair_batch_host(ctx, cost, PixType::U8, n_pairs, left, right, img_step, pair_stride, rows, cols, window_rad,
                           disparity_range, disp_left, disp_right, disp_step, disp_pair_stride, disp_elem_bytes);
}

int stereo_disparity_pair_batch_f32_host(stereo_ctx* ctx, int cost, int n_pairs, const float* left,
                                         const float* right, size_t img_step, size_t pair_stride, int rows, int cols,
                                         int window_rad, int disparity_range, void* disp_left, void* disp_right,
                                         size_t disp_step, size_t disp_pair_stride, int disp_elem_bytes) {
    return pair_batch_host(ctx, cost, PixType::F32, n_pairs, left, right, img_step, pair_stride, rows, cols, window_rad,
                           disparity_range, disp_left, disp_right, disp_step, disp_pair_stride, disp_elem_bytes);
}

} // extern "C"
