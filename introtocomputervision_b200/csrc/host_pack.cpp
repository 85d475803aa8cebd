// Host side of the pipelined CV_32FC1 entry points: a small persistent thread pool that converts float32 images to the
// u8 images the packed kernels consume (checking on the way that every pixel really is an integer in 0..255), so that
// the host link carries 1 byte per pixel instead of 4.  The reference uploads the CV_32FC1 Mats as they are
// (DisparitySSD.cu:171-174).  Plain C++ (compiled by the host compiler, no CUDA in here).
#include "host_pack.hpp"

#include <atomic>
#include <chrono>
#include <cstdint>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

namespace sb {

// One row segment; returns false as soon as the segment holds a pixel that is not an integer in 0..255 (NaN included:
// (int)NaN converts to INT_MIN, which fails both tests).  target_clones: resolved at load time for the CPU the library runs on.
#if defined(__x86_64__) && defined(__GNUC__) && !defined(__clang__)
__attribute__((target_clones("avx512f", "avx2", "default")))
#endif
bool pack_f32_u8_row(const float* __restrict__ src, uint8_t* __restrict__ dst, int n) {
    int bad = 0;
    for (int x = 0; x < n; ++x) {
        const float f = src[x];
        const int v = int(f);
        bad |= (float(v) != f) | (unsigned(v) > 255u);
        dst[x] = uint8_t(v);
    }
    return bad == 0;
}

#if defined(__x86_64__) && defined(__GNUC__) && !defined(__clang__)
#include <immintrin.h>
// AVX-512 row: 64 pixels per iteration, non-temporal stores into the (64-byte aligned) pinned staging - the staged bytes are
// read next by the DMA engine, not by a core, so they should not displace the source image from the caches nor cost a
// read-for-ownership of the destination lines.
__attribute__((target("avx512f,avx512bw")))
static bool pack_row_avx512(const float* __restrict__ src, uint8_t* __restrict__ dst, int n) {
    __m512i bad = _mm512_setzero_si512();
    __mmask16 ne = 0;
    int x = 0;
    const bool aligned = (reinterpret_cast<uintptr_t>(dst) & 63) == 0;
    for (; x + 64 <= n; x += 64) {
        const __m512 f0 = _mm512_loadu_ps(src + x), f1 = _mm512_loadu_ps(src + x + 16), f2 = _mm512_loadu_ps(src + x + 32), f3 = _mm512_loadu_ps(src + x + 48);
        const __m512i i0 = _mm512_cvttps_epi32(f0), i1 = _mm512_cvttps_epi32(f1), i2 = _mm512_cvttps_epi32(f2), i3 = _mm512_cvttps_epi32(f3);
        ne |= _mm512_cmp_ps_mask(_mm512_cvtepi32_ps(i0), f0, _CMP_NEQ_UQ) | _mm512_cmp_ps_mask(_mm512_cvtepi32_ps(i1), f1, _CMP_NEQ_UQ) |
              _mm512_cmp_ps_mask(_mm512_cvtepi32_ps(i2), f2, _CMP_NEQ_UQ) | _mm512_cmp_ps_mask(_mm512_cvtepi32_ps(i3), f3, _CMP_NEQ_UQ);
        bad = _mm512_or_si512(bad, _mm512_or_si512(_mm512_or_si512(i0, i1), _mm512_or_si512(i2, i3)));
        const __m128i b0 = _mm512_cvtepi32_epi8(i0), b1 = _mm512_cvtepi32_epi8(i1), b2 = _mm512_cvtepi32_epi8(i2), b3 = _mm512_cvtepi32_epi8(i3);
        __m512i out = _mm512_castsi128_si512(b0);
        out = _mm512_inserti32x4(out, b1, 1);
        out = _mm512_inserti32x4(out, b2, 2);
        out = _mm512_inserti32x4(out, b3, 3);
        if (aligned) _mm512_stream_si512(reinterpret_cast<__m512i*>(dst + x), out);
        else _mm512_storeu_si512(dst + x, out);
    }
    // any bit above the low 8 set in any lane <=> some value outside 0..255 (negative values have the sign bits set)
    int oob = _mm512_test_epi32_mask(bad, _mm512_set1_epi32(~0xFF)) != 0;
    for (; x < n; ++x) {
        const float f = src[x];
        const int v = int(f);
        oob |= (float(v) != f) | (unsigned(v) > 255u);
        dst[x] = uint8_t(v);
    }
    return !oob && ne == 0;
}
static const bool g_have_avx512 = __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw") && !getenv("STEREO_NO_AVX512");
#else
static const bool g_have_avx512 = false;
static bool pack_row_avx512(const float*, uint8_t*, int) { return false; }
#endif

// Workers and the caller poll for a bounded time before they block: inside one pipelined call the conversions follow each
// other within microseconds (one per work item), and a condition-variable wake-up of 15 threads costs about as much as
// converting a quarter of a 4K band.  After HOST_POOL_SPIN_US without work a worker sleeps on the condition variable, so an
// idle context costs nothing.
constexpr int HOST_POOL_SPIN_US = 250;

static inline void cpu_relax() {
#if defined(__x86_64__)
    __builtin_ia32_pause();
#else
    std::this_thread::yield();
#endif
}

struct HostPool::Impl {
    std::vector<std::thread> workers;
    std::mutex mu;
    std::condition_variable cv_work;
    const std::function<void(int)>* fn = nullptr;
    std::function<void(int)> job;       // the dispatch in flight (begin() .. end())
    bool open = false, inline_only = false;
    int n_tasks = 0;
    std::atomic<int> next{0};
    std::atomic<int> generation{0};     // bumped by run() after fn / n_tasks / next / pending are in place
    std::atomic<int> pending{0};        // workers that have not finished the current generation yet
    std::atomic<bool> stop{false};

    void loop() {
        int seen = 0;
        for (;;) {
            auto t0 = std::chrono::steady_clock::now();
            int g, polls = 0;
            while ((g = generation.load(std::memory_order_acquire)) == seen && !stop.load(std::memory_order_relaxed)) {
                cpu_relax();
                if ((++polls & 63) == 0 &&
                    std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - t0).count() > HOST_POOL_SPIN_US) {
                    std::unique_lock<std::mutex> lk(mu);
                    cv_work.wait(lk, [&] { return stop.load() || generation.load() != seen; });
                    t0 = std::chrono::steady_clock::now();
                }
            }
            if (stop.load()) return;
            seen = g;
            const std::function<void(int)>* f = fn;
            const int n = n_tasks;
            for (int i; (i = next.fetch_add(1, std::memory_order_relaxed)) < n;) (*f)(i);
            pending.fetch_sub(1, std::memory_order_release);
        }
    }
};

HostPool::HostPool(int threads) : impl_(new Impl), threads_(threads < 1 ? 1 : threads) {
    for (int i = 1; i < threads_; ++i) impl_->workers.emplace_back([this] { impl_->loop(); });
}

HostPool::~HostPool() {
    end();
    {
        std::lock_guard<std::mutex> lk(impl_->mu);
        impl_->stop.store(true);
    }
    impl_->cv_work.notify_all();
    for (auto& t : impl_->workers) t.join();
    delete impl_;
}

void HostPool::run(int n_tasks, const std::function<void(int)>& fn) {
    begin(n_tasks, fn);
    end();
}

// begin() publishes a dispatch and returns: the workers start on it while the caller does something else (the host
// pipeline enqueues the previous chunk's copies and launches); end() lets the caller take whatever tasks are left and
// returns when every task has finished.  One dispatch at a time; `fn` is copied.
void HostPool::begin(int n_tasks, const std::function<void(int)>& fn) {
    impl_->job = fn;
    impl_->n_tasks = n_tasks < 0 ? 0 : n_tasks;
    impl_->open = true;
    if (threads_ == 1 || impl_->n_tasks <= 1) { impl_->inline_only = true; return; }      // end() runs it on the caller
    impl_->inline_only = false;
    impl_->fn = &impl_->job;
    impl_->next.store(0, std::memory_order_relaxed);
    impl_->pending.store(int(impl_->workers.size()), std::memory_order_relaxed);
    impl_->generation.fetch_add(1, std::memory_order_seq_cst);
    { std::lock_guard<std::mutex> lk(impl_->mu); }          // orders the bump against a worker that is about to sleep
    impl_->cv_work.notify_all();
}

void HostPool::end() {
    if (!impl_->open) return;
    impl_->open = false;
    const int n_tasks = impl_->n_tasks;
    if (impl_->inline_only) { for (int i = 0; i < n_tasks; ++i) impl_->job(i); return; }
    for (int i; (i = impl_->next.fetch_add(1, std::memory_order_relaxed)) < n_tasks;) impl_->job(i);       // the calling thread works too
    // every worker has to pass through this generation (it reads fn / n_tasks) before the next begin() may change them
    for (int polls = 0; impl_->pending.load(std::memory_order_acquire) != 0;) {
        cpu_relax();
        if ((++polls & 1023) == 0) std::this_thread::yield();
    }
}

// tasks of ~64K pixels: enough of them to balance, few enough to keep the dispatch cost invisible; the tasks of all images
// of one upload chunk (left and right of every pair riding it) go out in ONE dispatch
void PackAsync::begin(HostPool& pool, const PackJob* jobs, int n_jobs) {
    n_ = n_jobs < 0 ? 0 : (n_jobs > MAXJ ? MAXJ : n_jobs);
    first_[0] = 0;
    for (int j = 0; j < n_; ++j) {
        jobs_[j] = jobs[j];
        rpt_[j] = (1 << 16) / (jobs[j].cols > 0 ? jobs[j].cols : 1);
        if (rpt_[j] < 1) rpt_[j] = 1;
        first_[j + 1] = first_[j] + (jobs[j].rows > 0 ? (jobs[j].rows + rpt_[j] - 1) / rpt_[j] : 0);
    }
    bad_.store(0, std::memory_order_relaxed);
    pool.begin(first_[n_], [this](int t) {
        int j = 0;
        while (t >= first_[j + 1]) ++j;
        const PackJob& J = jobs_[j];
        const int r0 = (t - first_[j]) * rpt_[j], r1 = r0 + rpt_[j] < J.rows ? r0 + rpt_[j] : J.rows;
        bool ok = true;
        for (int r = r0; r < r1; ++r) {
            const float* s = reinterpret_cast<const float*>(reinterpret_cast<const char*>(J.src) + size_t(r) * J.src_step);
            ok &= g_have_avx512 ? pack_row_avx512(s, J.dst + size_t(r) * J.dst_step, J.cols) : pack_f32_u8_row(s, J.dst + size_t(r) * J.dst_step, J.cols);
        }
#if defined(__x86_64__) && defined(__GNUC__)
        if (g_have_avx512) __builtin_ia32_sfence();       // the non-temporal stores are visible before the upload is enqueued
#endif
        if (!ok) bad_.store(1, std::memory_order_relaxed);
    });
}

bool PackAsync::end(HostPool& pool) {
    pool.end();
    return bad_.load() == 0;
}

bool pack_f32_u8_jobs(HostPool& pool, const PackJob* jobs, int n_jobs) {
    bool ok = true;
    PackAsync pa;
    for (int j = 0; j < n_jobs; j += PackAsync::MAXJ) {
        pa.begin(pool, jobs + j, n_jobs - j < PackAsync::MAXJ ? n_jobs - j : PackAsync::MAXJ);
        ok = pa.end(pool) && ok;
    }
    return ok;
}

bool pack_f32_u8(HostPool& pool, const float* src, size_t src_step, uint8_t* dst, size_t dst_step, int rows, int cols) {
    const PackJob job{src, src_step, dst, dst_step, rows, cols};
    return pack_f32_u8_jobs(pool, &job, 1);
}

int default_host_threads() {
    unsigned hw = std::thread::hardware_concurrency();
    if (hw == 0) hw = 1;
    // one process per GPU under a torchrun-style launcher: share the cores
    if (const char* e = getenv("LOCAL_WORLD_SIZE")) { const int w = atoi(e); if (w > 1) hw = hw / unsigned(w) ? hw / unsigned(w) : 1; }
    return int(hw > 16 ? 16 : hw);
}

} // namespace sb
