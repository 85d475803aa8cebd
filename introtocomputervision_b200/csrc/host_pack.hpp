// Host-side float32 -> u8 image packing for the pipelined CV_32FC1 entry points (host_pack.cpp).
#pragma once
#include <atomic>
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <functional>

namespace sb {

class HostPool {       // persistent worker threads; run() returns when every task has finished
public:
    explicit HostPool(int threads);
    ~HostPool();
    HostPool(const HostPool&) = delete;
    HostPool& operator=(const HostPool&) = delete;
    void run(int n_tasks, const std::function<void(int)>& fn);
    // asynchronous form: begin() hands the tasks to the workers and returns, end() joins in and waits (one dispatch at a time)
    void begin(int n_tasks, const std::function<void(int)>& fn);
    void end();
    int threads() const { return threads_; }
private:
    struct Impl;
    Impl* impl_;
    int threads_;
};

// rows x cols float32 (row stride src_step bytes) -> u8 (row stride dst_step bytes).  false: some pixel is not an integer
// in 0..255 (the u8 image is then meaningless).
bool pack_f32_u8(HostPool& pool, const float* src, size_t src_step, uint8_t* dst, size_t dst_step, int rows, int cols);
// several images in one dispatch of the pool (the images of one work item of the host pipeline)
struct PackJob { const float* src; size_t src_step; uint8_t* dst; size_t dst_step; int rows, cols; };
bool pack_f32_u8_jobs(HostPool& pool, const PackJob* jobs, int n_jobs);
// The same as a dispatch in flight: begin() returns while the workers convert, end() joins in, waits and reports whether
// every pixel was 8-bit.  The object (it holds the job list) must stay alive and unmoved in between; up to MAXJ images.
class PackAsync {
public:
    static constexpr int MAXJ = 16;
    void begin(HostPool& pool, const PackJob* jobs, int n_jobs);
    bool end(HostPool& pool);
private:
    PackJob jobs_[MAXJ];
    int first_[MAXJ + 1], rpt_[MAXJ], n_ = 0;
    std::atomic<int> bad_{0};
};

// min(16, hardware threads / LOCAL_WORLD_SIZE)
int default_host_threads();

} // namespace sb
