// Host-side float32 -> u8 image packing for the pipelined CV_32FC1 entry points (host_pack.cpp).
#pragma once
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

// min(16, hardware threads / LOCAL_WORLD_SIZE)
int default_host_threads();

} // namespace sb
