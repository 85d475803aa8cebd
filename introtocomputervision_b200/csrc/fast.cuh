// Packed u8 path (placeholder until the kernels land): reports "not supported" so that every call
// takes the exact path.
#pragma once
#include "common.cuh"
namespace sb {
static inline bool fast_supported(const Problem&) { return false; }
static inline size_t fast_scratch_bytes(stereo_ctx*, const Problem&) { return 0; }
static inline int fast_ctx_init(stereo_ctx*) { return STEREO_OK; }
static inline int run_fast(stereo_ctx*, const Problem&, cudaStream_t) { set_error("fast path not built"); return STEREO_ERR_UNSUPPORTED; }
}
