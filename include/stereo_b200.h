/*
 * stereo_b200.h — C ABI of the B200-native (sm_100a) dense stereo block matcher.
 *
 * This is the drop-in boundary for ONE path of tanmaniac/IntroToComputerVision: the Problem Set 2
 * window-based stereo matcher.  Every entry point cites the reference interface it replaces; paths
 * are relative to the reference repository root.
 *
 *   cuda::disparitySSD     ProblemSets/ps2_cpp/include/DisparitySSD.h:18-23
 *                          (wrapper being replaced: ProblemSets/ps2_cpp/lib/DisparitySSD.cu:143-207)
 *   cuda::disparityNCorr   ProblemSets/ps2_cpp/include/DisparityNCorr.h:19-24
 *                          (wrapper being replaced: ProblemSets/ps2_cpp/lib/DisparityNCorr.cu:177-251)
 *   disparitySSDPair / disparityNCorrPair   ProblemSets/ps2_cpp/src/main.cpp:21-48, 51-78
 *
 * RESULTS follow the reference's CPU semantics (serial::disparitySSD, lib/DisparitySSD.cpp:9-62 and
 * serial::disparityNCorr, lib/DisparityNCorr.cpp:12-71) — the reference's GPU kernels compute a
 * different window and thresholds and are not the parity target.
 *
 * Conventions
 *   - "ref" is the reference image (first argument of the reference functions), "tgt" the image
 *     that is searched.  L->R maps pass (left, right, -range, 0); R->L maps pass (right, left, 0,
 *     +range) — main.cpp:33,43.
 *   - Images are single channel, row-major; `*_step` arguments are row strides in BYTES
 *     (cv::Mat::step).  f32 images carry raw 0..255 intensities (no /255 — main.cpp:87-88) and may
 *     be non-integer/negative (main.cpp:140-153); u8 entry points are the same computation for
 *     images that are exactly 8-bit.
 *   - Disparity output element size is 1 (int8, the reference's CV_8SC1 with its `char` narrowing,
 *     DisparitySSD.cpp:59), 2 (int16) or 4 (int32); 2/4 hold the un-narrowed value, which is what
 *     >127-disparity searches need.
 *   - Every function returns STEREO_OK (0) or a negative stereo_status; nothing in this library
 *     calls exit() (the reference does: common/include/common/CudaCommon.cuh:11-22).
 *     stereo_last_error() returns a per-thread description of the last failure.
 *   - There is no CPU fallback: without a CUDA device every compute entry point fails with
 *     STEREO_ERR_NO_DEVICE.
 *   - A context owns all device buffers, streams and scratch; it is created once and reused (the
 *     reference re-allocates per call, DisparitySSD.cu:171-178).  A context is not thread-safe;
 *     use one per host thread.  Different contexts are independent (the reference's file-scope
 *     texture references, DisparitySSD.cu:19-20, made it non-reentrant).
 *   - Calls on one context execute in call order even when they are enqueued on different CUDA
 *     streams: all of them share the context's scratch arena, so a call enqueued on another stream
 *     than its predecessor first waits (on the device) for the predecessor's last kernel.  Use one
 *     context per stream for concurrent execution.
 */
#ifndef STEREO_B200_H_
#define STEREO_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define STEREO_B200_ABI_VERSION 1

typedef enum stereo_status {
    STEREO_OK = 0,
    STEREO_ERR_INVALID_ARG = -1,   /* null pointer, non-positive size, bad element size, step too small */
    STEREO_ERR_INVALID_RANGE = -2, /* min_disp > max_disp, or an NCC range that leaves a pixel without candidates */
    STEREO_ERR_NO_DEVICE = -3,     /* no CUDA device / wrong architecture (needs sm_100) */
    STEREO_ERR_CUDA = -4,          /* a CUDA runtime/driver call failed; see stereo_last_error() */
    STEREO_ERR_ALLOC = -5,         /* device or pinned-host allocation failed */
    STEREO_ERR_UNSUPPORTED = -6    /* parameter combination outside what the kernels cover */
} stereo_status;

typedef enum stereo_cost {
    STEREO_COST_SSD = 0,   /* sum of squared differences, argmin, first minimum wins   (DisparitySSD.cpp:45-57) */
    STEREO_COST_NCORR = 1  /* TM_CCORR_NORMED, argmax, first maximum wins              (DisparityNCorr.cpp:60-67) */
} stereo_cost;

/* Which kernel family served the last call (for tests and reports). */
typedef enum stereo_path {
    STEREO_PATH_NONE = 0,
    STEREO_PATH_EXACT_F32 = 1,  /* any float input, window and range: per-element reference arithmetic, no running sums */
    STEREO_PATH_FAST_U8 = 2,    /* integer-valued 0..255 inputs: packed dot-product running sums */
    STEREO_PATH_REFGPU = 4,     /* stereo_disparity_refgpu_*: the function the reference's GPU kernels compute */
    STEREO_PATH_FAST_F32 = 3    /* general float inputs of bounded range (noise / contrast variants, main.cpp:140-153,191-193):
                                   per-element round((l-r)^2) as exact int32 running sums (SSD), float32 running sums (NCC) */
} stereo_path;

typedef struct stereo_ctx stereo_ctx;

/* ---- library / context ------------------------------------------------------------------- */

int stereo_abi_version(void);
const char* stereo_last_error(void);
const char* stereo_status_string(int status);

/* Number of visible CUDA devices of compute capability 10.x (0 if none / no driver). */
int stereo_device_count(void);

/* Creates a context on `device` (ordinal).  Replaces the per-call GpuMat/Stream setup of
 * DisparitySSD.cu:163-182 and common::warmup() (common/src/CudaWarmup.cu:14-19). */
int stereo_ctx_create(int device, stereo_ctx** ctx_out);
void stereo_ctx_destroy(stereo_ctx* ctx);

/* Kernel family used by the most recent compute call on this context (stereo_path). */
int stereo_ctx_last_path(const stereo_ctx* ctx);
/* Device time of the most recent compute call's kernels in milliseconds (cudaEvent pair; the
 * reference logs the same quantity, DisparitySSD.cu:192-203).  <0 if unavailable.  For the pipelined
 * HOST entry points the pair brackets the compute stream, which waits for every band's upload: the
 * figure then includes upload time the kernels were blocked on (stereo_ctx_last_hot_kernel_ms is
 * kernels only).  CV_32FC1 DEVICE calls: the time the device idles while the host reads the classification
 * verdict is left out (classification kernels + the kernels of the chosen family). */
float stereo_ctx_last_kernel_ms(const stereo_ctx* ctx);
/* Device time (ms) of the most recent call's HOT kernels only (the packed cost/WTA kernels; one launch
 * covers up to 32 directions of equally shaped problems), summed over the launches that were measured
 * (at most 16 per call); *launches_measured (nullable) receives how many that was.  <0 if the call used
 * no hot kernel.  This is the number the roofline report divides by. */
float stereo_ctx_last_hot_kernel_ms(const stereo_ctx* ctx, int* launches_measured);
/* Directions (disparity maps) the measured hot launches of the most recent call computed. */
int stereo_ctx_last_hot_jobs(const stereo_ctx* ctx);
/* Number of kernel launches issued by the most recent compute call. */
int stereo_ctx_last_launches(const stereo_ctx* ctx);
/* Force a kernel family (debug/testing): 0 = automatic, else a stereo_path value. */
int stereo_ctx_force_path(stereo_ctx* ctx, int path);
/* Host entry points overlap upload, compute and download by cutting each image pair into row bands
 * (with the window halo) that flow through three streams.  bands = 0 (default) picks the band count
 * from the image size and the number of pairs; bands >= 1 forces it (1 = whole images).  Results do not
 * depend on it.  The reference uploads, computes and downloads one after the other
 * (DisparitySSD.cu:171-206). */
int stereo_ctx_set_pipe_bands(stereo_ctx* ctx, int bands);

/* How the pipelined host entry points cut a call of `n_pairs` rows x cols pairs into work items: row bands per pair
 * (bands_override = stereo_ctx_set_pipe_bands' value, 0 = automatic) and pairs per item (small images ride several
 * pairs per launch sequence).  Pure host arithmetic, no device needed. */
int stereo_host_pipeline_plan(int n_pairs, int rows, int cols, int bands_override, int* bands_per_pair, int* pairs_per_item);
/* The row bands of work item `item` of such a call: writes the ascending boundaries (first 0, last rows) to bounds[0 .. n-1]
 * and returns n (negative: error; cap = room in bounds).  With automatic bands, 4K-sized images are cut unevenly: the
 * first item of a call begins and the last item ends with a band of an eighth of the image - the first upload and the last
 * download are the only copies no kernel overlaps - and the items in between go as whole images. */
int stereo_host_pipeline_item_bands(int n_pairs, int rows, int cols, int bands_override, int item, int* bounds, int cap);

/* CV_32FC1 HOST entry points: images whose pixels are all integers in 0..255 (everything convertTo(CV_32FC1) produces,
 * main.cpp:87-88) are converted to u8 by `threads` host threads into pinned staging and uploaded as 1 byte per pixel; the
 * reference uploads the float Mats (DisparitySSD.cu:171-174).  threads = 0 (default): min(16, hardware threads /
 * LOCAL_WORLD_SIZE), and host conversion only from 8 threads up (below that the floats are uploaded and converted on the
 * device); threads = -1: never convert on the host; threads >= 1: that many, always.  Results do not depend on it. */
int stereo_ctx_set_host_threads(stereo_ctx* ctx, int threads);
int stereo_ctx_host_threads(const stereo_ctx* ctx);
/* The conversion itself, on `threads` host threads (pure host code, no device needed): rows x cols float32 -> u8;
 * *all_8bit = 1 when every pixel is an integer in 0..255 (otherwise the u8 image is meaningless). */
int stereo_host_pack_f32_u8(const float* src, size_t src_step, uint8_t* dst, size_t dst_step, int rows, int cols, int threads,
                            int* all_8bit);
/* Self-test of the worker pool behind the conversion (pure host code): `rounds` dispatches of 1..257 tasks on ONE pool of
 * `threads` threads (blocking and begin / end form alternating), with pauses long enough for the polling workers to fall asleep; STEREO_OK when every task of every
 * dispatch ran exactly once (STEREO_ERR_UNSUPPORTED with the offending task in stereo_last_error() otherwise). */
int stereo_host_pool_selftest(int threads, int rounds);

/* How the hot kernel of a batch of `n_pairs` (<= 16) equally shaped pair problems would be launched on a device with
 * `sm_count` SMs (pure host arithmetic, no device needed; tests pin the scheduling decisions with it).
 * schedule: 0 = linear split of the (tile, row) space, 1 = equal row segments per tile (fewer tiles than SMs),
 * 2 = row-band-major items strided over the CTAs (many tiles: neighbouring tiles walk the same rows together). */
typedef struct stereo_launch_plan_t {
    int fused;            /* both maps from one cost volume */
    int strip_px;         /* pixels per thread (K) */
    int strips_per_warp;  /* 1, or 2 for searches of at most 64 candidates */
    int groups;           /* 128- (64-) disparity groups */
    int tile_px;          /* pixels per CTA tile */
    int tiles;            /* tiles of the launch */
    int ctas;             /* grid size */
    int schedule;
    int rows_per_item;    /* rows per CTA share (schedule 0, 1) or per band (2) */
    int bands;            /* schedule 2: row bands per tile */
    int stages;           /* TMA pipeline stages */
    int smem_bytes;
    int border_kernel;    /* fused SSD: candidates centred in the right padding come from fused_border_kernel */
} stereo_launch_plan_t;
int stereo_launch_plan(int cost, int float_operands, int n_pairs, int rows, int cols, int window_rad, int disparity_range, int fuse,
                       int sm_count, stereo_launch_plan_t* plan);

/* Pair calls (stereo_disparity_pair_*): compute BOTH maps of a pair from one cost volume where the kernels allow it -
 * SSD with any disparity_range and window_rad <= 7 on 8-bit-valued or float images, NCC on 8-bit-valued images - the
 * reference's disparitySSDPair / disparityNCorrPair always want both (main.cpp:21-78) and the window cost of (x, d) in
 * one direction IS the window cost of (x + d, -d) in the other.  SSD results are identical either way (bit-exact), NCC
 * results agree like any two evaluations of the same scores (per-pixel instead of per-strip key scales); on (the default)
 * is faster.  0 switches back to one cost volume per direction. */
int stereo_ctx_set_fuse_pairs(stereo_ctx* ctx, int on);
/* Image pairs of the last call whose two maps came out of one cost volume. */
int stereo_ctx_last_fused_pairs(const stereo_ctx* ctx);

/* ---- single direction, HOST buffers (the drop-in form) ------------------------------------ */
/*
 * One disparity map for reference image `ref` searched in `tgt` over [min_disp, max_disp].
 * Replaces cuda::disparitySSD / cuda::disparityNCorr (host cv::Mat in, host cv::Mat out;
 * upload, kernels and download happen inside, synchronously — DisparitySSD.cu:171-206).
 *   disp_out        rows x cols elements of disp_elem_bytes (1, 2 or 4), row stride disp_step bytes
 *   best_out        optional (NULL to skip) rows x cols winning values, row stride best_step bytes:
 *                   SSD  -> int32 cost  (99999999 where no candidate exists, DisparitySSD.cpp:37)
 *                   NCC  -> float32 score (the result[] value minMaxLoc selected, DisparityNCorr.cpp:60-64)
 */
int stereo_disparity_f32_host(stereo_ctx* ctx, int cost,
                              const float* ref, size_t ref_step, const float* tgt, size_t tgt_step,
                              int rows, int cols, int window_rad, int min_disp, int max_disp,
                              void* disp_out, size_t disp_step, int disp_elem_bytes,
                              void* best_out, size_t best_step);

int stereo_disparity_u8_host(stereo_ctx* ctx, int cost,
                             const uint8_t* ref, size_t ref_step, const uint8_t* tgt, size_t tgt_step,
                             int rows, int cols, int window_rad, int min_disp, int max_disp,
                             void* disp_out, size_t disp_step, int disp_elem_bytes,
                             void* best_out, size_t best_step);

/* ---- single direction, DEVICE buffers, asynchronous on `cuda_stream` ----------------------- */
/* Same computation with every pointer a device pointer on the context's device.  `cuda_stream` is a
 * cudaStream_t passed as void* (NULL = the context's own stream; pass cudaStreamLegacy, (void*)1, for the
 * CUDA legacy default stream).  Returns after enqueueing. */
int stereo_disparity_f32_device(stereo_ctx* ctx, int cost,
                                const float* ref, size_t ref_step, const float* tgt, size_t tgt_step,
                                int rows, int cols, int window_rad, int min_disp, int max_disp,
                                void* disp_out, size_t disp_step, int disp_elem_bytes,
                                void* best_out, size_t best_step, void* cuda_stream);

int stereo_disparity_u8_device(stereo_ctx* ctx, int cost,
                               const uint8_t* ref, size_t ref_step, const uint8_t* tgt, size_t tgt_step,
                               int rows, int cols, int window_rad, int min_disp, int max_disp,
                               void* disp_out, size_t disp_step, int disp_elem_bytes,
                               void* best_out, size_t best_step, void* cuda_stream);

/* ---- both directions (the reference's *Pair helpers, main.cpp:21-78) ------------------------ */
/* left-referenced map over [-disparity_range, 0] into disp_left, right-referenced map over
 * [0, +disparity_range] (images swapped) into disp_right.  Host buffers, synchronous. */
int stereo_disparity_pair_f32_host(stereo_ctx* ctx, int cost,
                                   const float* left, size_t left_step, const float* right, size_t right_step,
                                   int rows, int cols, int window_rad, int disparity_range,
                                   void* disp_left, void* disp_right, size_t disp_step, int disp_elem_bytes);

int stereo_disparity_pair_u8_host(stereo_ctx* ctx, int cost,
                                  const uint8_t* left, size_t left_step, const uint8_t* right, size_t right_step,
                                  int rows, int cols, int window_rad, int disparity_range,
                                  void* disp_left, void* disp_right, size_t disp_step, int disp_elem_bytes);

/* Device buffers, asynchronous on `cuda_stream`. */
int stereo_disparity_pair_u8_device(stereo_ctx* ctx, int cost,
                                    const uint8_t* left, size_t left_step, const uint8_t* right, size_t right_step,
                                    int rows, int cols, int window_rad, int disparity_range,
                                    void* disp_left, void* disp_right, size_t disp_step, int disp_elem_bytes,
                                    void* cuda_stream);

/* The same for CV_32FC1 device images (what the reference's wrapper holds after its uploads, DisparitySSD.cu:171-174).
 * The kernel family depends on the pixel values (8-bit-valued, bounded float range, anything else), which the
 * host learns from a device-side classification: this call synchronises `cuda_stream` once before it enqueues
 * the matching kernels; the kernels themselves run asynchronously. */
int stereo_disparity_pair_f32_device(stereo_ctx* ctx, int cost,
                                     const float* left, size_t left_step, const float* right, size_t right_step,
                                     int rows, int cols, int window_rad, int disparity_range,
                                     void* disp_left, void* disp_right, size_t disp_step, int disp_elem_bytes,
                                     void* cuda_stream);

/* ---- batches of equally-shaped pairs (BASELINE config 5) ------------------------------------ */
/* `n_pairs` image pairs stored back to back: pair i's left image starts at left + i*pair_stride
 * bytes (same for right); outputs start at disp_* + i*disp_pair_stride bytes.  Device buffers,
 * asynchronous on `cuda_stream`. */
int stereo_disparity_pair_batch_u8_device(stereo_ctx* ctx, int cost, int n_pairs,
                                          const uint8_t* left, const uint8_t* right, size_t img_step,
                                          size_t pair_stride, int rows, int cols, int window_rad,
                                          int disparity_range, void* disp_left, void* disp_right,
                                          size_t disp_step, size_t disp_pair_stride, int disp_elem_bytes,
                                          void* cuda_stream);

/* Host buffers, synchronous: uploads, computes and downloads with copies overlapped across pairs. */
int stereo_disparity_pair_batch_u8_host(stereo_ctx* ctx, int cost, int n_pairs,
                                        const uint8_t* left, const uint8_t* right, size_t img_step,
                                        size_t pair_stride, int rows, int cols, int window_rad,
                                        int disparity_range, void* disp_left, void* disp_right,
                                        size_t disp_step, size_t disp_pair_stride, int disp_elem_bytes);

/* The same for CV_32FC1 host images (the type the reference's entry points take, DisparitySSD.cu:150): a batch
 * of the reference's disparitySSDPair / disparityNCorrPair calls (main.cpp:21-78) in one call.  img_step and
 * pair_stride are in BYTES.  8-bit-valued images ride the packed kernels with uploads, conversion, compute and
 * downloads overlapped across pairs; anything else is computed pair by pair on the exact path. */
int stereo_disparity_pair_batch_f32_host(stereo_ctx* ctx, int cost, int n_pairs,
                                         const float* left, const float* right, size_t img_step,
                                         size_t pair_stride, int rows, int cols, int window_rad,
                                         int disparity_range, void* disp_left, void* disp_right,
                                         size_t disp_step, size_t disp_pair_stride, int disp_elem_bytes);

/* ---- row-band form for sharding one large image across GPUs (BASELINE config 4) -------------- */
/* Computes output rows [row_begin, row_end) of the full rows x cols problem.  `ref`/`tgt` point at
 * FULL images (device); only rows within the band's halo are read.  disp_out points at the band's
 * first output row.  Results equal the corresponding rows of the full-image call, including the
 * reference's row-wrap reads at band seams. */
int stereo_disparity_band_u8_device(stereo_ctx* ctx, int cost,
                                    const uint8_t* ref, size_t ref_step, const uint8_t* tgt, size_t tgt_step,
                                    int rows, int cols, int row_begin, int row_end,
                                    int window_rad, int min_disp, int max_disp,
                                    void* disp_out, size_t disp_step, int disp_elem_bytes, void* cuda_stream);

/* Same, for callers that hold only a slab of the images (one rank of a row-band sharded job): `ref_halo`/
 * `tgt_halo` point at image row `halo_begin`; rows [halo_begin, halo_end) are present.  The slab must
 * contain the rows stereo_band_halo_rows() reports for the band — window_rad rows above and below plus
 * one more for the reference's row-wrap reads, clamped to the image — then the result equals the
 * corresponding rows of the full-image call bit for bit. */
int stereo_disparity_band_halo_u8_device(stereo_ctx* ctx, int cost,
                                         const uint8_t* ref_halo, size_t ref_step, const uint8_t* tgt_halo, size_t tgt_step,
                                         int rows, int cols, int row_begin, int row_end, int halo_begin, int halo_end,
                                         int window_rad, int min_disp, int max_disp,
                                         void* disp_out, size_t disp_step, int disp_elem_bytes, void* cuda_stream);

/* Both maps of the band in one call (a rank's share of a row-band sharded pair): left-referenced map over
 * [-disparity_range, 0] into disp_left, right-referenced map over [0, +disparity_range] into disp_right, each
 * pointing at the band's first output row; from one cost volume where stereo_ctx_set_fuse_pairs allows it. */
int stereo_disparity_pair_band_halo_u8_device(stereo_ctx* ctx, int cost,
                                              const uint8_t* left_halo, size_t left_step, const uint8_t* right_halo, size_t right_step,
                                              int rows, int cols, int row_begin, int row_end, int halo_begin, int halo_end,
                                              int window_rad, int disparity_range,
                                              void* disp_left, void* disp_right, size_t disp_step, int disp_elem_bytes,
                                              void* cuda_stream);

/* Input rows [*halo_begin, *halo_end) a band [row_begin, row_end) of a rows-high image needs (pure host
 * arithmetic, no device). */
int stereo_band_halo_rows(int rows, int row_begin, int row_end, int window_rad, int* halo_begin, int* halo_end);

/* ---- peer gather: the exchange step of the sharded configs, on the copy engines -----------------------
 * The reference has no multi-GPU path; BASELINE configs 4 and 5 gather the per-rank maps.
 * The hot kernel is persistent (one CTA per SM, full register file), so a collective that needs SMs cannot
 * overlap it.  These calls let each rank push its finished maps straight into every other rank's gather
 * buffer with device-to-device copies over NVLink (copy engines, no SM), stream-ordered after the kernels
 * that produced them, while the SMs already work on the next launch sequence.
 *   create  : device buffer other processes of this box can map + the 64-byte handle to send them
 *   open    : map a peer's buffer from its handle (a different process; never the creating one)
 *   push    : once `after_stream` reaches this point, copy src -> dst_ptrs[i] + dst_offset for every i
 *             (NULL entries are skipped; pointers may be local or opened peers)
 *   mark    : returns a ticket that completes when every push enqueued so far has finished
 *   wait    : makes `cuda_stream` wait for a ticket (at most 64 tickets are live) */
#define STEREO_IPC_HANDLE_BYTES 64
int stereo_peer_buffer_create(stereo_ctx* ctx, size_t bytes, void** dev_ptr, unsigned char* handle_out);
int stereo_peer_buffer_open(stereo_ctx* ctx, const unsigned char* handle, void** peer_ptr);
int stereo_peer_buffer_close(stereo_ctx* ctx, void* peer_ptr);
int stereo_peer_buffer_destroy(stereo_ctx* ctx, void* dev_ptr);
int stereo_peer_push(stereo_ctx* ctx, void* const* dst_ptrs, int n_dst, size_t dst_offset, const void* src, size_t bytes,
                     void* after_stream);
int stereo_peer_mark(stereo_ctx* ctx, int* ticket_out);
int stereo_peer_wait(stereo_ctx* ctx, int ticket, void* cuda_stream);

/* ---- device images and preprocessing ----------------------------------------------------- */
/*
 * What the ps2 executable does to an image between imread and the disparity call, on the device, so that an image pair
 * is uploaded ONCE (as 8-bit pixels) and serves both directions and every problem that uses it; the reference uploads
 * both float Mats again for every direction (DisparitySSD.cu:171-174).
 *   stereo_dev_alloc/free/upload/download : plain device buffers for callers without the CUDA runtime (synchronous copies)
 *   gray   : channels == 1: Mat::convertTo(CV_32FC1) unscaled (main.cpp:87-88); 3 or 4: cvtColor(COLOR_RGB2GRAY) on the
 *            interleaved data as loaded (BGR from imread, main.cpp:114-117) with OpenCV's fixed-point coefficients
 *            (shift 14 = OpenCV 3.4.1, 15 = later), then convertTo
 *   scale_add : out = a * scale (+ add, nullable): `left * 1.1f` (main.cpp:191-193) and `first + noise` (main.cpp:146-152),
 *            separately rounded float32 operations.  The noise field itself is cv::randn's sequential stream: host.
 */
int stereo_dev_alloc(stereo_ctx* ctx, size_t bytes, void** ptr);
int stereo_dev_free(stereo_ctx* ctx, void* ptr);
int stereo_dev_upload(stereo_ctx* ctx, void* dst, size_t dst_step, const void* src, size_t src_step, size_t row_bytes, int rows);
int stereo_dev_download(stereo_ctx* ctx, void* dst, size_t dst_step, const void* src, size_t src_step, size_t row_bytes, int rows);
int stereo_image_gray_f32_device(stereo_ctx* ctx, const uint8_t* img, size_t step, int rows, int cols, int channels, int shift,
                                 float* out, size_t out_step, void* cuda_stream);
int stereo_image_scale_add_f32_device(stereo_ctx* ctx, const float* a, size_t a_step, const float* add, size_t add_step, float scale,
                                      int rows, int cols, float* out, size_t out_step, void* cuda_stream);

/* ---- reference-GPU-semantics mode ------------------------------------------------- */
/*
 * Everything above reproduces the reference's CPU functions (the parity target).  Its GTX-1080 kernels compute a
 * different function: a (2R+1) x 2R window, clamp-to-edge addressing, every d in [min_disp, max_disp], float32 rolling
 * column sums restarted every 40 rows, SSD accepted only below 5e6 and NCC only above 0, -1 where nothing won
 * (lib/DisparitySSD.cu:16,27-141,177; lib/DisparityNCorr.cu:16,28-175,211).  These two entry points compute THAT function,
 * every float operation in the kernels' order, so maps the reference's GPU build produced can be reproduced.
 * disp_out: int8 (the kernels store `char`); best_out (nullable): the kernels' running best SSD / score map.
 * The range must fit the kernels' `char` loop variable.  Checked bit for bit against oracle_disparity_refgpu. */
int stereo_disparity_refgpu_f32_host(stereo_ctx* ctx, int cost,
                                     const float* ref, size_t ref_step, const float* tgt, size_t tgt_step,
                                     int rows, int cols, int window_rad, int min_disp, int max_disp,
                                     int8_t* disp_out, size_t disp_step, float* best_out, size_t best_step);
int stereo_disparity_refgpu_f32_device(stereo_ctx* ctx, int cost,
                                       const float* ref, size_t ref_step, const float* tgt, size_t tgt_step,
                                       int rows, int cols, int window_rad, int min_disp, int max_disp,
                                       int8_t* disp_out, size_t disp_step, float* best_out, size_t best_step,
                                       void* cuda_stream);

/* ---- one row band of a pair, HOST images (a device's share of a row-band sharded pair) -------------- */
/* left / right: FULL-image host origins (cv::Mat data pointers); only the slab of rows the band [row_begin, row_end) needs -
 * window halo plus the row of the SSD flat-index wrap, stereo_band_halo_rows - is uploaded.  disp_left / disp_right point
 * at the band's first output row.  Synchronous.  The f32 form converts the slab to u8 on the host; images that are not
 * 8-bit-valued return STEREO_ERR_UNSUPPORTED (callers take a whole-image call then). */
int stereo_disparity_pair_band_u8_host(stereo_ctx* ctx, int cost,
                                       const uint8_t* left, size_t left_step, const uint8_t* right, size_t right_step,
                                       int rows, int cols, int row_begin, int row_end, int window_rad, int disparity_range,
                                       void* disp_left, void* disp_right, size_t disp_step, int disp_elem_bytes);
int stereo_disparity_pair_band_f32_host(stereo_ctx* ctx, int cost,
                                        const float* left, size_t left_step, const float* right, size_t right_step,
                                        int rows, int cols, int row_begin, int row_end, int window_rad, int disparity_range,
                                        void* disp_left, void* disp_right, size_t disp_step, int disp_elem_bytes);

/* ---- several GPUs in ONE process (SURVEY.md §8b "context create/destroy taking a device list", §8e) ---------- */
/*
 * The reference is single-GPU (lib/DisparitySSD.cu:143-207).  A stereo_mgpu owns one stereo_ctx per listed device
 * (devices = NULL: every visible sm_100 device) and one host thread per device during a call.  The path has no exchange
 * inside the computation: batches are sharded by pair (contiguous blocks), one pair by row bands with halo; every device
 * uploads only its share and downloads its maps straight into the caller's host arrays - results are bit-identical to the
 * single-GPU calls.  A device ordinal may be listed more than once (several contexts on one GPU).
 */
typedef struct stereo_mgpu stereo_mgpu;
int stereo_mgpu_create(const int* devices, int n_devices, stereo_mgpu** out);
void stereo_mgpu_destroy(stereo_mgpu* mg);
int stereo_mgpu_device_count(const stereo_mgpu* mg);
/* The context of the index-th listed device (for stereo_ctx_set_* knobs); owned by the stereo_mgpu. */
stereo_ctx* stereo_mgpu_ctx(stereo_mgpu* mg, int index);
/* Batches (BASELINE config 5): same arguments as stereo_disparity_pair_batch_{u8,f32}_host. */
int stereo_mgpu_disparity_pair_batch_u8_host(stereo_mgpu* mg, int cost, int n_pairs,
                                             const uint8_t* left, const uint8_t* right, size_t img_step, size_t pair_stride,
                                             int rows, int cols, int window_rad, int disparity_range,
                                             void* disp_left, void* disp_right, size_t disp_step, size_t disp_pair_stride,
                                             int disp_elem_bytes);
int stereo_mgpu_disparity_pair_batch_f32_host(stereo_mgpu* mg, int cost, int n_pairs,
                                              const float* left, const float* right, size_t img_step, size_t pair_stride,
                                              int rows, int cols, int window_rad, int disparity_range,
                                              void* disp_left, void* disp_right, size_t disp_step, size_t disp_pair_stride,
                                              int disp_elem_bytes);
/* One pair in row bands (BASELINE config 4): same arguments as stereo_disparity_pair_{u8,f32}_host.  Float images that are
 * not 8-bit-valued are computed whole on the first device. */
int stereo_mgpu_disparity_pair_bands_u8_host(stereo_mgpu* mg, int cost,
                                             const uint8_t* left, size_t left_step, const uint8_t* right, size_t right_step,
                                             int rows, int cols, int window_rad, int disparity_range,
                                             void* disp_left, void* disp_right, size_t disp_step, int disp_elem_bytes);
int stereo_mgpu_disparity_pair_bands_f32_host(stereo_mgpu* mg, int cost,
                                              const float* left, size_t left_step, const float* right, size_t right_step,
                                              int rows, int cols, int window_rad, int disparity_range,
                                              void* disp_left, void* disp_right, size_t disp_step, int disp_elem_bytes);

/* Blocks until everything enqueued on `cuda_stream` (NULL = the context's stream) has finished and
 * returns any asynchronous error. */
int stereo_ctx_synchronize(stereo_ctx* ctx, void* cuda_stream);

#ifdef __cplusplus
} /* extern "C" */
#endif

#endif /* STEREO_B200_H_ */
