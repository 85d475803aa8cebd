// Instruction-mix microbenchmark (sm_100a): what issue rate does the CORE instruction mix of the fused SSD pair kernel
// reach when nothing but the pipes, the register file and the schedulers limit it?
//
// One "pixel step" of fast_row (4 candidates per lane, both maps) is, per warp: 6 IDP.2A (column sums, K + 2R columns for K
// pixels), 4 IADD3 (horizontal sums), 8 IMAD with an immediate multiplier (own and partner keys), 2 minima for the own key
// (VIMNMX3 + VIMNMX), 2 VIMNMX3 for the diagonals, 1 REDUX, 1 SHFL, ~1.5 LDS.128 of operands.  The loop below issues exactly
// that, with as many independent accumulators as the real kernel has (column sums: 24 of the 120), operands that stay in
// registers, no barriers, no global memory: its inst/clk/SMSP is the ceiling the real kernel's 0.69 is measured against.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mix mix.cu && ./mix
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} }while(0)
constexpr int ITERS = 2048;

__device__ __forceinline__ int dp2a(int a, unsigned b, int c) { int d; asm volatile("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }
__device__ __forceinline__ int add3(int a, int b, int c) { int d; asm volatile("{.reg .s32 t; add.s32 t, %1, %2; sub.s32 %0, t, %3;}" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }
__device__ __forceinline__ unsigned key256(int s, unsigned e) { unsigned d; asm volatile("mad.lo.u32 %0, %1, 256, %2;" : "=r"(d) : "r"(s), "r"(e)); return d; }
__device__ __forceinline__ unsigned min2(unsigned a, unsigned b) { unsigned d; asm volatile("min.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }
__device__ __forceinline__ unsigned min3(unsigned a, unsigned b, unsigned c) { unsigned d; asm volatile("{.reg .u32 t; min.u32 t, %1, %2; min.u32 %0, t, %3;}" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }

// MODE 0: the full step; 1: without REDUX / SHFL / LDS (arithmetic only); 2: arithmetic with the minima as plain VIMNMX
template <int MODE>
__global__ void __launch_bounds__(512, 1) kern(int* out, long long* cyc, int seed) {
    extern __shared__ int4 sm[];
    int col[4][11];                 // 4 x 11 column sums in flight (the real kernel: 4 x 30)
    int s[4];
    unsigned acc[4], held3 = ~0u, held1 = ~0u;
    int lp[4], rq[8], e2[8], el[4];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = make_int4(seed + i, seed * 3 + i, seed ^ i, i);
#pragma unroll
    for (int m = 0; m < 4; ++m) {
        s[m] = seed + m; acc[m] = ~0u - m; lp[m] = seed * (m + 3); el[m] = seed + 11 * m;
#pragma unroll
        for (int c = 0; c < 11; ++c) col[m][c] = threadIdx.x + c + m;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) { rq[i] = seed + 7 * i; e2[i] = seed * 5 + i; }
    __syncthreads();
    unsigned sink = 0;
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it += 4) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {          // four pixel steps per trip (the real loop is fully unrolled over a strip)
            // column sums: 6 IDP.2A on distinct accumulators (1.5 columns x 4 candidates)
            const int c0 = (k * 3) % 11, c1 = (k * 3 + 1) % 11;
#pragma unroll
            for (int m = 0; m < 4; ++m) col[m][c0] = dp2a(lp[k], unsigned(rq[(k + m) & 7]), col[m][c0]);
            if (k & 1) {
#pragma unroll
                for (int m = 0; m < 4; ++m) col[m][c1] = dp2a(lp[(k + 1) & 3], unsigned(rq[(k + m + 1) & 7]), col[m][c1]);
            }
            // horizontal sums, own and partner keys
            unsigned key[4], k2[4];
#pragma unroll
            for (int m = 0; m < 4; ++m) {
                s[m] = add3(s[m], col[m][c0], col[m][(c0 + 5) % 11]);
                key[m] = key256(s[m], unsigned(e2[(k + m) & 7]));
                k2[m] = key256(s[m], unsigned(el[k]));
            }
            // minima: own key of the pixel, diagonals of the partner
            unsigned best;
            if (MODE == 2) {
                best = min2(min2(key[0], key[1]), min2(key[2], key[3]));
                acc[0] = min2(acc[0], k2[0]); acc[1] = min2(acc[1], k2[1]); acc[2] = min2(acc[2], k2[2]); acc[3] = min2(acc[3], k2[3]);
            } else {
                best = min2(min3(key[0], key[1], key[2]), key[3]);
                acc[2] = min3(acc[2], held3, k2[2]);
                acc[0] = min3(acc[0], held1, k2[0]);
                held3 = k2[3]; held1 = k2[1];
            }
            unsigned done = acc[0];
            if (MODE == 0) {
                sink += __reduce_min_sync(0xffffffffu, best);
                done = __shfl_down_sync(0xffffffffu, done, 1);
                const int4 v = sm[(threadIdx.x + 64 * k) & 1023];                 // one LDS.128 of operands per step ...
                lp[k] = v.x; rq[k] = v.y; rq[k + 4] = v.z; e2[k] = v.w;
                if (k & 1) { const int4 w = sm[(threadIdx.x + 64 * k + 32) & 1023]; e2[k + 4] = w.x; el[k] = w.y; el[k - 1] = w.z; rq[k - 1] = w.w; }   // ... and a half
            } else {
                sink += best;
            }
            acc[0] = acc[1]; acc[1] = acc[2]; acc[2] = acc[3]; acc[3] = done;
        }
    }
    const long long t1 = clock64();
    int r = int(sink + held3 + held1);
#pragma unroll
    for (int m = 0; m < 4; ++m) {
        r += s[m] + int(acc[m]) + lp[m] + el[m];
#pragma unroll
        for (int c = 0; c < 11; ++c) r += col[m][c];
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) r += rq[i] + e2[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    if ((threadIdx.x & 31) == 0) cyc[blockIdx.x * 32 + (threadIdx.x >> 5)] = t1 - t0;
}

template <int MODE> static double run(int sms, int block, int* out, long long* cyc, std::vector<long long>& h, double inst_per_step) {
    const size_t smem = 160 * 1024;            // one CTA per SM, as the hot kernel
    CK(cudaFuncSetAttribute(kern<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    for (int rep = 0; rep < 2; ++rep) { kern<MODE><<<sms, block, smem>>>(out, cyc, 3); CK(cudaDeviceSynchronize()); }
    CK(cudaMemcpy(h.data(), cyc, sizeof(long long) * sms * 32, cudaMemcpyDeviceToHost));
    long long mx = 0;
    for (int b = 0; b < sms; ++b) for (int w = 0; w < block / 32; ++w) mx = std::max(mx, h[b * 32 + w]);
    return double(ITERS) * inst_per_step * (block / 128) / double(mx);
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    const int sms = p.multiProcessorCount;
    int* out; long long* cyc;
    CK(cudaMalloc(&out, sizeof(int) * sms * 512)); CK(cudaMalloc(&cyc, sizeof(long long) * sms * 32));
    std::vector<long long> h(sms * 32);
    printf("# %s, %d SMs; instructions per pixel step counted as the SASS of the loop body (cuobjdump), see the table\n", p.name, sms);
    printf("%-58s %6s %14s %14s\n", "mix", "w/SMSP", "inst/clk/SMSP", "clk/step/warp");
    for (int wps : {1, 2, 4}) {
        // SASS instructions per pixel step of the three loop bodies (cuobjdump -sass, CUDA 12.9): the full step adds address
        // arithmetic and moves for the operand loads to the 25.5 listed above
        const double full = 28.75, arith = 23.5, arith2 = 25.5;
        const double a = run<0>(sms, wps * 128, out, cyc, h, full), b = run<1>(sms, wps * 128, out, cyc, h, arith), c = run<2>(sms, wps * 128, out, cyc, h, arith2);
        printf("%-58s %6d %14.3f %14.1f\n", "pixel step: IDP 6, IADD3 4, IMAD 8, min 4, REDUX, SHFL, LDS 1.5", wps, a, full * wps / a);
        printf("%-58s %6d %14.3f %14.1f\n", "arithmetic only: IDP 6, IADD3 4, IMAD 8, min 4 (VIMNMX3)", wps, b, arith * wps / b);
        printf("%-58s %6d %14.3f %14.1f\n", "arithmetic only, minima as 7 VIMNMX", wps, c, arith2 * wps / c);
    }
    return 0;
}
