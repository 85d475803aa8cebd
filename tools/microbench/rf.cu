// Register-file read-bandwidth microbenchmark (sm_100a): same instructions as pipes.cu but with
// DISTINCT register operands per instruction, the way the stereo kernel issues them.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} }while(0)
constexpr int ITERS = 512, UNROLL = 8, CH = 8;

#define D_IDP3(i)   asm volatile("dp2a.lo.s32.u32 %0, %1, %2, %0;" : "+r"(x[i]) : "r"(a[i]), "r"(b[i]));
#define D_IDP2(i)   asm volatile("dp2a.lo.s32.u32 %0, %1, %2, %0;" : "+r"(x[i]) : "r"(a[0]), "r"(b[i]));
#define D_IDP1(i)   asm volatile("dp2a.lo.s32.u32 %0, %1, %2, %0;" : "+r"(x[i]) : "r"(a[0]), "r"(b[0]));
#define D_ADD3(i)   asm volatile("{.reg .s32 t; add.s32 t, %0, %1; sub.s32 %0, t, %2;}" : "+r"(x[i]) : "r"(a[i]), "r"(b[i]));
#define D_ADD2(i)   asm volatile("{.reg .s32 t; add.s32 t, %0, %1; sub.s32 %0, t, %2;}" : "+r"(x[i]) : "r"(a[i]), "r"(b[0]));
#define D_ADD1(i)   asm volatile("{.reg .s32 t; add.s32 t, %0, %1; sub.s32 %0, t, %2;}" : "+r"(x[i]) : "r"(a[0]), "r"(b[0]));
#define D_MAD3(i)   asm volatile("mad.lo.s32 %0, %1, %2, %0;" : "+r"(x[i]) : "r"(a[i]), "r"(b[i]));
#define D_MADI(i)   asm volatile("mad.lo.s32 %0, %1, 256, %0;" : "+r"(x[i]) : "r"(a[i]));
#define D_MADK(i)   asm volatile("mad.lo.s32 %0, %1, %2, %0;" : "+r"(x[i]) : "r"(a[i]), "r"(b[0]));
#define D_FMA3(i)   asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(f[i]) : "f"(g[i]), "f"(h[i]));
#define D_FMA2(i)   asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(f[i]) : "f"(g[0]), "f"(h[i]));
#define D_FMA1(i)   asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(f[i]) : "f"(g[0]), "f"(h[0]));
#define D_MIN2(i)   asm volatile("min.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(a[i]));
#define D_MIN3(i)   asm volatile("{.reg .u32 t; min.u32 t, %0, %1; min.u32 %0, t, %2;}" : "+r"(x[i]) : "r"(a[i]), "r"(b[i]));
#define D_LEA(i)    asm volatile("{.reg .s32 t; shl.b32 t, %1, 8; add.s32 %0, t, %0;}" : "+r"(x[i]) : "r"(a[i]));
#define REP8(M) M(0) M(1) M(2) M(3) M(4) M(5) M(6) M(7)
#define ALT8(A,B) A(0) B(0) A(1) B(1) A(2) B(2) A(3) B(3) A(4) B(4) A(5) B(5) A(6) B(6) A(7) B(7)

template<int TEST> __global__ void __launch_bounds__(1024,1) kern(int* out, long long* cyc, int seed)
{
  int x[CH], a[CH], b[CH]; float f[CH], g[CH], h[CH];
  #pragma unroll
  for (int i = 0; i < CH; i++) { x[i] = threadIdx.x + i; a[i] = seed * (i + 3); b[i] = seed + 7 * i; f[i] = i; g[i] = 1.0f + 1e-6f * (seed + i); h[i] = 0.5f * i + seed; }
  __syncthreads();
  long long t0 = clock64();
  #pragma unroll 1
  for (int it = 0; it < ITERS; it++) {
    #pragma unroll
    for (int u = 0; u < UNROLL; u++) {
      if (TEST == 0) { REP8(D_IDP3) }
      if (TEST == 1) { REP8(D_IDP2) }
      if (TEST == 2) { REP8(D_IDP1) }
      if (TEST == 3) { REP8(D_ADD3) }
      if (TEST == 4) { REP8(D_ADD2) }
      if (TEST == 5) { REP8(D_ADD1) }
      if (TEST == 6) { REP8(D_MAD3) }
      if (TEST == 7) { REP8(D_MADI) }
      if (TEST == 8) { REP8(D_MADK) }
      if (TEST == 9) { REP8(D_FMA3) }
      if (TEST == 10) { REP8(D_FMA2) }
      if (TEST == 11) { REP8(D_FMA1) }
      if (TEST == 12) { REP8(D_MIN2) }
      if (TEST == 13) { REP8(D_MIN3) }
      if (TEST == 14) { ALT8(D_IDP2, D_ADD3) }
      if (TEST == 15) { ALT8(D_IDP1, D_ADD1) }
      if (TEST == 16) { ALT8(D_IDP3, D_ADD3) }
      if (TEST == 17) { ALT8(D_IDP2, D_MIN2) }
      if (TEST == 18) { ALT8(D_IDP2, D_LEA) }
      if (TEST == 19) { ALT8(D_MADI, D_ADD3) }
      if (TEST == 20) { REP8(D_LEA) }
    }
  }
  long long t1 = clock64();
  int acc = 0; float facc = 0;
  #pragma unroll
  for (int i = 0; i < CH; i++) { acc += x[i] + a[i] + b[i]; facc += f[i] + g[i] + h[i]; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc + (int)facc;
  if ((threadIdx.x & 31) == 0) cyc[blockIdx.x * 32 + (threadIdx.x >> 5)] = t1 - t0;
}
struct Test { int id; const char* name; int n; };
static const Test tests[] = {
 {0,"IDP 3 distinct",8},{1,"IDP a shared (2 distinct)",8},{2,"IDP a,b shared (1 distinct)",8},
 {3,"IADD3 3 distinct",8},{4,"IADD3 2 distinct",8},{5,"IADD3 1 distinct",8},
 {6,"IMAD 3 distinct",8},{7,"IMAD imm (2 distinct)",8},{8,"IMAD k shared (2 distinct)",8},
 {9,"FFMA 3 distinct",8},{10,"FFMA 2 distinct",8},{11,"FFMA 1 distinct",8},
 {12,"VIMNMX 2 distinct",8},{13,"VIMNMX3 3 distinct",8},{20,"LEA 2 distinct",8},
 {14,"alt IDP(2d) / IADD3(3d)",16},{15,"alt IDP(1d) / IADD3(1d)",16},{16,"alt IDP(3d) / IADD3(3d)",16},
 {17,"alt IDP(2d) / VIMNMX(2d)",16},{18,"alt IDP(2d) / LEA(2d)",16},{19,"alt IMADimm(2d) / IADD3(3d)",16},
};
template<int T> void launch(int grid, int block, int* out, long long* cyc) { kern<T><<<grid, block>>>(out, cyc, 3); }
typedef void (*LaunchFn)(int,int,int*,long long*);
static LaunchFn fn(int id) { switch(id){
#define C(n) case n: return launch<n>;
 C(0)C(1)C(2)C(3)C(4)C(5)C(6)C(7)C(8)C(9)C(10)C(11)C(12)C(13)C(14)C(15)C(16)C(17)C(18)C(19)C(20)
 } return nullptr; }
int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0)); int sms = p.multiProcessorCount;
  int* out; long long* cyc; CK(cudaMalloc(&out, sizeof(int)*sms*1024)); CK(cudaMalloc(&cyc, sizeof(long long)*sms*32));
  std::vector<long long> h(sms*32);
  printf("%-34s %6s %14s\n", "test", "w/SMSP", "inst/clk/SMSP");
  for (const Test& t : tests) for (int wps : {1, 2, 4}) {
    int block = wps*128; LaunchFn f = fn(t.id);
    f(sms, block, out, cyc); CK(cudaDeviceSynchronize()); f(sms, block, out, cyc); CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(h.data(), cyc, sizeof(long long)*sms*32, cudaMemcpyDeviceToHost));
    long long mx = 0; for (int b = 0; b < sms; b++) for (int w = 0; w < block/32; w++) mx = std::max(mx, h[b*32+w]);
    printf("%-34s %6d %14.3f\n", t.name, wps, (double)ITERS*UNROLL*t.n*wps/(double)mx);
  }
  return 0;
}
