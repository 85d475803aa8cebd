// Instruction-issue microbenchmark for sm_100a (B200): measures warp-instructions / cycle / SMSP
// for the integer/FP32 instructions the stereo cost kernels are built from, alone and in the
// mixes the kernels actually issue.  Output: one line per (test, warps-per-SMSP).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <string>

#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} }while(0)

constexpr int ITERS = 512;
constexpr int UNROLL = 8;   // body repeats per loop iteration
constexpr int CH = 8;       // independent chains

#define A_FFMA(i)  asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(fa), "f"(fb));
#define A_FADD(i)  asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(fa));
#define A_FMUL(i)  asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(fa));
#define A_IMAD(i)  asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(ia), "r"(ib));
#define A_IADD3(i) asm volatile("{.reg .s32 t; add.s32 t, %0, %1; sub.s32 %0, t, %2;}" : "+r"(x[i]) : "r"(ia), "r"(ib));
#define A_IADD(i)  asm volatile("add.s32 %0, %0, %1;" : "+r"(x[i]) : "r"(ia));
#define A_DP2A(i)  asm volatile("dp2a.lo.s32.u32 %0, %1, %2, %0;" : "+r"(x[i]) : "r"(ia), "r"(ib));
#define A_DP4A(i)  asm volatile("dp4a.s32.u32 %0, %1, %2, %0;" : "+r"(x[i]) : "r"(ia), "r"(ib));
#define A_MIN(i)   asm volatile("min.s32 %0, %0, %1;" : "+r"(x[i]) : "r"(y[i]));
#define A_MINMAX(i) asm volatile("min.s32 %0, %0, %1; max.s32 %0, %0, %2;" : "+r"(x[i]) : "r"(y[i]), "r"(z[i]));
#define A_LOP(i)   asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(ia), "r"(ib));
#define A_SHL(i)   asm volatile("shf.l.wrap.b32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(ia), "r"(ib));
#define A_PRMT(i)  asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(ia), "r"(ib));
#define A_I2F(i)   asm volatile("{.reg .f32 t; cvt.rn.f32.s32 t, %0; mov.b32 %0, t;}" : "+r"(x[i]));
#define A_F2I(i)   asm volatile("{.reg .s32 t; cvt.rzi.s32.f32 t, %0; mov.b32 %0, t;}" : "+f"(f[i]));
#define A_SHFL(i)  asm volatile("shfl.sync.bfly.b32 %0, %0, 1, 0x1f, 0xffffffff;" : "+r"(x[i]));
#define A_REDUX(i) asm volatile("redux.sync.min.s32 %0, %0, 0xffffffff;" : "+r"(x[i]));
#define A_LDS32(i) asm volatile("ld.shared.b32 %0, [%1];" : "=r"(x[i]) : "r"(saddr + (i)*128));
#define A_LDS128(i) asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(x[i]), "=r"(y[i]), "=r"(z[i]), "=r"(w[i]) : "r"(saddr4 + (i)*512));
#define A_LDS128B(i) asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(x[i]), "=r"(y[i]), "=r"(z[i]), "=r"(w[i]) : "r"(sbase + (i)*16));
#define A_LDS64(i) asm volatile("ld.shared.v2.b32 {%0,%1}, [%2];" : "=r"(x[i]), "=r"(y[i]) : "r"(saddr2 + (i)*256));
#define A_FSEL(i)  asm volatile("{.reg .pred p; setp.gt.f32 p, %1, %0; selp.f32 %0, %1, %0, p; selp.s32 %2, %3, %2, p;}" : "+f"(f[i]), "+f"(g[i]), "+r"(x[i]) : "r"(ia));
#define A_FMAX(i)  asm volatile("max.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(g[i]));
#define A_LEA(i)   asm volatile("{.reg .s32 t; shl.b32 t, %0, 9; add.s32 %0, t, %1;}" : "+r"(x[i]) : "r"(ib));

#define REP8(M) M(0) M(1) M(2) M(3) M(4) M(5) M(6) M(7)

// mixes (each REP8 group is 8 instrs on independent chains)
#define MIX_DP_ADD3      REP8(A_DP2A) REP8(A_IADD3)
#define MIX_KERNEL       REP8(A_DP2A) REP8(A_IADD3) REP8(A_IMAD) REP8(A_MIN)
#define MIX_KERNEL_LEA   REP8(A_DP2A) REP8(A_IADD3) REP8(A_LEA) REP8(A_MIN)
#define MIX_KERNEL15     REP8(A_DP2A) REP8(A_DP2A) REP8(A_DP2A) REP8(A_IADD3) REP8(A_IADD3) REP8(A_IADD3) REP8(A_IMAD) REP8(A_IMAD) REP8(A_MIN) REP8(A_MIN)
#define MIX_FFMA_KERNEL  REP8(A_FFMA) REP8(A_FFMA) REP8(A_IADD3) REP8(A_IMAD) REP8(A_MIN)
#define MIX_FFMA_IADD    REP8(A_FFMA) REP8(A_IADD3)
#define MIX_FFMA_IMAD    REP8(A_FFMA) REP8(A_IMAD)
#define MIX_FFMA_DP      REP8(A_FFMA) REP8(A_DP2A)
#define MIX_IMAD_IADD    REP8(A_IMAD) REP8(A_IADD3)
#define MIX_FFMA3_IADD   REP8(A_FFMA) REP8(A_FFMA) REP8(A_FFMA) REP8(A_IADD3)
#define MIX_NCC          REP8(A_DP2A) REP8(A_IADD3) REP8(A_FFMA) REP8(A_FSEL)
#define A_FFMA3(i) asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(g[i]) : "f"(f[i]), "f"(fa), "f"(fb));
#define A_KEYLOP(i) asm volatile("lop3.b32 %0, %1, 0xffffff80, %2, 0xf8;" : "=r"(y[i]) : "r"(x[i]), "r"(ia));
#define MIX_NCC5      REP8(A_DP2A) REP8(A_IADD3) REP8(A_FFMA) REP8(A_KEYLOP) REP8(A_MIN)
#define MIX_NCC7      REP8(A_DP2A) REP8(A_IADD3) REP8(A_I2F) REP8(A_FMUL) REP8(A_KEYLOP) REP8(A_MIN)
#define MIX_I2F_DP    REP8(A_I2F) REP8(A_DP2A)
#define MIX_I2F_ADD3  REP8(A_I2F) REP8(A_IADD3)
#define MIX_DP_LDS       REP8(A_DP2A) REP8(A_DP2A) REP8(A_DP2A) REP8(A_DP2A) A_LDS128(0)

template<int TEST> __global__ void __launch_bounds__(1024,1) kern(int* out, long long* cyc, int ia, int ib, float fa, float fb)
{
  __shared__ __align__(16) int sm[8192];
  for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = i * ia;
  __syncthreads();
  int x[CH], y[CH], z[CH], w[CH]; float f[CH], g[CH];
  #pragma unroll
  for (int i = 0; i < CH; i++) { x[i] = threadIdx.x + i; y[i] = ia + i; z[i]=0; w[i]=0; f[i] = threadIdx.x * 0.5f + i; g[i] = fa * i; }
  unsigned sbase = (unsigned)__cvta_generic_to_shared(sm);
  unsigned lane = threadIdx.x & 31;
  unsigned saddr = sbase + lane * 4; unsigned saddr2 = sbase + lane * 8; unsigned saddr4 = sbase + lane * 16;
  (void)saddr; (void)saddr2; (void)saddr4;
  __syncthreads();
  long long t0 = clock64();
  #pragma unroll 1
  for (int it = 0; it < ITERS; it++) {
    #pragma unroll
    for (int u = 0; u < UNROLL; u++) {
      if (TEST == 0) { REP8(A_FFMA) }
      if (TEST == 1) { REP8(A_FADD) }
      if (TEST == 2) { REP8(A_IMAD) }
      if (TEST == 3) { REP8(A_IADD3) }
      if (TEST == 4) { REP8(A_DP2A) }
      if (TEST == 5) { REP8(A_DP4A) }
      if (TEST == 6) { REP8(A_MIN) }
      if (TEST == 7) { REP8(A_LOP) }
      if (TEST == 8) { REP8(A_SHL) }
      if (TEST == 9) { REP8(A_PRMT) }
      if (TEST == 10) { REP8(A_I2F) }
      if (TEST == 11) { REP8(A_F2I) }
      if (TEST == 12) { REP8(A_SHFL) }
      if (TEST == 13) { REP8(A_REDUX) }
      if (TEST == 14) { REP8(A_LDS32) }
      if (TEST == 15) { REP8(A_LDS64) }
      if (TEST == 16) { REP8(A_LDS128) }
      if (TEST == 17) { REP8(A_LDS128B) }
      if (TEST == 18) { REP8(A_FSEL) }
      if (TEST == 19) { REP8(A_FMAX) }
      if (TEST == 20) { REP8(A_IADD) }
      if (TEST == 21) { REP8(A_LEA) }
      if (TEST == 22) { REP8(A_FMUL) }
      if (TEST == 23) { REP8(A_MINMAX) }
      if (TEST == 30) { MIX_DP_ADD3 }
      if (TEST == 31) { MIX_KERNEL }
      if (TEST == 32) { MIX_KERNEL_LEA }
      if (TEST == 33) { MIX_KERNEL15 }
      if (TEST == 34) { MIX_FFMA_KERNEL }
      if (TEST == 35) { MIX_FFMA_IADD }
      if (TEST == 36) { MIX_FFMA_IMAD }
      if (TEST == 37) { MIX_FFMA_DP }
      if (TEST == 38) { MIX_IMAD_IADD }
      if (TEST == 39) { MIX_FFMA3_IADD }
      if (TEST == 40) { MIX_NCC }
      if (TEST == 41) { MIX_DP_LDS }
      if (TEST == 42) { MIX_NCC5 }
      if (TEST == 43) { MIX_NCC7 }
      if (TEST == 44) { MIX_I2F_DP }
      if (TEST == 45) { MIX_I2F_ADD3 }
    }
  }
  long long t1 = clock64();
  int acc = 0; float facc = 0;
  #pragma unroll
  for (int i = 0; i < CH; i++) { acc += x[i] + y[i] + z[i] + w[i]; facc += f[i] + g[i]; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc + (int)facc;
  if ((threadIdx.x & 31) == 0) cyc[blockIdx.x * 32 + (threadIdx.x >> 5)] = t1 - t0;
}

struct Test { int id; const char* name; int instr_per_body; };
static const Test tests[] = {
  {0,"FFMA",8},{1,"FADD",8},{22,"FMUL",8},{2,"IMAD",8},{3,"IADD3(a+b-c)",8},{20,"IADD",8},{21,"LEA(shl+add)",8},{4,"IDP.2A",8},{5,"IDP.4A",8},{6,"VIMNMX3 (2 mins fused)",4},{23,"VIMNMX min,max",16},
  {7,"LOP3",8},{8,"SHF",8},{9,"PRMT",8},{10,"I2F",8},{11,"F2I",8},{12,"SHFL.BFLY",8},{13,"REDUX.MIN",8},
  {14,"LDS.32",8},{15,"LDS.64",8},{16,"LDS.128",8},{17,"LDS.128 bcast",8},{18,"FSETP+2SEL (3 instr)",24},{19,"FMNMX",8},
  {30,"mix DP2A:IADD3 1:1",16},{31,"mix DP2A:IADD3:IMAD:MIN",32},{32,"mix DP2A:IADD3:LEA:MIN",32},{33,"mix 3DP:3ADD3:2IMAD:2MIN",80},
  {34,"mix 2FFMA:IADD3:IMAD:MIN",40},{35,"mix FFMA:IADD3",16},{36,"mix FFMA:IMAD",16},{37,"mix FFMA:DP2A",16},{38,"mix IMAD:IADD3",16},{39,"mix 3FFMA:IADD3",32},
  {40,"mix DP2A:IADD3:FFMA:FSEL3",48},{41,"mix 32DP2A:1LDS128",33},{42,"mix NCC R<=5 DP:ADD3:FFMA:LOP:MIN3",36},{43,"mix NCC R>=6 DP:ADD3:I2F:FMUL:LOP:MIN3",44},{44,"mix I2F:DP2A",16},{45,"mix I2F:IADD3",16},
};

template<int T> void launch(int grid, int block, int* out, long long* cyc, cudaStream_t s) { kern<T><<<grid, block, 0, s>>>(out, cyc, 3, 5, 1.0001f, 0.5f); }
typedef void (*LaunchFn)(int,int,int*,long long*,cudaStream_t);
static LaunchFn fn(int id) {
  switch(id){
#define C(n) case n: return launch<n>;
  C(0)C(1)C(2)C(3)C(4)C(5)C(6)C(7)C(8)C(9)C(10)C(11)C(12)C(13)C(14)C(15)C(16)C(17)C(18)C(19)C(20)C(21)C(22)C(23)
  C(30)C(31)C(32)C(33)C(34)C(35)C(36)C(37)C(38)C(39)C(40)C(41)C(42)C(43)C(44)C(45)
  }
  return nullptr;
}

int main(int argc, char** argv) {
  int dev = 0; CK(cudaSetDevice(dev));
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, dev));
  int sms = p.multiProcessorCount;
  printf("# device %s, %d SMs, clockRate %d kHz\n", p.name, sms, p.clockRate);
  int* out; long long* cyc; CK(cudaMalloc(&out, sizeof(int) * sms * 1024)); CK(cudaMalloc(&cyc, sizeof(long long) * sms * 32));
  std::vector<long long> h(sms * 32);
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  printf("%-32s %5s %12s %12s %10s\n", "test", "w/SMSP", "inst/clk/SMSP", "lanes/clk/SM", "eff_MHz");
  for (const Test& t : tests) {
    for (int wps : {1, 2, 4, 8}) {
      int block = wps * 128;
      LaunchFn f = fn(t.id);
      f(sms, block, out, cyc, 0); CK(cudaDeviceSynchronize());
      CK(cudaEventRecord(e0)); f(sms, block, out, cyc, 0); CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
      CK(cudaMemcpy(h.data(), cyc, sizeof(long long) * sms * 32, cudaMemcpyDeviceToHost));
      long long mx = 0; for (int b = 0; b < sms; b++) for (int w = 0; w < block / 32; w++) mx = std::max(mx, h[b * 32 + w]);
      double inst = (double)ITERS * UNROLL * t.instr_per_body;       // per warp
      double ipc = inst * wps / (double)mx;                           // per SMSP
      printf("%-32s %5d %12.3f %12.1f %10.0f\n", t.name, wps, ipc, ipc * 4 * 32, (double)mx / (ms * 1e3));
    }
  }
  return 0;
}
