// Co-issue micro-benchmark for B200 (sm_100a): can the FP64 pipe (DFMA/DADD) and the integer-multiply
// pipe (IMAD.WIDE.U32) run at their own peak rates at the same time, from the same warp or from
// different warps of one SM?  Decides whether a 52-bit-limb DFMA Montgomery multiplier can run BESIDE
// the 32-bit-limb IMAD one (DESIGN.md "open levers").
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dfma_mix dfma_mix.cu && ./dfma_mix
// One JSON line per mix: chip-wide G instr/s of each counted class (CUDA events).
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <algorithm>
#include <cuda_runtime.h>

#define ITERS 8192
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)

// Two fused wide pairs chained through the carry flag (the multiplier's row primitive, halved).
#define WIDE2(A0, A1, MA, B0, B1)                                                         \
  asm volatile("{ .reg .u32 l0, h0, l1, h1;\n\tmov.b64 {l0, h0}, %0;\n\tmov.b64 {l1, h1}, %1;\n\t" \
               "mad.lo.cc.u32 l0, %2, %3, l0;\n\tmadc.hi.cc.u32 h0, %2, %3, h0;\n\t"      \
               "madc.lo.cc.u32 l1, %2, %4, l1;\n\tmadc.hi.u32 h1, %2, %4, h1;\n\t"        \
               "mov.b64 %0, {l0, h0};\n\tmov.b64 %1, {l1, h1}; }"                         \
               : "+l"(A0), "+l"(A1) : "r"(MA), "r"(B0), "r"(B1));

// One 52x52-bit limb product the way a DFMA multiplier does it: hi and lo halves by two DFMA.RZ and one
// DADD, both accumulated as 64-bit integer bit patterns (IADD3 + IMAD.X each).
#define DPROD(X, Y, SH, SL)                                                               \
  { asm volatile("" : "+d"(X)); double hi_ = __fma_rz(X, Y, 0x1.0p104); double sub_ = 0x1.0000000000001p104 - hi_;    \
    double lo_ = __fma_rz(X, Y, sub_);                                                    \
    SH += (unsigned long long)__double_as_longlong(hi_);                                  \
    SL += (unsigned long long)__double_as_longlong(lo_);                                  \
    asm volatile("" : "+l"(SH), "+l"(SL)); }

enum { MODE_WIDE = 0, MODE_DFMA = 1, MODE_THREAD_MIX = 2, MODE_WARP_MIX = 3, MODE_DPROD = 4, MODE_WARP_MIX_DPROD = 5 };

// K = DFMAs per 2 wides in the same-thread mix; ODD_EVERY = 1 of every ODD_EVERY warps is an fp64 warp
template <int MODE, int K, int FP_OF, int FP_NUM>
__global__ void __launch_bounds__(256) kern(uint32_t* out, uint32_t seed) {
  unsigned long long acc[8], acc2[8], sh[8], sl[8];
  uint32_t a[8];
  double d[8], x[8];
  uint32_t b = seed * 2654435761u + threadIdx.x, b2 = b ^ 0x1234567u;
  double db = 1.0000001 + threadIdx.x * 1e-9;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    a[i] = (b + i * 0x9e3779b9u) ^ seed; acc[i] = i + threadIdx.x; acc2[i] = acc[i] * 3;
    d[i] = 1.0 + i; sh[i] = i; sl[i] = 2 * i; x[i] = (double)((a[i] & 0xfffff) | 0x1000000) * 1048576.0 + (double)i;
  }
  const int warp = threadIdx.x >> 5;
  const bool fp_warp = (warp % FP_OF) < FP_NUM;
#pragma unroll 1
  for (int it = 0; it < ITERS / 4; it++) {
#pragma unroll
    for (int rep = 0; rep < 4; rep++) {
      if (MODE == MODE_WIDE) {
#pragma unroll
        for (int i = 0; i < 8; i++) WIDE2(acc[i], acc2[i], a[i], b, b2)
      } else if (MODE == MODE_DFMA) {
#pragma unroll
        for (int i = 0; i < 8; i++) asm volatile("fma.rn.f64 %0, %0, %1, %0;" : "+d"(d[i]) : "d"(db));
      } else if (MODE == MODE_THREAD_MIX) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
          WIDE2(acc[i], acc2[i], a[i], b, b2)
#pragma unroll
          for (int k = 0; k < K; k++) asm volatile("fma.rn.f64 %0, %0, %1, %0;" : "+d"(d[(i + k) & 7]) : "d"(db));
        }
      } else if (MODE == MODE_WARP_MIX) {
        if (fp_warp) {
#pragma unroll
          for (int q = 0; q < 4; q++)
#pragma unroll
            for (int i = 0; i < 8; i++) asm volatile("fma.rn.f64 %0, %0, %1, %0;" : "+d"(d[i]) : "d"(db));
        } else {
#pragma unroll
          for (int i = 0; i < 8; i++) WIDE2(acc[i], acc2[i], a[i], b, b2)
        }
      } else if (MODE == MODE_DPROD) {
#pragma unroll
        for (int i = 0; i < 8; i++) DPROD(x[i], x[(i + 1) & 7], sh[i], sl[i])
      } else if (MODE == MODE_WARP_MIX_DPROD) {
        if (fp_warp) {
#pragma unroll
          for (int i = 0; i < 8; i++) DPROD(x[i], x[(i + 1) & 7], sh[i], sl[i])
        } else {
#pragma unroll
          for (int i = 0; i < 8; i++) WIDE2(acc[i], acc2[i], a[i], b, b2)
        }
      }
    }
  }
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++)
    s += (uint32_t)acc[i] ^ (uint32_t)(acc[i] >> 32) ^ (uint32_t)acc2[i] ^ (uint32_t)(acc2[i] >> 32) ^ (uint32_t)d[i] ^
         (uint32_t)sh[i] ^ (uint32_t)(sh[i] >> 32) ^ (uint32_t)sl[i] ^ (uint32_t)(sl[i] >> 32);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

typedef void (*kern_t)(uint32_t*, uint32_t);

// wide_frac / fp_frac: fraction of the threads that run each class; per-step counts per thread per rep
static void run(const char* name, kern_t kfn, int bps, int nsm, double wide_per_rep, double wide_frac,
                double fp_per_rep, double fp_frac, const char* fp_unit) {
  int grid = nsm * bps, block = 256;
  uint32_t* out; CK(cudaMalloc(&out, (size_t)grid * block * 4));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int w = 0; w < 3; w++) kfn<<<grid, block>>>(out, 12345u + w);
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < 5; r++) {
    CK(cudaEventRecord(e0)); kfn<<<grid, block>>>(out, 777u + r); CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1)); float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); best = std::min(best, ms);
  }
  double threads = (double)grid * block;
  double wide = threads * wide_frac * ITERS * wide_per_rep, fp = threads * fp_frac * ITERS * fp_per_rep;
  printf("{\"mix\": \"%s\", \"blocks_per_sm\": %d, \"ms\": %.4f, \"wide_T_per_s\": %.3f, \"%s_T_per_s\": %.3f}\n",
         name, bps, best, wide / best * 1e-9, fp_unit, fp / best * 1e-9);
  CK(cudaFree(out));
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  printf("{\"device\": \"%s\", \"sms\": %d}\n", p.name, p.multiProcessorCount);
  int nsm = p.multiProcessorCount;
  for (int bps : {2, 4}) {
    run("wide_only", kern<MODE_WIDE, 0, 2, 1>, bps, nsm, 16, 1.0, 0, 0, "dfma");
    run("dfma_only", kern<MODE_DFMA, 0, 2, 1>, bps, nsm, 0, 0, 8, 1.0, "dfma");
    run("same_thread 2wide+1dfma", kern<MODE_THREAD_MIX, 1, 2, 1>, bps, nsm, 16, 1.0, 8, 1.0, "dfma");
    run("same_thread 2wide+2dfma", kern<MODE_THREAD_MIX, 2, 2, 1>, bps, nsm, 16, 1.0, 16, 1.0, "dfma");
    run("same_thread 2wide+4dfma", kern<MODE_THREAD_MIX, 4, 2, 1>, bps, nsm, 16, 1.0, 32, 1.0, "dfma");
    run("warp_mix 1of2 dfma(x4)", kern<MODE_WARP_MIX, 0, 2, 1>, bps, nsm, 16, 0.5, 32, 0.5, "dfma");
    run("warp_mix 1of4 dfma(x4)", kern<MODE_WARP_MIX, 0, 4, 1>, bps, nsm, 16, 0.75, 32, 0.25, "dfma");
    run("dprod_only (2dfma+dadd+2add64)", kern<MODE_DPROD, 0, 2, 1>, bps, nsm, 0, 0, 8, 1.0, "dprod");
    run("warp_mix 1of2 dprod", kern<MODE_WARP_MIX_DPROD, 0, 2, 1>, bps, nsm, 16, 0.5, 8, 0.5, "dprod");
    run("warp_mix 1of4 dprod", kern<MODE_WARP_MIX_DPROD, 0, 4, 1>, bps, nsm, 16, 0.75, 8, 0.25, "dprod");
    run("warp_mix 3of4 dprod", kern<MODE_WARP_MIX_DPROD, 0, 4, 3>, bps, nsm, 16, 0.25, 8, 0.75, "dprod");
  }
  return 0;
}
