// Integer-pipe micro-benchmark for B200 (sm_100a): issue throughput, per SM per clock, of the
// instruction forms the Fr Montgomery kernels are built from.  Its IMAD.WIDE figure is the
// roofline denominator ("integer-multiply pipe peak") used by bench.py and DESIGN.md.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o imad_peak imad_peak.cu && ./imad_peak
// Prints one JSON line per mix: ops/clk/SM (clock64 deltas) and Gop/s chip-wide (CUDA events).
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>

#define ITERS 16384
#define CHAINS 8

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)

// Each BODY(i) is one or more instructions on chain i; accumulators live in 64-bit register pairs so the
// loop body contains only the instructions under test (checked with cuobjdump -sass).
#define DEFINE_KERNEL(NAME, BODY)                                                         \
  __global__ void __launch_bounds__(256) NAME(uint32_t* out, long long* cyc, uint32_t seed) { \
    unsigned long long acc[CHAINS], acc2[CHAINS];                                         \
    uint32_t a[CHAINS], x[CHAINS];                                                        \
    uint32_t b = seed * 2654435761u + threadIdx.x, b2 = b ^ 0x1234567u;                   \
    double d[CHAINS], db = 1.0000001 + threadIdx.x * 1e-9;                                \
    _Pragma("unroll") for (int i = 0; i < CHAINS; i++) {                                  \
      a[i] = b + i * 0x9e3779b9u; asm volatile("xor.b32 %0, %0, %1;" : "+r"(a[i]) : "r"(seed)); acc[i] = i + threadIdx.x + ((unsigned long long)blockIdx.x << 32); \
      acc2[i] = acc[i] * 3; x[i] = a[i] ^ 0x55aa55aau; d[i] = 1.0 + i;                    \
    }                                                                                     \
    long long t0 = clock64();                                                             \
    _Pragma("unroll 1") for (int it = 0; it < ITERS / 8; it++) {                          \
      _Pragma("unroll") for (int rep = 0; rep < 8; rep++) {                               \
        _Pragma("unroll") for (int i = 0; i < CHAINS; i++) { BODY }                       \
      }                                                                                   \
    }                                                                                     \
    long long t1 = clock64();                                                             \
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;                                      \
    uint32_t s = 0;                                                                       \
    _Pragma("unroll") for (int i = 0; i < CHAINS; i++)                                    \
      s += (uint32_t)acc[i] ^ (uint32_t)(acc[i] >> 32) ^ (uint32_t)acc2[i] ^ (uint32_t)(acc2[i] >> 32) ^ x[i] ^ a[i] ^ (uint32_t)d[i]; \
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + (uint32_t)db + b2;                   \
  }

// 32x32+64 -> 64, no carry: IMAD.WIDE.U32
#define WIDE_STEP(i) asm volatile("{ .reg .u32 l0, h0;\n\tmov.b64 {l0, h0}, %0;\n\tmad.wide.u32 %0, l0, %1, %0; }" : "+l"(acc[i]) : "r"(b));
DEFINE_KERNEL(k_wide, { WIDE_STEP(i) })
// mad.lo.cc + madc.hi.cc on a 64-bit pair, carry captured by addc: IMAD.WIDE.U32 Rd, Pc, ... + IADD3.X
DEFINE_KERNEL(k_wide_cc, {
  asm volatile("{ .reg .u32 l0, h0;\n\tmov.b64 {l0, h0}, %0;\n\t"
               "mad.lo.cc.u32 l0, %2, %3, l0;\n\tmadc.hi.cc.u32 h0, %2, %3, h0;\n\taddc.u32 %1, %1, 0;\n\t"
               "mov.b64 %0, {l0, h0}; }"
               : "+l"(acc[i]), "+r"(x[i]) : "r"(a[i]), "r"(b));
})
// two fused pairs chained through the carry flag: IMAD.WIDE.U32 (carry out) then IMAD.WIDE.U32.X (carry in)
DEFINE_KERNEL(k_wide_cc2, {
  asm volatile("{ .reg .u32 l0, h0, l1, h1;\n\tmov.b64 {l0, h0}, %0;\n\tmov.b64 {l1, h1}, %1;\n\t"
               "mad.lo.cc.u32 l0, %2, %3, l0;\n\tmadc.hi.cc.u32 h0, %2, %3, h0;\n\t"
               "madc.lo.cc.u32 l1, %2, %4, l1;\n\tmadc.hi.u32 h1, %2, %4, h1;\n\t"
               "mov.b64 %0, {l0, h0};\n\tmov.b64 %1, {l1, h1}; }"
               : "+l"(acc[i]), "+l"(acc2[i]) : "r"(a[i]), "r"(b), "r"(b2));
})
// four fused pairs in one carry chain + capture (the multiplier's row primitive): 4 wide + 1 IADD3.X
DEFINE_KERNEL(k_chain4, {
  asm volatile("{ .reg .u32 l0, h0, l1, h1;\n\tmov.b64 {l0, h0}, %0;\n\tmov.b64 {l1, h1}, %1;\n\t"
               "mad.lo.cc.u32 l0, %3, %4, l0;\n\tmadc.hi.cc.u32 h0, %3, %4, h0;\n\t"
               "madc.lo.cc.u32 l1, %3, %5, l1;\n\tmadc.hi.cc.u32 h1, %3, %5, h1;\n\t"
               "madc.lo.cc.u32 l0, %6, %4, l0;\n\tmadc.hi.cc.u32 h0, %6, %4, h0;\n\t"
               "madc.lo.cc.u32 l1, %6, %5, l1;\n\tmadc.hi.cc.u32 h1, %6, %5, h1;\n\t"
               "addc.u32 %2, %2, 0;\n\t"
               "mov.b64 %0, {l0, h0};\n\tmov.b64 %1, {l1, h1}; }"
               : "+l"(acc[i]), "+l"(acc2[i]), "+r"(x[i]) : "r"(a[i]), "r"(b), "r"(b2), "r"(seed));
})
DEFINE_KERNEL(k_lo, { asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(b), "r"(a[i])); })
DEFINE_KERNEL(k_hi, { asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(b), "r"(a[i])); })
DEFINE_KERNEL(k_iadd3, { asm volatile("add.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(a[i])); asm volatile("xor.b32 %0, %0, %1;" : "+r"(a[i]) : "r"(x[i])); })
DEFINE_KERNEL(k_addc, {
  asm volatile("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, %1, %3;" : "+r"(x[i]), "+r"(a[i]) : "r"(b2), "r"(b));
})
// one wide MAD + one / two ALU ops per chain step: do the fma and alu pipes overlap?
DEFINE_KERNEL(k_wide_plus_add, {
  WIDE_STEP(i)
  asm volatile("add.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(b));
})
DEFINE_KERNEL(k_wide_plus_2add, {
  WIDE_STEP(i)
  asm volatile("add.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(b));
  asm volatile("xor.b32 %0, %0, %1;" : "+r"(a[i]) : "r"(x[i]));
})
DEFINE_KERNEL(k_dfma, { asm volatile("fma.rn.f64 %0, %0, %1, %0;" : "+d"(d[i]) : "d"(db)); })

typedef void (*kern_t)(uint32_t*, long long*, uint32_t);

static void run(const char* name, kern_t kfn, int counted_per_body, int blocks_per_sm, int nsm, int clock_khz) {
  int grid = nsm * blocks_per_sm, block = 256;
  uint32_t* out; long long* cyc;
  CK(cudaMalloc(&out, (size_t)grid * block * 4));
  CK(cudaMalloc(&cyc, (size_t)grid * 8));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int w = 0; w < 3; w++) kfn<<<grid, block>>>(out, cyc, 12345u + w);
  CK(cudaDeviceSynchronize());
  float best_ms = 1e30f;
  for (int r = 0; r < 5; r++) {
    CK(cudaEventRecord(e0)); kfn<<<grid, block>>>(out, cyc, 777u + r); CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1)); float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); best_ms = std::min(best_ms, ms);
  }
  std::vector<long long> h(grid); CK(cudaMemcpy(h.data(), cyc, grid * 8, cudaMemcpyDeviceToHost));
  std::sort(h.begin(), h.end());
  double med_cyc = (double)h[grid / 2];
  double warp_instr_per_block = (double)ITERS * CHAINS * counted_per_body * (block / 32);
  // all resident blocks of an SM overlap for ~med_cyc cycles
  double lanes_per_clk_sm = warp_instr_per_block * 32.0 * blocks_per_sm / med_cyc;
  double total_ops = (double)grid * block * ITERS * CHAINS * counted_per_body;
  printf("{\"mix\": \"%s\", \"blocks_per_sm\": %d, \"thread_ops_per_clk_per_sm\": %.2f, \"gops\": %.1f, \"ms\": %.4f, \"median_cycles\": %.0f, \"eff_mhz\": %.0f}\n",
         name, blocks_per_sm, lanes_per_clk_sm, total_ops / best_ms * 1e-6, best_ms, med_cyc,
         med_cyc / (best_ms * 1e-3) * 1e-6);
  CK(cudaFree(out)); CK(cudaFree(cyc));
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz\": %d}\n", p.name, p.multiProcessorCount, khz);
  int nsm = p.multiProcessorCount;
  for (int bps : {2, 4, 8}) {
    run("imad_wide_u32", k_wide, 1, bps, nsm, khz);
    run("mad_lo_cc+madc_hi_cc(+addc)", k_wide_cc, 1, bps, nsm, khz);
    run("wide_cc_pair_chain(2 wide)", k_wide_cc2, 2, bps, nsm, khz);
    run("chain4(4 wide + addc, count wide)", k_chain4, 4, bps, nsm, khz);
    run("imad_lo", k_lo, 1, bps, nsm, khz);
    run("imad_hi", k_hi, 1, bps, nsm, khz);
    run("add+xor(2 alu)", k_iadd3, 2, bps, nsm, khz);
    run("add.cc+addc(2 alu)", k_addc, 2, bps, nsm, khz);
    run("wide+1add (count wide)", k_wide_plus_add, 1, bps, nsm, khz);
    run("wide+2alu (count wide)", k_wide_plus_2add, 1, bps, nsm, khz);
    run("dfma", k_dfma, 1, bps, nsm, khz);
  }
  return 0;
}
