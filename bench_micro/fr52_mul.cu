// Throughput of one Montgomery multiplication on B200: 8 x 32-bit limbs on IMAD.WIDE.U32 (csrc/fr.cuh) against
// 5 x 52-bit limbs on DFMA (csrc/fr52.cuh).  Each thread runs CHAINS independent chains x <- x * y.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../gnark_crypto_primitives_b200/csrc -I. -o fr52_mul fr52_mul.cu
// -DF52_EXPERIMENT=1|2|3 builds the timing-only variants (xor instead of add / split 32-bit adds / no DADD).
// The DFMA result of thread 0 is checked on the host with the same header compiled for the CPU (round toward zero).
#include <cstdio>
#include <cstdlib>
#include <cfenv>
#include <algorithm>
#include <cuda_runtime.h>
#include "fr.cuh"
#include "fr52.cuh"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)
#define ITERS 2048

template <int CHAINS>
__global__ void __launch_bounds__(128) k_u32(uint32_t* out, uint32_t seed) {
  uint32_t x[CHAINS][8], y[8];
#pragma unroll
  for (int l = 0; l < 8; l++) {
    y[l] = (seed * 2654435761u + threadIdx.x * 977u + l * 0x9e3779b9u) >> (l == 7 ? 4 : 0);
#pragma unroll
    for (int c = 0; c < CHAINS; c++) x[c][l] = (y[l] ^ (0x1234567u * (c + 1))) >> (l == 7 ? 4 : 0);
  }
#pragma unroll 1
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int c = 0; c < CHAINS; c++) gcp::fr_mul(x[c], x[c], y);
  }
  uint32_t s = 0;
#pragma unroll
  for (int c = 0; c < CHAINS; c++)
#pragma unroll
    for (int l = 0; l < 8; l++) s ^= x[c][l];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__host__ __device__ inline void seed_el(double (&x)[5], uint32_t seed, uint32_t tid, int c) {
  for (int l = 0; l < 5; l++) {
    unsigned long long v = ((unsigned long long)(seed * 2654435761u + tid * 977u + l * 0x9e3779b9u + c * 0x85ebca6bu) << 20) ^
                           (0x5bd1e995ull * (l + 1) * (c + 3));
    v &= (l == 4) ? 0x3fffffffffffull : 0xfffffffffffffull;
    x[l] = (double)v;
  }
}

template <int CHAINS>
__global__ void __launch_bounds__(128) k_f52(double* out, uint32_t seed) {
  double x[CHAINS][5], y[5];
  seed_el(y, seed, threadIdx.x, 7);
#pragma unroll
  for (int c = 0; c < CHAINS; c++) seed_el(x[c], seed, threadIdx.x, c);
#pragma unroll 1
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int c = 0; c < CHAINS; c++) gcp::f52::mul(x[c], x[c], y);
  }
  double* o = out + (size_t)(blockIdx.x * blockDim.x + threadIdx.x) * 5;
#pragma unroll
  for (int l = 0; l < 5; l++) {
    double s = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; c++) s = (c == 0) ? x[c][l] : s;  // chain 0 only (checked on the host)
    o[l] = s;
  }
  if (CHAINS > 1) {  // keep the other chains alive
    double t = 0;
#pragma unroll
    for (int c = 1; c < CHAINS; c++) t += x[c][0];
    if (t == -1.0) o[0] = t;
  }
}

template <typename F>
static float time_it(F launch) {
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int w = 0; w < 2; w++) launch(1u + w);
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < 5; r++) {
    CK(cudaEventRecord(e0)); launch(777u); CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1)); float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); best = std::min(best, ms);
  }
  return best;
}

template <int CHAINS>
static void bench(int nsm, int bps) {
  int grid = nsm * bps, block = 128;
  uint32_t* o32; double* o52;
  CK(cudaMalloc(&o32, (size_t)grid * block * 4));
  CK(cudaMalloc(&o52, (size_t)grid * block * 5 * 8));
  float ms32 = time_it([&](uint32_t s) { k_u32<CHAINS><<<grid, block>>>(o32, s); });
  float ms52 = time_it([&](uint32_t s) { k_f52<CHAINS><<<grid, block>>>(o52, s); });
  double muls = (double)grid * block * ITERS * CHAINS;
  // host check of thread 0, chain 0
  double h[5]; CK(cudaMemcpy(h, o52, sizeof(h), cudaMemcpyDeviceToHost));
  int old = fegetround(); fesetround(FE_TOWARDZERO);
  double x[5], y[5]; seed_el(y, 777u, 0, 7); seed_el(x, 777u, 0, 0);
  for (int it = 0; it < ITERS; it++) gcp::f52::mul(x, x, y);
  fesetround(old);
  bool ok = true; for (int l = 0; l < 5; l++) ok = ok && (x[l] == h[l]);
  printf("{\"chains\": %d, \"blocks_per_sm\": %d, \"imad_gmul_per_s\": %.2f, \"dfma_gmul_per_s\": %.2f, \"ratio\": %.3f, \"dfma_matches_host\": %s}\n",
         CHAINS, bps, muls / ms32 * 1e-6, muls / ms52 * 1e-6, ms32 / ms52, ok ? "true" : "false");
  CK(cudaFree(o32)); CK(cudaFree(o52));
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  printf("{\"device\": \"%s\", \"sms\": %d}\n", p.name, p.multiProcessorCount);
  for (int bps : {2, 4, 6, 8}) {
    bench<1>(p.multiProcessorCount, bps);
    bench<2>(p.multiProcessorCount, bps);
  }
  return 0;
}
