// iden3-compatible MiMC7 over BN254 Fr: 91 rounds of x <- (x + h + c_i)^7, Miyaguchi-Preneel chaining with field
// addition.  Values of (*MiMC).Sum, /root/reference/hash/native/bn254/mimc7/mimc.go:47-54 (encrypt :80-87,
// pow7 :74-78, at most 62 inputs :9,33-38), constants.go:9-25 (constants[0] = 0).  One thread per hash.
#pragma once
#include "fr.cuh"
#include "kernels.h"

namespace gcp {

constexpr int MIMC7_ROUNDS = 91;
constexpr int MIMC7_MAX_INPUTS = 62;
__device__ __constant__ u32 c_mimc7[MIMC7_ROUNDS * 8];  // Montgomery form

__global__ void __launch_bounds__(128) mimc7_kernel(const u32* __restrict__ in, int len, size_t n, u32* __restrict__ out,
                                                    u8* __restrict__ status, int mont) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  bool canon = true;
  u32 h[8];
  fr_set_zero(h);
#pragma unroll 1
  for (int j = 0; j < len; j++) {
    u32 raw[8], d[8], x[8];
    load_fr(raw, in + (idx * (size_t)len + j) * 8);
    canon = canon && fr_is_canonical(raw);
    if (mont)
      fr_copy(d, raw);
    else
      fr_to_mont(d, raw);
    fr_copy(x, d);
#pragma unroll 1
    for (int i = 0; i < MIMC7_ROUNDS; i++) {
      u32 c[8], s[8], s2[8], s4[8], s3[8];
#pragma unroll
      for (int l = 0; l < 8; l++) c[l] = c_mimc7[i * 8 + l];
      fr_add(s, x, h);
      fr_add(s, s, c);
      fr_sqr(s2, s);
      fr_sqr(s4, s2);
      fr_mul(s3, s, s2);
      fr_mul(x, s3, s4);  // s^7 = s * s^2 * s^4
    }
    // r = x + h ; h = h + r + d   (mimc.go:50-51, :86)
    fr_add(x, x, h);
    fr_add(h, h, x);
    fr_add(h, h, d);
  }
  u32 res[8];
  if (mont) {
    fr_copy(res, h);
    fr_canon(res);
  } else {
    fr_from_mont(res, h);
  }
  if (!canon) fr_set_zero(res);
  store_fr(out + idx * 8, res);
  status[idx] = canon ? GCP_STATUS_OK : GCP_STATUS_NONCANONICAL;
}

}  // namespace gcp
