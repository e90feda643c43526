// Coordinate conversion at the iden3 boundary:
//   FromTEtoRTE / FromRTEtoTE /root/reference/ecc/format/twistededwards.go:29-48 (scalingFactor :17)
// The proof gadgets built on variable-base scalar multiplication (AssertDecrypt, DecryptionProof.Verify, EdDSA) live in
// varbase.cuh.
#pragma once
#include "edwards.cuh"
#include "elgamal.cuh"

namespace gcp {

// ---- TE <-> RTE: x' = x * (-f) or x / (-f), y unchanged (ecc/format/twistededwards.go:29-48) -------------------------
__global__ void te_rte_kernel(const u32* __restrict__ in, size_t n_points, u32* __restrict__ out, u8* __restrict__ status,
                              int to_rte) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_points) return;
  const u32 negf[8] = {0xc9603c7bu, 0x5c62c8e0u, 0x8fabc7f1u, 0xf8382911u, 0x6aa07f4du, 0x7d53da81u, 0x6ba06ab6u, 0x1da7c5b3u};
  const u32 negf_inv[8] = {0xb1b017d8u, 0x61d380bfu, 0x8415d72eu, 0x7f5d8063u, 0x294f7a18u, 0x77e18e30u, 0x305733c2u, 0x10d2ede5u};
  u32 x[8], y[8], r[8];
  load_fr(x, in + idx * 16);
  load_fr(y, in + idx * 16 + 8);
  bool canon = fr_is_canonical(x) && fr_is_canonical(y);
  // mont_mul(x, c*R) = x*c in whichever representation x is in (standard or Montgomery)
  if (to_rte)
    fr_mul(r, x, negf);
  else
    fr_mul(r, x, negf_inv);
  fr_canon(r);
  if (!canon) {
    fr_set_zero(r);
    fr_set_zero(y);
  }
  store_fr(out + idx * 16, r);
  store_fr(out + idx * 16 + 8, y);
  status[idx] = canon ? GCP_STATUS_OK : GCP_STATUS_NONCANONICAL;
}

}  // namespace gcp
