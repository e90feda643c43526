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
  const u32 negf[8] = GCP_NEG_F_MONT;
  const u32 negf_inv[8] = GCP_NEG_F_INV_MONT;
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
