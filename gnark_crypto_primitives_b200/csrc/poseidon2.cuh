// Width-2 Poseidon2 over BN254 Fr (t = 2, rF = 6, rP = 50, x^5) and the Merkle-Damgard hasher the reference builds on
// it: values of HashPoseidon2.Hash, /root/reference/hash/native/bn254/poseidon2/native.go:30-63 (perm2 :27), and of
// the gadget HashPoseidon2Gnark, gnark.go:18-54 (MinMaxHint hints.go:10-19).  The permutation itself is gnark-crypto's
// ecc/bn254/fr/poseidon2 (un-vendored): external matrix circ(2,1) before the first round and after every full round,
// internal matrix [[2,1],[1,3]] after every partial round, round = matmul(sbox(state + key)).  The 62 round keys are
// DATA (one array per context, installed from data/poseidon2_bn254_t2.bin or by gcp_poseidon2_set_round_keys), kept in
// Montgomery form in global memory: every thread of a warp reads the same key, so the loads are uniform and L1-resident.
// One thread per item; 62 S-boxes (124 squarings + 62 multiplies) per permutation, no matrix multiplications.
#pragma once
#include "fr.cuh"
#include "kernels.h"

namespace gcp {

constexpr int P2_RF = 6;
constexpr int P2_RP = 50;
constexpr int P2_KEYS = P2_RF * 2 + P2_RP;  // 62, flattened in round order

__device__ __forceinline__ void p2_load_key(u32 (&k)[8], const u32* __restrict__ keys, int idx) {
  load_fr(k, keys + (size_t)idx * 8);
}

__device__ __forceinline__ void p2_sbox(u32 (&x)[8]) {
  u32 x2[8], x4[8];
  fr_sqr(x2, x);
  fr_sqr(x4, x2);
  fr_mul(x, x4, x);
}

__device__ __forceinline__ void p2_external(u32 (&s0)[8], u32 (&s1)[8]) {
  u32 t[8];
  fr_add(t, s0, s1);
  fr_add(s0, s0, t);
  fr_add(s1, s1, t);
}

// state in lazy Montgomery form (< 2r) in and out
__device__ __noinline__ void poseidon2_permute(u32 (&s0)[8], u32 (&s1)[8], const u32* __restrict__ keys) {
  p2_external(s0, s1);
  int kp = 0;
#pragma unroll 1
  for (int i = 0; i < P2_RF + P2_RP; i++) {
    const bool full = i < P2_RF / 2 || i >= P2_RF / 2 + P2_RP;
    u32 k[8];
    p2_load_key(k, keys, kp++);
    fr_add(s0, s0, k);
    p2_sbox(s0);
    if (full) {  // uniform over the warp
      p2_load_key(k, keys, kp++);
      fr_add(s1, s1, k);
      p2_sbox(s1);
      p2_external(s0, s1);
    } else {
      u32 t[8];
      fr_add(t, s0, s1);
      fr_add(s0, s0, t);
      fr_add(s1, s1, s1);
      fr_add(s1, s1, t);
    }
  }
}

// element in the caller's format -> lazy Montgomery; `std` receives the canonical integer (needed for the min/max order).
// Canonical format: a 32-byte value >= r is reduced mod r like native.go:37-39 (SafeBigInt) does.  Montgomery format is
// gnark-crypto's fr.Element memory, whose invariant is < r: a value >= r there is status NONCANONICAL.
__device__ __forceinline__ bool p2_ingest(u32 (&mont)[8], u32 (&std)[8], const u32* p, int is_mont) {
  u32 raw[8];
  load_fr(raw, p);
  if (is_mont) {
    if (!fr_is_canonical(raw)) return false;
    fr_copy(mont, raw);
    fr_from_mont(std, raw);
  } else {
    const u32 P1[8] = GCP_P_LIMBS;
#pragma unroll 1
    for (int it = 0; it < 5 && !fr_is_canonical(raw); it++) {  // 2^256 / r < 5.3
      u32 t[8];
      sub256(t, raw, P1);
      fr_copy(raw, t);
    }
    fr_copy(std, raw);
    fr_to_mont(mont, raw);
  }
  return true;
}

__device__ __forceinline__ bool lt256(const u32 (&a)[8], const u32 (&b)[8]) {
  u32 t[8];
  return sub256(t, a, b) != 0;
}

__device__ __forceinline__ void p2_emit(u32* out, u32 (&x)[8], int is_mont, bool ok) {
  u32 res[8];
  if (is_mont) {
    fr_copy(res, x);
    fr_canon(res);
  } else {
    fr_from_mont(res, x);
  }
  if (!ok) fr_set_zero(res);
  store_fr(out, res);
}

// HashPoseidon2.Hash: len = 2 (internal node, ordered min/max) or 3 (leaf: key, value, flag)
__global__ void __launch_bounds__(128) poseidon2_hash_kernel(const u32* __restrict__ keys, const u32* __restrict__ in, int len,
                                                             size_t n, u32* __restrict__ out, u8* __restrict__ status,
                                                             int is_mont) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  const u32* row = in + idx * (size_t)len * 8;
  bool ok = true;
  u32 m0[8], m1[8], m2[8], a_std[8], b_std[8];
  ok = p2_ingest(m0, a_std, row, is_mont) && ok;
  ok = p2_ingest(m1, b_std, row + 8, is_mont) && ok;
  if (len == 3) {
    u32 c_std[8];
    ok = p2_ingest(m2, c_std, row + 16, is_mont) && ok;
  } else if (lt256(b_std, a_std)) {  // native.go:42-44: bytes.Compare(safe[0], safe[1]) > 0 => swap
#pragma unroll
    for (int l = 0; l < 8; l++) {
      u32 t = m0[l];
      m0[l] = m1[l];
      m1[l] = t;
    }
  }
  u32 cv[8], s1[8];
  fr_set_zero(cv);
#pragma unroll 1
  for (int j = 0; j < len; j++) {  // native.go:47-61: st = {cv, m}; Permutation; cv = st[1] + m
    u32 m[8];
#pragma unroll
    for (int l = 0; l < 8; l++) m[l] = j == 0 ? m0[l] : (j == 1 ? m1[l] : m2[l]);
    fr_copy(s1, m);
    poseidon2_permute(cv, s1, keys);
    fr_add(cv, s1, m);
  }
  p2_emit(out + idx * 8, cv, is_mont, ok);
  status[idx] = ok ? GCP_STATUS_OK : GCP_STATUS_NONCANONICAL;
}

// perm2.Permutation(st[:]) on n states of two elements (native.go:55)
__global__ void __launch_bounds__(128) poseidon2_permutation_kernel(const u32* __restrict__ keys, const u32* __restrict__ in,
                                                                    size_t n, u32* __restrict__ out, u8* __restrict__ status,
                                                                    int is_mont) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  u32 s0[8], s1[8], t[8];
  bool ok = p2_ingest(s0, t, in + idx * 16, is_mont);
  ok = p2_ingest(s1, t, in + idx * 16 + 8, is_mont) && ok;
  poseidon2_permute(s0, s1, keys);
  p2_emit(out + idx * 16, s0, is_mont, ok);
  p2_emit(out + idx * 16 + 8, s1, is_mont, ok);
  status[idx] = ok ? GCP_STATUS_OK : GCP_STATUS_NONCANONICAL;
}

}  // namespace gcp
