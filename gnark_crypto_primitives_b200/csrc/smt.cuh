// Sparse-Merkle-tree proof verification kernels (value semantics of the circomlib/arbo verifier gadget).
//
// Reference: /root/reference/tree/smt/verifier.go:171-242 (VerifierWithLeafHashFlag) with
// lev_ins.go:43-77, verifier_sm.go:5-14, verifier_level.go:8-17, hash.go:10-27, utils.go:11-56.
//
// For boolean enabled/fnc/isOld0 the gadget's field arithmetic reduces to (derivation in DESIGN.md, checked
// against the literal oracle in tests/):
//   lidx   = 1 + (index of the last non-zero sibling among siblings[0..n-2]), or 0 if none     (LevInsFlag)
//   states : top for levels < lidx; at lidx the machine enters  new (fnc=0) [+ i0 if isOld0],
//            old (fnc=1, isOld0=0) or i0 (fnc=1, isOld0=1);  na afterwards                      (VerifierSM)
//   flagStates  = !(fnc==0 && isOld0==1)                 (exactly one state set at the end)
//   level[lidx] = fnc==0 ? hash1New : (isOld0 ? 0 : hash1Old)
//   level[i]    = H(bit_i(key) ? (sib_i, level[i+1]) : (level[i+1], sib_i))   for i < lidx      (VerifierLevel)
//   flagLevIns  = siblings[n-1] == 0 ;  flagKeyReuse = !(fnc && !isOld0 && oldKey == key)
//   flagRoot    = level[0] == root ;  flag = AND of the four ; enabled==0 => flag = 1, level[0] = 0
// The gadget hashes all n levels but multiplies levels >= lidx by stTop = 0, so they never reach an output;
// this kernel hashes lidx levels.  Assertions of the gadget (key < 2^n, boolean selectors) become status codes.
#pragma once
#include "poseidon.cuh"
#include "kernels.h"

namespace gcp {


__device__ __forceinline__ void load_elem(u32 (&m)[8], bool& canonical, const u32* p, int mont) {
  u32 x[8];
  load_fr(x, p);
  canonical = canonical && fr_is_canonical(x);
  if (mont) {
#pragma unroll
    for (int l = 0; l < 8; l++) m[l] = x[l];
  } else {
    fr_to_mont(m, x);
  }
}

// key as an integer (standard form) -> needed for the path bits and the range assertion
__device__ __forceinline__ void key_integer(u32 (&k)[8], const u32* p, int mont) {
  u32 x[8];
  load_fr(x, p);
  if (mont) {
    fr_from_mont(k, x);
  } else {
#pragma unroll
    for (int l = 0; l < 8; l++) k[l] = x[l];
  }
}

// Pass 1: validate selectors / key range, pick the leaf the state machine will inject and hash it (t = 4).
__global__ void __launch_bounds__(128) smt_leaf_kernel(SmtArgs a) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= a.n) return;
  u32 fnc = a.fnc ? a.fnc[idx] : 0u;
  u32 is0 = a.is_old0 ? a.is_old0[idx] : 0u;
  u32 en = a.enabled ? a.enabled[idx] : 1u;
  u8 st = GCP_STATUS_OK;
  if ((fnc | is0 | en) > 1u) st = GCP_STATUS_NOT_BOOLEAN;

  bool canon = true;
  u32 key[8], val[8], one[8];
  load_elem(key, canon, a.keys + idx * 8, a.mont);
  load_elem(val, canon, a.values + idx * 8, a.mont);
  u32 okey[8], oval[8];
  if (a.old_keys) {
    load_elem(okey, canon, a.old_keys + idx * 8, a.mont);
    load_elem(oval, canon, a.old_values + idx * 8, a.mont);
  }
  {
    u32 r[8];
    load_fr(r, a.roots + idx * a.root_stride);
    canon = canon && fr_is_canonical(r);
  }
  if (!canon && st == GCP_STATUS_OK) st = GCP_STATUS_NONCANONICAL;

  // lowBits(key, n): asserts key < 2^n  (utils.go:11-13)
  if (st == GCP_STATUS_OK) {
    u32 ki[8];
    key_integer(ki, a.keys + idx * 8, a.mont);
    u32 hi = 0;
#pragma unroll
    for (int l = 0; l < 8; l++) {
      int lo_bit = l * 32;
      if (a.n_levels <= lo_bit)
        hi |= ki[l];
      else if (a.n_levels < lo_bit + 32)
        hi |= ki[l] >> (a.n_levels - lo_bit);
    }
    if (hi) st = GCP_STATUS_KEY_RANGE;
  }

  u32 leaf[8];
#pragma unroll
  for (int l = 0; l < 8; l++) {
    leaf[l] = 0;
    one[l] = FR_ONE[l];
  }
  if (st == GCP_STATUS_OK && en == 1u) {
    bool use_new = (fnc == 0u);
    bool use_old = (fnc == 1u && is0 == 0u);
    if (use_new || use_old) {
      if (use_old && a.old_keys) {
#pragma unroll
        for (int l = 0; l < 8; l++) {
          key[l] = okey[l];
          val[l] = oval[l];
        }
      }
      poseidon_hash3(leaf, key, val, one);  // Hash1: H(key, value, 1)
    }
  }
  store_fr(a.leaf + idx * 8, leaf);
  a.status[idx] = st;
}

// Pass 2: walk the sibling array from the leaf end: skip the zero tail, then fold Hash2 up to the root.
__global__ void __launch_bounds__(128) smt_path_kernel(SmtArgs a) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= a.n) return;
  const int n = a.n_levels;
  u8 st = a.status[idx];
  u32 fnc = a.fnc ? a.fnc[idx] : 0u;
  u32 is0 = a.is_old0 ? a.is_old0[idx] : 0u;
  u32 en = a.enabled ? a.enabled[idx] : 1u;
  const u32* sib = a.siblings + idx * (size_t)n * 8;

  u32 acc[8];
  load_fr(acc, a.leaf + idx * 8);
  u32 key[8];
  key_integer(key, a.keys + idx * 8, a.mont);

  bool live = (st == GCP_STATUS_OK) && (en == 1u);
  bool canon = true;
  bool last_zero = true;
  bool started = false;  // true once a non-zero sibling among [0, n-2] has been seen (i < lidx from then on)
  for (int i = n - 1; i >= 0; i--) {
    if (!live) break;
    u32 x[8];
    load_fr(x, sib + (size_t)i * 8);
    canon = canon && fr_is_canonical(x);
    bool nz = !is_zero256(x);
    if (i == n - 1) {
      last_zero = !nz;  // LevInsFlag rule 1; this sibling never takes part in the fold
      continue;
    }
    started = started || nz;
    if (!started) continue;
    u32 s[8];
    if (a.mont) {
#pragma unroll
      for (int l = 0; l < 8; l++) s[l] = x[l];
    } else {
      fr_to_mont(s, x);
    }
    u32 bit = (key[i >> 5] >> (i & 31)) & 1u;
    u32 lft[8], rgt[8];
#pragma unroll
    for (int l = 0; l < 8; l++) {  // Switcher (utils.go:50-56)
      lft[l] = bit ? s[l] : acc[l];
      rgt[l] = bit ? acc[l] : s[l];
    }
    poseidon_hash2(acc, lft, rgt);
  }

  u8 flag = 0;
  u32 root_c[8];
#pragma unroll
  for (int l = 0; l < 8; l++) root_c[l] = 0;
  if (st == GCP_STATUS_OK) {
    if (en == 0u) {
      flag = 1;  // every check bypassed; level[0] = 0
    } else if (!canon) {
      st = GCP_STATUS_NONCANONICAL;
    } else {
      if (a.mont) {
#pragma unroll
        for (int l = 0; l < 8; l++) root_c[l] = acc[l];
        fr_canon(root_c);
      } else {
        fr_from_mont(root_c, acc);
      }
      u32 root[8];
      load_fr(root, a.roots + idx * a.root_stride);
      bool flag_root = eq256(root_c, root);
      bool flag_states = !(fnc == 0u && is0 == 1u);
      bool keys_equal = true;
      if (a.old_keys) {
        u32 ok[8], kk[8];
        load_fr(ok, a.old_keys + idx * 8);
        load_fr(kk, a.keys + idx * 8);
        keys_equal = eq256(ok, kk);
      }
      bool flag_key_reuse = !(fnc == 1u && is0 == 0u && keys_equal);
      flag = (flag_root && flag_states && flag_key_reuse && last_zero) ? 1 : 0;
    }
  }
  a.flags[idx] = (st == GCP_STATUS_OK) ? flag : 0;
  a.status[idx] = st;
  if (a.out_roots) {
    if (st != GCP_STATUS_OK) {
#pragma unroll
      for (int l = 0; l < 8; l++) root_c[l] = 0;
    }
    store_fr(a.out_roots + idx * 8, root_c);
  }
}

}  // namespace gcp
