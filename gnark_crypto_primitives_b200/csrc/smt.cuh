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
#include "poseidon2.cuh"
#include "kernels.h"

namespace gcp {


// The tree's hash function is a plug in the reference (utils.Hasher, utils/hashers.go:10; every gadget of tree/smt takes
// hFn).  HASHER 0: PoseidonHasher (:25-27), the hash of arbo's circomlib-compatible trees and the default;
// HASHER 1: Poseidon2Hasher (:35-37) = HashPoseidon2Gnark (hash/native/bn254/poseidon2/gnark.go:18-54), the width-2
// Merkle-Damgard hasher: a node hashes its two children ordered (min, max) (the MinMax hint, :34-44), a leaf hashes
// (key, value, 1) in that order; its round keys are the context's (poseidon2.cuh; permutation parity unpinned).
__device__ __forceinline__ void p2_absorb(u32 (&cv)[8], const u32 (&m)[8], const u32* __restrict__ hk) {
  u32 s1[8];
  fr_copy(s1, m);
  poseidon2_permute(cv, s1, hk);  // native.go:47-61: st = {cv, m}; Permutation; cv = st[1] + m
  fr_add(cv, s1, m);
}

// Hash2 of the tree kernels: the pair-schedule permutation (poseidon.cuh: 59 784 instead of 61 384 wide multiplies) as an
// OUT-OF-LINE function whose operands and result are structs passed by value (registers).  Inlined into smt_path_kernel the
// same body lost 1.7 % (capped at 128 registers ptxas spends 6 k more non-multiply instructions per hash on it, ncu:
// profiles/r02_smt_pair_schedule_ncu.csv; uncapped it takes 158 registers, three blocks per SM); as a function it keeps its
// own register allocation - the one poseidon_fixed_kernel<3> gets - and the caller's per-level state is saved around one
// call per level (140 B of stack traffic per 60 k multiplies).  Dense proofs 786 -> 797 k/s at 2^17, census-like
// 5.13 -> 5.22 M/s, processor 2.48 -> 2.57 M transitions/s.  GCP_SMT_HASH2_INLINE restores the inlined one-round form.
#ifndef GCP_SMT_HASH2_INLINE
struct SmtE8 {
  u32 v[8];
};
__device__ __noinline__ SmtE8 poseidon_hash2_pairs_ool(SmtE8 l, SmtE8 r) {
  u32 s[3][8];
#pragma unroll
  for (int i = 0; i < 8; i++) {
    s[0][i] = 0;
    s[1][i] = l.v[i];
    s[2][i] = r.v[i];
  }
  SmtE8 o;
  poseidon_permute_const<3, true>(s, o.v);
  return o;
}
#endif
template <int HASHER>
__device__ __forceinline__ void smt_hash2(u32 (&out)[8], const u32 (&l)[8], const u32 (&r)[8], const u32* __restrict__ hk) {
  if constexpr (HASHER == 0) {
#ifndef GCP_SMT_HASH2_INLINE
    SmtE8 a, b;
    fr_copy(a.v, l);
    fr_copy(b.v, r);
    a = poseidon_hash2_pairs_ool(a, b);
    fr_copy(out, a.v);
#else
    poseidon_hash2(out, l, r);
#endif
  } else {
    u32 ls[8], rs[8];
    fr_from_mont(ls, l);  // canonical integers: the order is on the values, not on their representations
    fr_from_mont(rs, r);
    const bool swap = lt256(rs, ls);
    u32 cv[8], m[8];
    fr_set_zero(cv);
#pragma unroll
    for (int i = 0; i < 8; i++) m[i] = swap ? r[i] : l[i];
    p2_absorb(cv, m, hk);
#pragma unroll
    for (int i = 0; i < 8; i++) m[i] = swap ? l[i] : r[i];
    p2_absorb(cv, m, hk);
    fr_copy(out, cv);
  }
}

template <int HASHER>
__device__ __forceinline__ void smt_hash1(u32 (&out)[8], const u32 (&key)[8], const u32 (&val)[8], const u32 (&one)[8],
                                          const u32* __restrict__ hk) {
  if constexpr (HASHER == 0) {
    poseidon_hash3(out, key, val, one);
  } else {
    u32 cv[8];
    fr_set_zero(cv);
    p2_absorb(cv, key, hk);
    p2_absorb(cv, val, hk);
    p2_absorb(cv, one, hk);
    fr_copy(out, cv);
  }
}

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
template <int HASHER>
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
      if (a.leaf_hash_form) {
        fr_copy(leaf, val);                   // the caller's hash1New / hash1Old (lazy Montgomery after load_elem)
      } else {
        smt_hash1<HASHER>(leaf, key, val, one, a.hkeys);  // Hash1: H(key, value, 1)
      }
    }
  }
  store_fr(a.leaf + idx * 8, leaf);
  a.status[idx] = st;
}

// ---------------------------------------------------------------------------------------------------------
// Pass 0 (HBM-bound): stream the whole sibling array once and reduce every proof to
//   lidx (1 + index of the last non-zero sibling among [0, n-2]), "siblings[n-1] == 0", "every sibling < r".
// One warp per proof, lane = one 16-byte piece: a warp-wide LDG.128 covers 512 contiguous bytes (four whole
// 128-byte lines), all loads of a proof are issued before the first use.  This is the proof-streaming kernel whose
// achieved GB/s is reported against the measured HBM copy bandwidth.
// ---------------------------------------------------------------------------------------------------------
constexpr int SMT_SCAN_MAX_ITERS = 16;  // 253 levels * 2 pieces / 32 lanes, rounded up
constexpr u32 SMT_INFO_LAST_ZERO = 1u, SMT_INFO_CANONICAL = 2u;

__global__ void __launch_bounds__(256) smt_scan_kernel(const u32* __restrict__ siblings, size_t n, int n_levels,
                                                       u16* __restrict__ lidx_out, u8* __restrict__ info_out,
                                                       u32* __restrict__ hist) {
  const int lane = threadIdx.x & 31;
  const size_t warp_global = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const size_t n_warps = ((size_t)gridDim.x * blockDim.x) >> 5;
  const int pieces = n_levels * 2;                      // 16-byte pieces per proof
  const int iters = (pieces + 31) >> 5;
  // r as two 128-bit halves, most significant limb last
  const u32 r_lo[4] = {GCP_P0, GCP_P1, GCP_P2, GCP_P3}, r_hi[4] = {GCP_P4, GCP_P5, GCP_P6, GCP_P7};
  for (size_t proof = warp_global; proof < n; proof += n_warps) {
    const uint4* base = reinterpret_cast<const uint4*>(siblings + proof * (size_t)n_levels * 8);
    uint4 v[SMT_SCAN_MAX_ITERS];
#pragma unroll
    for (int it = 0; it < SMT_SCAN_MAX_ITERS; it++) {
      int piece = it * 32 + lane;
      v[it] = (it < iters && piece < pieces) ? __ldcs(base + piece) : make_uint4(0, 0, 0, 0);
    }
    int last_nz = -1;          // highest non-zero sibling index among [0, n-2]
    bool last_zero = true, canon = true;
    const unsigned full = 0xffffffffu;
    const bool high = (lane & 1) != 0;                  // odd lanes hold limbs 4..7 of a sibling
#pragma unroll
    for (int it = 0; it < SMT_SCAN_MAX_ITERS; it++) {
      if (it < iters) {
        const uint4 x = v[it];
        const bool nz = (x.x | x.y | x.z | x.w) != 0;
        // bit p of m <-> piece p non-zero; sibling k of this iteration owns pieces 2k, 2k+1
        const unsigned m = __ballot_sync(full, nz);
        unsigned nz_mask = (m | (m >> 1)) & 0x55555555u;
        const int sib_in_iter = min(16, n_levels - it * 16);
        if (sib_in_iter < 16) nz_mask &= (1u << (2 * sib_in_iter)) - 1u;
        // canonical: only a top limb >= the top limb of r can make a sibling >= r; the full compare runs in that rare case
        const unsigned suspects = __ballot_sync(full, high && x.w >= GCP_P7);
        if (suspects) {
          const u32 w[4] = {x.x, x.y, x.z, x.w};
          const u32* ref = high ? r_hi : r_lo;
          bool lt = false, eq = true;
#pragma unroll
          for (int l = 3; l >= 0; l--) {
            lt = lt || (eq && w[l] < ref[l]);
            eq = eq && (w[l] == ref[l]);
          }
          bool o_lt = __shfl_xor_sync(full, lt, 1);
          bool o_eq = __shfl_xor_sync(full, eq, 1);
          bool sib_lt = o_lt || (o_eq && lt);           // on even lanes: hi < r_hi, or hi == r_hi and lo < r_lo
          bool valid = !high && (it * 16 + (lane >> 1)) < n_levels;
          if (__ballot_sync(full, valid && !sib_lt)) canon = false;
        }
        if (it * 16 + 15 >= n_levels - 1) {             // this iteration contains sibling n-1
          int k = (n_levels - 1) - it * 16;
          if (k >= 0 && k < 16) {
            if (nz_mask & (1u << (2 * k))) last_zero = false;
            nz_mask &= ~(1u << (2 * k));
          }
        }
        if (nz_mask) last_nz = it * 16 + ((31 - __clz(nz_mask)) >> 1);
      }
    }
    if (lane == 0) {
      int lidx = last_nz + 1;
      lidx_out[proof] = (u16)lidx;
      info_out[proof] = (u8)((last_zero ? SMT_INFO_LAST_ZERO : 0u) | (canon ? SMT_INFO_CANONICAL : 0u));
      if (hist) atomicAdd(hist + lidx, 1u);
    }
  }
}

// Counting sort of proof indices by lidx, longest paths first: exclusive prefix over 256 bins (one block), then scatter.
__global__ void smt_sort_prefix_kernel(const u32* __restrict__ hist, u32* __restrict__ cursor, u32 n) {
  __shared__ u32 sh[256];
  int t = threadIdx.x;
  sh[t] = hist[255 - t];                                // descending lidx
  __syncthreads();
  if (t == 0) {
    u32 acc = 0, uniform = 0;
    for (int i = 0; i < 256; i++) {
      u32 c = sh[i];
      if (c == n) uniform = 1;                          // every proof has the same path length: keep memory order
      sh[i] = acc;
      acc += c;
    }
    cursor[256] = uniform;
  }
  __syncthreads();
  cursor[255 - t] = sh[t];
}

__global__ void smt_sort_scatter_kernel(const u16* __restrict__ lidx, size_t n, u32* __restrict__ cursor, u32* __restrict__ perm) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  if (cursor[256]) {
    perm[idx] = (u32)idx;
    return;
  }
  u32 pos = atomicAdd(cursor + lidx[idx], 1u);
  perm[pos] = (u32)idx;
}

// Pass 2: fold Hash2 from level lidx-1 up to the root.
//
// Proofs are visited through `perm` (sorted by lidx), so the 32 proofs of a warp have (nearly) equal path lengths:
// no divergence over the path length and no reads of the zero tail.  Sibling staging: the proof-major array gives
// every proof SMT_CH consecutive levels as one contiguous, 128-byte aligned run, so the warp streams a chunk of SMT_CH
// levels for all its 32 proofs with 8 cp.async instructions of 32 x 16 B, each covering four whole 128-byte lines
// (vectorised, every fetched sector used), into a double-buffered shared-memory tile; rows are padded to 144 B so that
// each lane's LDS.128 reads of its own row are bank-conflict free.  The next chunk is in flight while the current one
// is hashed, so HBM latency never reaches the integer pipe.
constexpr int SMT_CH = 4;                       // levels per staged chunk: 4 x 32 B = one 128-byte line per proof
constexpr int SMT_ROW_WORDS = SMT_CH * 8 + 4;   // 36 words = 144 B row stride
constexpr int SMT_WARPS = 4;                    // 128 threads per block (64 and 256 measured slower / do not fit static smem)

__device__ __forceinline__ void smt_stage_chunk(u32* tile, const u32* __restrict__ siblings, u32 my_proof, bool my_valid,
                                                int n_levels, int chunk, int lane) {
  const int part = lane & 7;                    // 16-byte piece of the 128-byte run
  const int level = chunk * SMT_CH + (part >> 1);
#pragma unroll
  for (int t = 0; t < 8; t++) {
    const int p = t * 4 + (lane >> 3);          // slot within the warp
    const u32 proof = __shfl_sync(0xffffffffu, my_proof, p);
    const bool valid = __shfl_sync(0xffffffffu, my_valid, p);
    u32* dst = tile + p * SMT_ROW_WORDS + part * 4;
    if (valid && level < n_levels) {
      const u32* src = siblings + ((size_t)proof * (size_t)n_levels + (size_t)chunk * SMT_CH) * 8 + part * 4;
      unsigned saddr = (unsigned)__cvta_generic_to_shared(dst);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(saddr), "l"(src) : "memory");
    } else {
      *reinterpret_cast<uint4*>(dst) = make_uint4(0, 0, 0, 0);
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
}

// four blocks per SM (128 registers).  With the inlined Hash2 (GCP_SMT_HASH2_INLINE) ptxas lands on 128 by itself and a
// bound of 1 would make it take 136.
#ifndef GCP_SMT_PATH_MIN_BLOCKS
#define GCP_SMT_PATH_MIN_BLOCKS 4
#endif
#define GCP_SMT_PATH_BOUNDS __launch_bounds__(SMT_WARPS * 32, GCP_SMT_PATH_MIN_BLOCKS)
template <int HASHER>
__global__ void GCP_SMT_PATH_BOUNDS smt_path_kernel(SmtArgs a, const u32* __restrict__ perm, const u16* __restrict__ lidx_arr,
                                                       const u8* __restrict__ info_arr) {
  __shared__ __align__(16) u32 tiles[2][SMT_WARPS][32 * SMT_ROW_WORDS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const size_t warp_base = (size_t)blockIdx.x * blockDim.x + warp * 32;
  if (warp_base >= a.n) return;                 // whole warp out of range
  const bool in_range = warp_base + lane < a.n;
  const size_t idx = in_range ? perm[warp_base + lane] : 0;
  const int n = a.n_levels;
  u8 st = in_range ? a.status[idx] : (u8)GCP_STATUS_NONCANONICAL;
  u32 fnc = a.fnc ? a.fnc[idx] : 0u;
  u32 is0 = a.is_old0 ? a.is_old0[idx] : 0u;
  u32 en = a.enabled ? a.enabled[idx] : 1u;
  const u32 info = in_range ? info_arr[idx] : 0u;
  const int lidx = in_range ? (int)lidx_arr[idx] : 0;

  u32 acc[8];
  load_fr(acc, a.leaf + idx * 8);
  u32 key[8];
  key_integer(key, a.keys + idx * 8, a.mont);

  const bool canon = (info & SMT_INFO_CANONICAL) != 0;
  const bool last_zero = (info & SMT_INFO_LAST_ZERO) != 0;
  const bool live = in_range && (st == GCP_STATUS_OK) && (en == 1u) && canon;
  // the warp walks chunks from its longest live path downwards
  int my_top = live ? lidx : 0;
  int warp_top = my_top;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) warp_top = max(warp_top, __shfl_xor_sync(0xffffffffu, warp_top, o));
  const int first_chunk = (warp_top + SMT_CH - 1) / SMT_CH - 1;   // chunk holding level warp_top - 1 (or -1: nothing to hash)
  if (first_chunk >= 0) smt_stage_chunk(tiles[first_chunk & 1][warp], a.siblings, (u32)idx, live, n, first_chunk, lane);
#pragma unroll 1
  for (int c = first_chunk; c >= 0; c--) {
    if (c > 0) {
      smt_stage_chunk(tiles[(c - 1) & 1][warp], a.siblings, (u32)idx, live, n, c - 1, lane);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncwarp();
    const u32* row = tiles[c & 1][warp] + lane * SMT_ROW_WORDS;
    const int hi = min(n, c * SMT_CH + SMT_CH) - 1;
#pragma unroll 1
    for (int i = hi; i >= c * SMT_CH; i--) {
      if (!live || i >= lidx) continue;
      u32 x[8];
      {
        const uint4* q = reinterpret_cast<const uint4*>(row + (i - c * SMT_CH) * 8);
        uint4 v0 = q[0], v1 = q[1];
        x[0] = v0.x; x[1] = v0.y; x[2] = v0.z; x[3] = v0.w;
        x[4] = v1.x; x[5] = v1.y; x[6] = v1.z; x[7] = v1.w;
      }
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
      smt_hash2<HASHER>(acc, lft, rgt, a.hkeys);
    }
    __syncwarp();  // everyone is done with this buffer before it is refilled two chunks later
  }
  if (!in_range) return;

  u8 flag = 0;
  u32 root_c[8];
#pragma unroll
  for (int l = 0; l < 8; l++) root_c[l] = 0;
  if (st == GCP_STATUS_OK) {
    if (!canon) {
      st = GCP_STATUS_NONCANONICAL;  // a sibling >= r is not a field element, whatever the selectors say
    } else if (en == 0u) {
      flag = 1;  // every check bypassed; level[0] = 0
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

// ---------------------------------------------------------------------------------------------------------
// Processor: state transition (insert / update / delete / nop), values of smt.Processor
// (/root/reference/tree/smt/processor.go:10-72, processor_sm.go:7-17, processor_level.go:10-27).
//
// For boolean selectors the six-state machine reduces to (checked against the literal oracle in tests/):
//   enabled = fnc0 | fnc1 ;  lidx as in the verifier ;  LevIns asserts siblings[n-1] == 0 when enabled
//   levels < lidx            : top   -> old_i = H(sw(b_i; old_{i+1}, sib_i)),  new_i = H(sw(b_i; new_{i+1}, sib_i))
//   level lidx, fnc0 = 0     : upd   -> old = hash1Old, new = hash1New                        (update)
//   level lidx, fnc0 = 1, isOld0 : old0 -> old = 0, new = hash1New                          (insert into an empty slot)
//   level lidx.., fnc0 = 1, !isOld0 : bot while the old and new key bits agree (old = hash1Old, new = H(sw(b_i; new_{i+1}, 0))),
//                              new1 at the first level j where they differ (old = hash1Old, new = H(sw(b_j; hash1New, hash1Old)))
//   b_i = bit i of the NEW key.  The final-state assertion fails when no terminal state is reached (keys equal).
//   topL/topR = fnc0 & fnc1 ? (new_0, old_0) : (old_0, new_0) ;  assert oldRoot == topL ;  newRoot = enabled ? topR : oldRoot
//   assert !(!fnc0 & fnc1 & oldKey != newKey)
// Assertion failures are reported as status GCP_STATUS_ASSERTION with newRoot = 0.
// ---------------------------------------------------------------------------------------------------------

__device__ __forceinline__ bool key_in_range(const u32 (&ki)[8], int n_levels) {
  u32 hi = 0;
#pragma unroll
  for (int l = 0; l < 8; l++) {
    int lo_bit = l * 32;
    if (n_levels <= lo_bit)
      hi |= ki[l];
    else if (n_levels < lo_bit + 32)
      hi |= ki[l] >> (n_levels - lo_bit);
  }
  return hi == 0;
}

template <int HASHER>
__global__ void __launch_bounds__(128) smt_process_kernel(SmtProcessArgs a) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= a.n) return;
  const int n = a.n_levels;
  const u32 fnc0 = a.fnc0[idx], fnc1 = a.fnc1[idx], is0 = a.is_old0[idx];
  u8 st = GCP_STATUS_OK;
  if ((fnc0 | fnc1 | is0) > 1u) st = GCP_STATUS_NOT_BOOLEAN;
  const bool enabled = (fnc0 | fnc1) != 0;

  bool canon = true;
  u32 okey[8], oval[8], nkey[8], nval[8], oroot[8];
  load_elem(okey, canon, a.old_keys + idx * 8, a.mont);
  load_elem(oval, canon, a.old_values + idx * 8, a.mont);
  load_elem(nkey, canon, a.new_keys + idx * 8, a.mont);
  load_elem(nval, canon, a.new_values + idx * 8, a.mont);
  load_fr(oroot, a.old_roots + idx * 8);
  canon = canon && fr_is_canonical(oroot);
  const u32* sib = a.siblings + idx * (size_t)n * 8;
  u32 ok_int[8], nk_int[8];
  key_integer(ok_int, a.old_keys + idx * 8, a.mont);
  key_integer(nk_int, a.new_keys + idx * 8, a.mont);

  // sibling scan from the leaf end: canonical check, siblings[n-1] == 0, lidx
  bool last_zero = true;
  int lidx = 0;
  {
    bool found = false;
    for (int i = n - 1; i >= 0; i--) {
      u32 x[8];
      load_fr(x, sib + (size_t)i * 8);
      canon = canon && fr_is_canonical(x);
      bool nz = !is_zero256(x);
      if (i == n - 1) {
        last_zero = !nz;
      } else if (nz && !found) {
        found = true;
        lidx = i + 1;
      }
    }
  }
  if (st == GCP_STATUS_OK && !canon) st = GCP_STATUS_NONCANONICAL;
  if (st == GCP_STATUS_OK && !(key_in_range(ok_int, n) && key_in_range(nk_int, n))) st = GCP_STATUS_KEY_RANGE;
  if (st == GCP_STATUS_OK && enabled && !last_zero) st = GCP_STATUS_ASSERTION;  // LevIns (lev_ins.go:16-20)
  bool keys_equal = eq256(ok_int, nk_int);
  if (st == GCP_STATUS_OK && fnc0 == 0u && fnc1 == 1u && !keys_equal) st = GCP_STATUS_ASSERTION;  // processor.go:64-70

  // terminal level of the state machine
  int jterm = lidx;
  if (st == GCP_STATUS_OK && enabled && fnc0 == 1u && is0 == 0u) {
    jterm = -1;
    for (int i = lidx; i < n; i++) {
      u32 x = ((ok_int[i >> 5] ^ nk_int[i >> 5]) >> (i & 31)) & 1u;
      if (x) {
        jterm = i;
        break;
      }
    }
    if (jterm < 0) st = GCP_STATUS_ASSERTION;  // no terminal state: processor.go:47
  }

  u32 out[8];
#pragma unroll
  for (int l = 0; l < 8; l++) out[l] = 0;
  if (st == GCP_STATUS_OK) {
    if (!enabled) {
      load_fr(out, a.old_roots + idx * 8);  // nop: newRoot = oldRoot
    } else {
      // leaf hashes (one inlined copy of the t = 4 permutation), or the caller's (ProcessorWithLeafHash)
      u32 h1old[8], h1new[8], one[8];
#pragma unroll
      for (int l = 0; l < 8; l++) one[l] = FR_ONE[l];
      if (a.leaf_hash_form) {
        fr_copy(h1old, oval);
        fr_copy(h1new, nval);
      }
#pragma unroll 1
      for (int h = 0; h < 2 && !a.leaf_hash_form; h++) {
        u32 kk[8], vv[8], res[8];
#pragma unroll
        for (int l = 0; l < 8; l++) {
          kk[l] = h ? nkey[l] : okey[l];
          vv[l] = h ? nval[l] : oval[l];
        }
        smt_hash1<HASHER>(res, kk, vv, one, a.hkeys);
#pragma unroll
        for (int l = 0; l < 8; l++) {
          if (h)
            h1new[l] = res[l];
          else
            h1old[l] = res[l];
        }
      }
      const bool is_upd = (fnc0 == 0u), is_old0 = (fnc0 == 1u && is0 == 1u);
      u32 acc_old[8], acc_new[8], zero[8];
#pragma unroll
      for (int l = 0; l < 8; l++) zero[l] = 0;
      // terminal level
      if (is_upd) {
        fr_copy(acc_old, h1old);
        fr_copy(acc_new, h1new);
      } else if (is_old0) {
        fr_copy(acc_old, zero);
        fr_copy(acc_new, h1new);
      } else {
        fr_copy(acc_old, h1old);
        fr_copy(acc_new, h1new);  // becomes H(sw(b_j; hash1New, hash1Old)) in the first loop iteration below
      }
      // walk up: i = jterm (new1 only), jterm-1 .. lidx (bot), lidx-1 .. 0 (top); one inlined copy of Hash2
      const int start = (is_upd || is_old0) ? lidx - 1 : jterm;
      for (int i = start; i >= 0; i--) {
        const u32 bit = (nk_int[i >> 5] >> (i & 31)) & 1u;
        const bool top = i < lidx;
        u32 s[8];
        if (top) {
          u32 x[8];
          load_fr(x, sib + (size_t)i * 8);
          if (a.mont)
            fr_copy(s, x);
          else
            fr_to_mont(s, x);
        }
#pragma unroll 1
        for (int h = 0; h < 2; h++) {
          if (h == 0 && !top) continue;  // the old chain only hashes on top levels
          u32 child[8], other[8];
          if (h == 0) {
            fr_copy(child, acc_old);
            fr_copy(other, s);
          } else {
            fr_copy(child, acc_new);
            if (top)
              fr_copy(other, s);
            else if (i == jterm)
              fr_copy(other, h1old);
            else
              fr_copy(other, zero);
          }
          u32 lft[8], rgt[8], res[8];
#pragma unroll
          for (int l = 0; l < 8; l++) {
            lft[l] = bit ? other[l] : child[l];
            rgt[l] = bit ? child[l] : other[l];
          }
          smt_hash2<HASHER>(res, lft, rgt, a.hkeys);
          if (h == 0)
            fr_copy(acc_old, res);
          else
            fr_copy(acc_new, res);
        }
      }
      u32 old_c[8], new_c[8];
      if (a.mont) {
        fr_copy(old_c, acc_old);
        fr_copy(new_c, acc_new);
        fr_canon(old_c);
        fr_canon(new_c);
      } else {
        fr_from_mont(old_c, acc_old);
        fr_from_mont(new_c, acc_new);
      }
      const bool both = (fnc0 & fnc1) != 0;
      bool match = both ? eq256(new_c, oroot) : eq256(old_c, oroot);  // ForceEqualIfEnabled, processor.go:60
      if (!match) {
        st = GCP_STATUS_ASSERTION;
      } else {
#pragma unroll
        for (int l = 0; l < 8; l++) out[l] = both ? old_c[l] : new_c[l];
      }
    }
  }
  store_fr(a.new_roots + idx * 8, out);
  a.status[idx] = st;
}

// ---------------------------------------------------------------------------------------------------------
// Processor on the verifier's pipeline: scan (HBM-bound, lidx / last-zero / canonical per proof) -> counting sort by
// lidx -> prep (validation, leaf hashes, the levels below the insertion level, which need no sibling) -> path (both
// chains folded over the sibling levels by warps of equal path length, siblings staged with cp.async).  The
// thread-per-proof kernel above re-read every sibling twice with 32-byte strided loads and diverged over the path
// length; it stays as the reference form for tiny batches (launch_smt_process picks).
// ---------------------------------------------------------------------------------------------------------
template <int HASHER>
__global__ void __launch_bounds__(128) smt_process_prep_kernel(SmtProcessArgs a, const u16* __restrict__ lidx_arr,
                                                               const u8* __restrict__ info_arr, u32* __restrict__ acc_old_out,
                                                               u32* __restrict__ acc_new_out) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= a.n) return;
  const int n = a.n_levels;
  const u32 fnc0 = a.fnc0[idx], fnc1 = a.fnc1[idx], is0 = a.is_old0[idx];
  u8 st = GCP_STATUS_OK;
  if ((fnc0 | fnc1 | is0) > 1u) st = GCP_STATUS_NOT_BOOLEAN;
  const bool enabled = (fnc0 | fnc1) != 0;
  bool canon = true;
  u32 okey[8], oval[8], nkey[8], nval[8], oroot[8];
  load_elem(okey, canon, a.old_keys + idx * 8, a.mont);
  load_elem(oval, canon, a.old_values + idx * 8, a.mont);
  load_elem(nkey, canon, a.new_keys + idx * 8, a.mont);
  load_elem(nval, canon, a.new_values + idx * 8, a.mont);
  load_fr(oroot, a.old_roots + idx * 8);
  canon = canon && fr_is_canonical(oroot);
  u32 ok_int[8], nk_int[8];
  key_integer(ok_int, a.old_keys + idx * 8, a.mont);
  key_integer(nk_int, a.new_keys + idx * 8, a.mont);
  const u32 info = info_arr[idx];
  const int lidx = (int)lidx_arr[idx];
  canon = canon && (info & SMT_INFO_CANONICAL) != 0;
  const bool last_zero = (info & SMT_INFO_LAST_ZERO) != 0;
  if (st == GCP_STATUS_OK && !canon) st = GCP_STATUS_NONCANONICAL;
  if (st == GCP_STATUS_OK && !(key_in_range(ok_int, n) && key_in_range(nk_int, n))) st = GCP_STATUS_KEY_RANGE;
  if (st == GCP_STATUS_OK && enabled && !last_zero) st = GCP_STATUS_ASSERTION;  // LevIns (lev_ins.go:16-20)
  const bool keys_equal = eq256(ok_int, nk_int);
  if (st == GCP_STATUS_OK && fnc0 == 0u && fnc1 == 1u && !keys_equal) st = GCP_STATUS_ASSERTION;  // processor.go:64-70
  int jterm = lidx;
  if (st == GCP_STATUS_OK && enabled && fnc0 == 1u && is0 == 0u) {
    jterm = -1;
    for (int i = lidx; i < n; i++) {
      if (((ok_int[i >> 5] ^ nk_int[i >> 5]) >> (i & 31)) & 1u) {
        jterm = i;
        break;
      }
    }
    if (jterm < 0) st = GCP_STATUS_ASSERTION;  // no terminal state: processor.go:47
  }
  u32 acc_old[8], acc_new[8];
  fr_set_zero(acc_old);
  fr_set_zero(acc_new);
  if (st == GCP_STATUS_OK && enabled) {
    u32 h1old[8], h1new[8], one[8];
#pragma unroll
    for (int l = 0; l < 8; l++) one[l] = FR_ONE[l];
    if (a.leaf_hash_form) {
      fr_copy(h1old, oval);
      fr_copy(h1new, nval);
    }
#pragma unroll 1
    for (int h = 0; h < 2 && !a.leaf_hash_form; h++) {
      u32 kk[8], vv[8], res[8];
#pragma unroll
      for (int l = 0; l < 8; l++) {
        kk[l] = h ? nkey[l] : okey[l];
        vv[l] = h ? nval[l] : oval[l];
      }
      smt_hash1<HASHER>(res, kk, vv, one, a.hkeys);
#pragma unroll
      for (int l = 0; l < 8; l++) {
        if (h)
          h1new[l] = res[l];
        else
          h1old[l] = res[l];
      }
    }
    const bool is_upd = (fnc0 == 0u), is_old0 = (fnc0 == 1u && is0 == 1u);
    fr_copy(acc_new, h1new);
    if (!is_old0) fr_copy(acc_old, h1old);  // old0: the old chain starts from the empty slot (0)
    if (!is_upd && !is_old0) {
      // new1 at jterm: new = H(sw(b; hash1New, hash1Old)); bot on jterm-1 .. lidx: new = H(sw(b; new, 0)); the old chain
      // keeps hash1Old through these levels
      for (int i = jterm; i >= lidx; i--) {
        const u32 bit = (nk_int[i >> 5] >> (i & 31)) & 1u;
        u32 other[8], lft[8], rgt[8];
        if (i == jterm)
          fr_copy(other, h1old);
        else
          fr_set_zero(other);
#pragma unroll
        for (int l = 0; l < 8; l++) {
          lft[l] = bit ? other[l] : acc_new[l];
          rgt[l] = bit ? acc_new[l] : other[l];
        }
        smt_hash2<HASHER>(acc_new, lft, rgt, a.hkeys);
      }
    }
  }
  store_fr(acc_old_out + idx * 8, acc_old);
  store_fr(acc_new_out + idx * 8, acc_new);
  a.status[idx] = st;
}

template <int MIN_BLOCKS, int HASHER>
__global__ void __launch_bounds__(SMT_WARPS * 32, MIN_BLOCKS) smt_process_path_kernel(SmtProcessArgs a, const u32* __restrict__ perm,
                                                                       const u16* __restrict__ lidx_arr,
                                                                       const u32* __restrict__ acc_old_in,
                                                                       const u32* __restrict__ acc_new_in) {
  __shared__ __align__(16) u32 tiles[2][SMT_WARPS][32 * SMT_ROW_WORDS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const size_t warp_base = (size_t)blockIdx.x * blockDim.x + warp * 32;
  if (warp_base >= a.n) return;
  const bool in_range = warp_base + lane < a.n;
  const size_t idx = in_range ? perm[warp_base + lane] : 0;
  const int n = a.n_levels;
  u8 st = in_range ? a.status[idx] : (u8)GCP_STATUS_NONCANONICAL;
  const u32 fnc0 = a.fnc0[idx], fnc1 = a.fnc1[idx];
  const bool enabled = (fnc0 | fnc1) != 0;
  const int lidx = in_range ? (int)lidx_arr[idx] : 0;
  const bool live = in_range && st == GCP_STATUS_OK && enabled;
  u32 acc[2][8];  // [0] old chain, [1] new chain
  load_fr(acc[0], acc_old_in + idx * 8);
  load_fr(acc[1], acc_new_in + idx * 8);
  u32 key[8];
  key_integer(key, a.new_keys + idx * 8, a.mont);
  int warp_top = live ? lidx : 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) warp_top = max(warp_top, __shfl_xor_sync(0xffffffffu, warp_top, o));
  const int first_chunk = (warp_top + SMT_CH - 1) / SMT_CH - 1;
  if (first_chunk >= 0) smt_stage_chunk(tiles[first_chunk & 1][warp], a.siblings, (u32)idx, live, n, first_chunk, lane);
#pragma unroll 1
  for (int c = first_chunk; c >= 0; c--) {
    if (c > 0) {
      smt_stage_chunk(tiles[(c - 1) & 1][warp], a.siblings, (u32)idx, live, n, c - 1, lane);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncwarp();
    const u32* row = tiles[c & 1][warp] + lane * SMT_ROW_WORDS;
    const int hi = min(n, c * SMT_CH + SMT_CH) - 1;
#pragma unroll 1
    for (int i = hi; i >= c * SMT_CH; i--) {
      if (!live || i >= lidx) continue;
      u32 s[8];
      {
        const uint4* q = reinterpret_cast<const uint4*>(row + (i - c * SMT_CH) * 8);
        uint4 v0 = q[0], v1 = q[1];
        u32 x[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
        if (a.mont)
          fr_copy(s, x);
        else
          fr_to_mont(s, x);
      }
      const u32 bit = (key[i >> 5] >> (i & 31)) & 1u;
      // top level: old_i = H(sw(b; old_{i+1}, sib)), new_i = H(sw(b; new_{i+1}, sib)); one inlined copy of Hash2
#pragma unroll 1
      for (int h = 0; h < 2; h++) {
        u32 lft[8], rgt[8];
#pragma unroll
        for (int l = 0; l < 8; l++) {
          const u32 child = h ? acc[1][l] : acc[0][l];
          lft[l] = bit ? s[l] : child;
          rgt[l] = bit ? child : s[l];
        }
        u32 res[8];
        smt_hash2<HASHER>(res, lft, rgt, a.hkeys);
#pragma unroll
        for (int l = 0; l < 8; l++) {
          if (h)
            acc[1][l] = res[l];
          else
            acc[0][l] = res[l];
        }
      }
    }
    __syncwarp();
  }
  if (!in_range) return;
  u32 out[8];
  fr_set_zero(out);
  if (st == GCP_STATUS_OK) {
    if (!enabled) {
      load_fr(out, a.old_roots + idx * 8);  // nop: newRoot = oldRoot
    } else {
      u32 old_c[8], new_c[8], oroot[8];
      load_fr(oroot, a.old_roots + idx * 8);
      if (a.mont) {
        fr_copy(old_c, acc[0]);
        fr_copy(new_c, acc[1]);
        fr_canon(old_c);
        fr_canon(new_c);
      } else {
        fr_from_mont(old_c, acc[0]);
        fr_from_mont(new_c, acc[1]);
      }
      const bool both = (fnc0 & fnc1) != 0;
      const bool match = both ? eq256(new_c, oroot) : eq256(old_c, oroot);  // ForceEqualIfEnabled, processor.go:60
      if (!match) {
        st = GCP_STATUS_ASSERTION;
      } else {
#pragma unroll
        for (int l = 0; l < 8; l++) out[l] = both ? old_c[l] : new_c[l];
      }
    }
  }
  store_fr(a.new_roots + idx * 8, out);
  a.status[idx] = st;
}

// Hash1 input rows (tree/smt/hash.go:10-19): (key, values..., 1), one thread per element
__global__ void smt_leaf_rows_kernel(const u32* __restrict__ keys, const u32* __restrict__ values, int n_values, size_t n,
                                     u32* __restrict__ rows, int mont) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int arity = n_values + 2;
  if (idx >= n * (size_t)arity) return;
  const size_t item = idx / arity;
  const int j = (int)(idx - item * arity);
  u32 v[8];
  if (j == 0) {
    load_fr(v, keys + item * 8);
  } else if (j <= n_values) {
    load_fr(v, values + (item * (size_t)n_values + (j - 1)) * 8);
  } else {
#pragma unroll
    for (int l = 0; l < 8; l++) v[l] = mont ? FR_ONE[l] : (l == 0 ? 1u : 0u);
  }
  store_fr(rows + idx * 8, v);
}

// ---------------------------------------------------------------------------------------------------
// arbo packed siblings -> Assignment.Siblings rows
// ---------------------------------------------------------------------------------------------------
// What the callers of this path do on the CPU before they can fill smt.Assignment: arbo.UnpackSiblings on the
// byte string GenProof returns, then pad with zeros (or cut) to `levels`
// (/root/reference/tree/smt/wrapper_arbo.go:63-76,166-179, testutil/utils.go:152-166).  Wire format (arbo
// PackSiblings, un-vendored dependency, restated in oracle/smt.py::pack_siblings):
//   [u16 LE full length][u16 LE bitmap length L][L bytes bitmap, bit i = byte i/8 bit i%8][32 B LE per set bit]
// One warp per proof; lane l expands levels l, l+32, ...  A proof arbo would reject (length field mismatch, bitmap
// running past the string, a set bit whose 32 bytes are cut short) gets bad = GCP_STATUS_MALFORMED and an all-zero
// row.  Siblings are canonical little-endian integers on the wire; for the Montgomery element format they are
// converted here (a sibling >= r is stored as it is, so the verifier's scan reports it as non-canonical).
struct SmtUnpackArgs {
  const u8* packed;    // the chunk's packed bytes
  const u64* offsets;  // n + 1 absolute offsets into the caller's blob
  u64 base;            // offset of packed[0] in the caller's blob
  u64 packed_bytes;    // bytes available behind `packed`
  size_t n;
  int n_levels;
  u32* siblings;       // n x n_levels x 8
  u8* bad;             // n
  int mont;
  // arbo's post-insert flow (wrapper_arbo.go:170-172): when both are given, the LAST unpacked sibling of proof i is dropped
  // where is_old0[i] == 0 && fnc1[i] == 0 (GenProof ran after the add: that sibling is the displaced old leaf)
  const u8* drop_is_old0;
  const u8* drop_fnc1;
};

__device__ __forceinline__ void load_unaligned32(u32 (&r)[8], const u8* p) {
  const uintptr_t a = (uintptr_t)p;
  const u32* w = (const u32*)(a & ~(uintptr_t)3);
  const u32 sh = (u32)(a & 3) * 8;
  u32 prev = __ldg(w);
#pragma unroll
  for (int k = 0; k < 8; k++) {
    u32 next = (k < 7 || sh) ? __ldg(w + k + 1) : 0;
    r[k] = sh ? __funnelshift_r(prev, next, sh) : prev;
    prev = next;
  }
}

__global__ void __launch_bounds__(128) smt_unpack_kernel(SmtUnpackArgs a) {
  const size_t proof = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (proof >= a.n) return;
  const u64 beg = a.offsets[proof], end = a.offsets[proof + 1];
  bool bad = end < beg || beg < a.base || end - a.base > a.packed_bytes;
  const u64 len = bad ? 0 : end - beg;
  const u8* b = a.packed + (bad ? 0 : beg - a.base);
  u32 L = 0;
  if (!bad) bad = len < 4 || len > 0xffff;
  if (!bad) {
    const u32 full = (u32)b[0] | ((u32)b[1] << 8);
    L = (u32)b[2] | ((u32)b[3] << 8);
    bad = full != (u32)len || 4 + (u64)L > len;
  }
  const u8* bitmap = b + 4;
  const u8* data = bitmap + L;
  const u32 dlen = bad ? 0 : (u32)len - 4 - L;
  const u32 avail = dlen / 32;
  if (!bad && (dlen & 31)) {
    // arbo walks every bit of the bitmap: a set bit that starts inside the data but is cut short is an error
    u32 cnt = 0;
    for (u32 j = lane; j < L; j += 32) cnt += __popc((u32)bitmap[j]);
#pragma unroll
    for (int o = 16; o; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    bad = cnt > avail;
  }
  // index of the last unpacked sibling (the highest set bit of the bitmap) where the post-insert rule drops it
  int drop = -1;
  if (!bad && a.drop_is_old0 && a.drop_is_old0[proof] == 0 && a.drop_fnc1[proof] == 0) {
    int top = -1;
    for (u32 j = lane; j < L; j += 32)
      if (bitmap[j]) top = (int)j;
#pragma unroll
    for (int o = 16; o; o >>= 1) top = max(top, __shfl_xor_sync(0xffffffffu, top, o));
    if (top < 0)
      bad = true;  // nothing to drop: the reference's slice expression panics on an empty sibling list
    else
      drop = top * 8 + (31 - __clz((u32)bitmap[top]));
  }
  for (int i = lane; i < a.n_levels; i += 32) {
    u32 v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (!bad && i != drop && (u32)i < 8 * L && ((bitmap[i >> 3] >> (i & 7)) & 1)) {
      u32 rank = __popc((u32)bitmap[i >> 3] & ((1u << (i & 7)) - 1));
      for (int j = 0; j < (i >> 3); j++) rank += __popc((u32)bitmap[j]);
      if (rank < avail) {
        load_unaligned32(v, data + (size_t)rank * 32);
        if (a.mont && fr_is_canonical(v)) {
          fr_to_mont(v, v);
          fr_canon(v);
        }
      }
    }
    store_fr(a.siblings + (proof * (size_t)a.n_levels + i) * 8, v);
  }
  if (lane == 0) a.bad[proof] = bad ? GCP_STATUS_MALFORMED : 0;
}

// a malformed proof never reaches the gadget: the caller's unpack step returns an error (wrapper_arbo.go:64-67)
__global__ void smt_apply_bad_kernel(const u8* bad, size_t n, u8* flags, u8* status, u32* out_roots) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || !bad[i]) return;
  flags[i] = 0;
  status[i] = bad[i];
  if (out_roots) {
#pragma unroll
    for (int l = 0; l < 8; l++) out_roots[i * 8 + l] = 0;
  }
}

}  // namespace gcp
