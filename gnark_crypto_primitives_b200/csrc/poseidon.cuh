// Poseidon permutation over BN254 Fr (circomlib "optimized" schedule), device side.
//
// Computes exactly the value of (*Poseidon).Sum, /root/reference/hash/native/bn254/poseidon/poseidon.go:116-183:
//   state = [0, in...] ; ark(C,0) ; 3 x {sigma all, ark, mix(M)} ; sigma all, ark, mix(P) ;
//   RP x {sigma(s0), +C, s0' = <S[0..t), state>, s_k += s0 * S[t+k-1]} ; 3 x {sigma all, ark, mix(M)} ;
//   sigma all ; out = <M[.][0], state>
// with mix(i) = sum_j m[j][i] * in[j]  (poseidon.go:213-224) and the table sizes of constants.go.
// All arithmetic is exact mod r, so any evaluation order yields the same canonical output; here every
// matrix row is ONE lazy dot product (t*64 wide multiplies + one 72-multiply Montgomery reduction).
//
// Table layout (device, Montgomery form, canonical): C[8t+RP] | S[(2t-1)RP] | M[t*t] | P[t*t], each element
// 8 x u32.  M and P are stored as in the reference: element (j,i) at j*t+i.
#pragma once
#include "fr.cuh"
#include "kernels.h"

namespace gcp {


constexpr int POSEIDON_RP[16] = {56, 57, 56, 60, 60, 63, 64, 63, 60, 66, 60, 65, 70, 60, 64, 68};  // poseidon.go:119

// Constant-memory copies for the two hot widths: t=3 (Hash2, SMT node) and t=4 (Hash1, SMT leaf).
constexpr int POS3_ELEMS = (8 * 3 + 57) + 5 * 57 + 9 + 9;  // 384
constexpr int POS4_ELEMS = (8 * 4 + 56) + 7 * 56 + 16 + 16;  // 512
__device__ __constant__ u32 c_pos3[POS3_ELEMS * 8];
// t = 3 pair schedule (below): one derived constant per pair of partial rounds, d_q = S_q[1] c_{q-1,1} + S_q[2] c_{q-1,2}
// for q = 2, 4, ..., 56 (c_{q,k} = S_q[t+k-1] is the rank-1 coefficient), computed on the device from the uploaded table.
#ifndef GCP_POS3_PAIRS
#define GCP_POS3_PAIRS 1
#endif
constexpr int POS3_PAIR_COUNT = (57 - 1) / 2;  // 28
__device__ __constant__ u32 c_pos3_pair[POS3_PAIR_COUNT * 8];
__device__ u32 g_pos3_pair[POS3_PAIR_COUNT * 8];
__device__ __constant__ u32 c_pos4[POS4_ELEMS * 8];

template <int T>
struct ConstTab;
template <>
struct ConstTab<3> {
  static constexpr int RP = 57;
  __device__ static __forceinline__ const u32* base() { return c_pos3; }
};
template <>
struct ConstTab<4> {
  static constexpr int RP = 56;
  __device__ static __forceinline__ const u32* base() { return c_pos4; }
};

__device__ __forceinline__ void load_const(u32 (&r)[8], const u32* p) {
#pragma unroll
  for (int i = 0; i < 8; i++) r[i] = p[i];
}

// x <- x^5   (poseidon.go:199-203)
__device__ __forceinline__ void sigma(u32 (&x)[8]) {
  u32 x2[8], x4[8];
  fr_sqr(x2, x);
  fr_sqr(x4, x2);
  fr_mul(x, x4, x);
}

// Register-resident permutation for small T with tables in constant memory.
// in/out: lazy Montgomery. s[0] must be the capacity element (0).
//
// Code-size discipline: the whole permutation is ONE loop over the 8 + RP rounds and contains a handful of
// multiplier bodies - (a) the S-box squaring (run twice) and multiply per element, (b) one lazy dot-product row (T
// products + one reduction) for the full rounds, (c) the fused partial-round body (dot row + T-1 rank-1 updates
// advancing row by row together).  Elements are brought to position 0 by rotating the register file instead of unrolling over
// the state index.  The first version unrolled every round body (~260 KB of SASS) and was bound by instruction
// fetch (ncu: stall_no_instruction, icc hit rate 89 %); this one is ~40 KB and stays in the instruction cache
// (hit rate 99.95 %).
template <int T>
__device__ __forceinline__ void rotate_left(u32 (&s)[T][8]) {
#pragma unroll
  for (int l = 0; l < 8; l++) {
    u32 t0 = s[0][l];
#pragma unroll
    for (int j = 0; j + 1 < T; j++) s[j][l] = s[j + 1][l];
    s[T - 1][l] = t0;
  }
}

template <int T, bool PAIR_SCHEDULE = false>
__device__ __forceinline__ void poseidon_permute_const(u32 (&s)[T][8], u32 (&out)[8]) {
  constexpr int RP = ConstTab<T>::RP;
  const u32* C = ConstTab<T>::base();
  const u32* S = C + (8 * T + RP) * 8;
  const u32* M = S + (2 * T - 1) * RP * 8;
  const u32* P = M + T * T * 8;
  const u32 P2[8] = GCP_2P_LIMBS;
  u32 cst[8];

  // ark(C, 0)
#pragma unroll
  for (int j = 0; j < T; j++) {
    load_const(cst, C + j * 8);
    fr_add(s[j], s[j], cst);
  }

  // t = 3, partial rounds in PAIRS (bit-identical values, 64 fewer wide multiplies per round on average).  A partial
  // round is s0' = <S[0..t), (x, s1, s2)>, s_k += x c_k with x the S-boxed s0: t + (t-1) products and t reductions.  The
  // s_k of round A are only ever read by round B's dot row and rank-1 update, and both are linear in them:
  //   round A:  x0 = sbox(s0);  s0 = S_A0 x0 + S_A1 s1 + S_A2 s2                       3 products, 1 reduction
  //   round B:  x1 = sbox(s0);  s0 = S_B0 x1 + S_B1 s1 + S_B2 s2 + d x0                4 products, 1 reduction
  //             s_k += c_Ak x0 + c_Bk x1   (k = 1, 2)                                  4 products, 2 reductions
  // with d = S_B1 c_A1 + S_B2 c_A2 precomputed: 11 products + 4 reductions instead of 10 + 6 per two rounds (one unit =
  // 64 wide multiplies).  RP = 57 is odd: the first partial round runs as a B round with x0 = 0.  Round A's dot row is
  // the full rounds' row body with stride 1.  Bounds (r / 2^256 = 0.18904, products of values a r and b r reduce to
  // < (0.18904 sum(a b) + 1) r): x0 is kept canonical (< r, off the critical path), y = x^5 + c < 2.89 r unreduced,
  // after A: s0 < (2.89 + 2 + 2) 0.189 + 1 = 2.31 r, after B: s0 < (2.77 + 2 + 2 + 1) 0.189 + 1 = 2.47 r - every squaring
  // input stays below 2^255 = 2.645 r; the rank-1 sum is < (1 + 2.89) 0.189 + 1 = 1.74 r, added to s_k < 2 r and brought
  // back under 2 r by fr_add.
  // PAIR_SCHEDULE: the batch kernel (Hash2 135.5 -> 137.5 M/s) and, through the out-of-line poseidon_hash2_pairs_ool of
  // smt.cuh, the tree kernels.  INLINED into smt_path_kernel the eight registers of x0 and the longer round-B body do not
  // fit four blocks per SM (158 registers, or 128 with 4.6 % more executed instructions: dense proofs 786 -> 764 / 773 k
  // per second at 2^17, profiles/r02_poseidon_pair_schedule.jsonl).
  constexpr bool PAIRS = (T == 3) && PAIR_SCHEDULE && (GCP_POS3_PAIRS != 0);
  u32 xh[8];  // canonical x0 of the pending round A (0: none)
#pragma unroll
  for (int l = 0; l < 8; l++) xh[l] = 0;
  u32 n[T][8];
#pragma unroll 1
  for (int r = 0; r < 8 + RP; r++) {
    const bool full = (r < 4) || (r >= 4 + RP);
    const bool last = (r == 7 + RP);
    const bool phase_a = PAIRS && !full && (((r - 4) & 1) != 0);
    // round constants added after the S-box: full rounds c[(r+1)T + j] (first half), c[(r+1)T + RP - ... ] (second
    // half, poseidon.go:172), partial rounds c[5T + (r-4)] (poseidon.go:154); none in the last round.
    const u32* crow = (r < 4) ? C + (r + 1) * T * 8 : (full ? C + ((r - RP + 1) * T + RP) * 8 : C + (5 * T + (r - 4)) * 8);
    const int nsig = full ? T : 1;
#pragma unroll 1
    for (int k = 0; k < nsig; k++) {
      // (a) S-box on s[0]: y = x*x, y = y*y, y = y*x   (poseidon.go:199-203)
      u32 y[8];
#pragma unroll
      for (int l = 0; l < 8; l++) y[l] = s[0][l];
#pragma unroll 1
      for (int it = 0; it < 2; it++) fr_sqr(y, y);
      fr_mul(y, y, s[0]);
      if (!last) {
        load_const(cst, crow + k * 8);
        if (full || T > 3)
          fr_add(y, y, cst);
        else
          add256(y, y, cst);  // partial rounds, t <= 3: left unreduced (< 2.75 r), see the bound note below
      }
#pragma unroll
      for (int l = 0; l < 8; l++) s[0][l] = y[l];
      if (full) rotate_left<T>(s);
    }
    if (full || phase_a) {
      // (b) matrix rows as lazy dot products: n_i = sum_j m[j][i] s_j (M, or P after round 3); the last round only
      // needs column 0.  Row-major over the terms: after product rows I of every term limb I is complete, so reduction
      // row I follows at once and its serial chain overlaps with the product rows still to come.  Round A of a pair
      // (t = 3) is ONE such row over S[0..t) (coefficient stride 1), result to s0 without the conditional subtraction.
      const u32* coef = full ? ((r == 3) ? P : M) : S + (2 * T - 1) * (r - 4) * 8;
      const int cstride = full ? T * 8 : 8;
      const int nrows = (last || phase_a) ? 1 : T;
      if (phase_a) {
        const u32 P1[8] = GCP_P_LIMBS;
#pragma unroll
        for (int l = 0; l < 8; l++) xh[l] = s[0][l];
        cond_sub(xh, P2);
        cond_sub(xh, P1);
      }
#pragma unroll 1
      for (int i = 0; i < nrows; i++) {
        Wide w;
        wide_zero(w);
        u32 c = 0;
        const u32* cf = coef + i * 8;
#define GCP_DOT_ROW(I)                                                                  \
  _Pragma("unroll") for (int j = 0; j < T; j++) mac_row<I>(w, s[j], cf[j * cstride + I]); \
  redc_row<I>(w, c);
        GCP_DOT_ROW(0) GCP_DOT_ROW(1) GCP_DOT_ROW(2) GCP_DOT_ROW(3) GCP_DOT_ROW(4) GCP_DOT_ROW(5) GCP_DOT_ROW(6) GCP_DOT_ROW(7)
#undef GCP_DOT_ROW
        rotate_left<T>(n);
        wide_redc_finish(w, c, n[T - 1]);
        if (!phase_a) cond_sub(n[T - 1], P2);
      }
      if (phase_a) {
#pragma unroll
        for (int l = 0; l < 8; l++) s[0][l] = n[T - 1][l];
      } else if (!last) {
#pragma unroll
        for (int j = 0; j < T; j++)
#pragma unroll
          for (int l = 0; l < 8; l++) s[j][l] = n[j][l];
      }
    } else if constexpr (PAIRS) {
      // round B of a pair (or the single first partial round, xh = 0): see the schedule above
      const int q = r - 4;
      const u32* srow = S + (2 * T - 1) * q * 8;
      const u32* prow = srow - (2 * T - 1) * 8;  // round A's row (q = 0: the tail of C, multiplied by xh = 0)
      const u32* dq = c_pos3_pair + (q >= 2 ? (q / 2 - 1) : 0) * 8;
      Wide wd, wk[T - 1];
      wide_zero(wd);
#pragma unroll
      for (int k = 0; k < T - 1; k++) wide_zero(wk[k]);
      u32 cd = 0, ck[T - 1];
#pragma unroll
      for (int k = 0; k < T - 1; k++) ck[k] = 0;
#define GCP_PAIR_ROW(I)                                                                       \
  _Pragma("unroll") for (int j = 0; j < T; j++) mac_row<I>(wd, s[j], srow[j * 8 + I]);        \
  mac_row<I>(wd, xh, dq[I]);                                                                  \
  redc_row<I>(wd, cd);                                                                        \
  _Pragma("unroll") for (int k = 0; k < T - 1; k++) {                                         \
    mac_row<I>(wk[k], s[0], srow[(T + k) * 8 + I]);                                           \
    mac_row<I>(wk[k], xh, prow[(T + k) * 8 + I]);                                             \
    redc_row<I>(wk[k], ck[k]);                                                                \
  }
      GCP_PAIR_ROW(0) GCP_PAIR_ROW(1) GCP_PAIR_ROW(2) GCP_PAIR_ROW(3)
      GCP_PAIR_ROW(4) GCP_PAIR_ROW(5) GCP_PAIR_ROW(6) GCP_PAIR_ROW(7)
#undef GCP_PAIR_ROW
#pragma unroll
      for (int k = 1; k < T; k++) {
        u32 prod[8];
        wide_redc_finish(wk[k - 1], ck[k - 1], prod);
        fr_add(s[k], s[k], prod);
      }
      wide_redc_finish(wd, cd, s[0]);
    } else {
      // partial round (poseidon.go:152-166): n0 = sum_j S[j] s_j and s_k += s_0 * S[T+k-1] all read the post-S-box state
      // and are independent, so their T accumulators advance row by row together: T independent carry / reduction
      // chains in flight per thread.
      const u32* srow = S + (2 * T - 1) * (r - 4) * 8;
      Wide wd, wk[T - 1];
      wide_zero(wd);
#pragma unroll
      for (int k = 0; k < T - 1; k++) wide_zero(wk[k]);
      u32 cd = 0, ck[T - 1];
#pragma unroll
      for (int k = 0; k < T - 1; k++) ck[k] = 0;
#define GCP_PARTIAL_ROW(I)                                                                    \
  _Pragma("unroll") for (int j = 0; j < T; j++) mac_row<I>(wd, s[j], srow[j * 8 + I]);        \
  redc_row<I>(wd, cd);                                                                        \
  _Pragma("unroll") for (int k = 0; k < T - 1; k++) {                                         \
    mac_row<I>(wk[k], s[0], srow[(T + k) * 8 + I]);                                           \
    redc_row<I>(wk[k], ck[k]);                                                                \
  }
      GCP_PARTIAL_ROW(0) GCP_PARTIAL_ROW(1) GCP_PARTIAL_ROW(2) GCP_PARTIAL_ROW(3)
      GCP_PARTIAL_ROW(4) GCP_PARTIAL_ROW(5) GCP_PARTIAL_ROW(6) GCP_PARTIAL_ROW(7)
#undef GCP_PARTIAL_ROW
#pragma unroll
      for (int k = 1; k < T; k++) {
        u32 prod[8];
        wide_redc_finish(wk[k - 1], ck[k - 1], prod);
        fr_add(s[k], s[k], prod);
      }
      wide_redc_finish(wd, cd, s[0]);
      // Bounds for t <= 3 (r/2^256 = 0.18904): with s_0 < 2.28 r the S-box gives x^2 < 1.98 r, x^4 < 1.74 r,
      // x^5 < 1.75 r, so y = x^5 + c < 2.75 r; the dot row is then < (2.75 + 2 (t-1)) r^2 / 2^256 + r < 2.28 r again.
      // Every squaring input stays below 2^255 = 2.645 r and every sum below 2^256, so neither y nor the new s_0
      // needs a conditional subtraction on the round-to-round critical path.  t = 4 would reach 2.65 r: it reduces.
      if constexpr (T > 3) cond_sub(s[0], P2);
    }
  }
#pragma unroll
  for (int l = 0; l < 8; l++) out[l] = n[T - 1][l];
}

// Hash2 (tree/smt/hash.go:21-27): Poseidon(l, r), lazy Montgomery in and out.
__device__ __forceinline__ void poseidon_hash2(u32 (&out)[8], const u32 (&l)[8], const u32 (&r)[8]) {
  u32 s[3][8];
#pragma unroll
  for (int i = 0; i < 8; i++) {
    s[0][i] = 0;
    s[1][i] = l[i];
    s[2][i] = r[i];
  }
#ifndef GCP_SMT_PAIRS
#define GCP_SMT_PAIRS 0  // the tree kernels keep one partial round per iteration (measured: see poseidon_permute_const)
#endif
  poseidon_permute_const<3, (GCP_SMT_PAIRS != 0)>(s, out);
}

// Poseidon(a, b, c) — Hash1 (tree/smt/hash.go:10-19) is poseidon_hash3(key, value, 1).
__device__ __forceinline__ void poseidon_hash3(u32 (&out)[8], const u32 (&a)[8], const u32 (&b)[8],
                                               const u32 (&c)[8]) {
  u32 s[4][8];
#pragma unroll
  for (int i = 0; i < 8; i++) {
    s[0][i] = 0;
    s[1][i] = a[i];
    s[2][i] = b[i];
    s[3][i] = c[i];
  }
  poseidon_permute_const<4>(s, out);
}

}  // namespace gcp

// ------------------------------------------------------------------------------------------
// Generic width (t = 2..17): tables in global memory, state in per-thread local arrays.
// Used for every arity other than 2 and 3 and by MultiHash (16-input chunks are t = 17).
// ------------------------------------------------------------------------------------------
namespace gcp {

constexpr int POSEIDON_MAX_T = 17;
// Lazy dot products in the generic kernel: how many (state x constant) products share one Montgomery reduction.
// With r / 2^256 = 0.1891, constants < r and n terms:  state < 2r:  sum < 2n r^2 must stay below 2^512 - r 2^256
// (0.0715 n < 0.811, n <= 11) and the reduced value 0.378 n r + r below 2^256 = 5.29 r (n <= 11);  state < r (canonical):
// 0.0357 n < 0.811 and 0.189 n + 1 < 5.29, n <= 22.  So t <= 11 takes the whole row in one reduction as it is, and
// t = 12..17 keeps the state canonical (one conditional subtraction of r per state write) and does the same.  Two
// conditional subtractions of 2r bring the row result (< 5.2 r) back under 2r.  (The first version reduced every 5
// terms: 4 reductions per row at t = 17.)
// Round B of a partial-round pair (below) adds one term (a canonical value times a constant) to its dot row: with the
// state < 2r that is (2t + 1) r^2, so the non-canonical form now stops at t = 10.
constexpr int LAZY_DOT_NONCANON_MAX = 10;

__device__ __forceinline__ bool generic_state_canonical(int t) { return t > LAZY_DOT_NONCANON_MAX; }

__device__ __forceinline__ void load_global_const(u32 (&r)[8], const u32* p) {
  const uint4* q = reinterpret_cast<const uint4*>(p);
  uint4 x = __ldg(q), y = __ldg(q + 1);
  r[0] = x.x; r[1] = x.y; r[2] = x.z; r[3] = x.w;
  r[4] = y.x; r[5] = y.y; r[6] = y.z; r[7] = y.w;
}

// out = sum_j coef[j * stride] * s[j]  over j in [0, t)   (lazy Montgomery, < 2r; < r when the state is kept canonical)
// ex_c != nullptr: one more term ex_x * ex_c (ex_x canonical: round B of a partial-round pair)
__device__ __noinline__ void generic_dot(u32 (&out)[8], const u32 (*s)[8], const u32* coef, int stride, int t,
                                         const u32* ex_x = nullptr, const u32* ex_c = nullptr) {
  const u32 P2[8] = GCP_2P_LIMBS;
  const u32 P1[8] = GCP_P_LIMBS;
  Wide w;
  wide_zero(w);
#pragma unroll 1
  for (int j = 0; j < t; j++) {
    u32 c[8], x[8];
    load_global_const(c, coef + (size_t)j * stride * 8);
#pragma unroll
    for (int l = 0; l < 8; l++) x[l] = s[j][l];
    wide_mac(w, x, c);
  }
  if (ex_c) {
    u32 c[8], x[8];
    load_global_const(c, ex_c);
#pragma unroll
    for (int l = 0; l < 8; l++) x[l] = ex_x[l];
    wide_mac(w, x, c);
  }
  u32 total[8];
  wide_redc(w, total);
  cond_sub(total, P2);
  cond_sub(total, P2);
  if (generic_state_canonical(t)) cond_sub(total, P1);
#pragma unroll
  for (int l = 0; l < 8; l++) out[l] = total[l];
}

__device__ __noinline__ void generic_sigma_ark(u32 (&x)[8], const u32* c, bool canonical) {
  const u32 P1[8] = GCP_P_LIMBS;
  u32 k[8];
  sigma(x);
  load_global_const(k, c);
  fr_add(x, x, k);
  if (canonical) cond_sub(x, P1);
}

// s[0..t) in, lazy Montgomery, s[0] = 0.  Result in out.  poseidon.go:116-183.
__device__ __forceinline__ void poseidon_permute_generic(u32 (*s)[8], u32 (*n)[8], u32 (&out)[8],
                                                         const PoseidonTable& tab) {
  const int t = tab.t, rp = tab.RP;
  const bool canonical = generic_state_canonical(t);
  const u32 P1[8] = GCP_P_LIMBS;
  u32 x[8], k[8];
#pragma unroll 1
  for (int j = 0; j < t; j++) {
    load_global_const(k, tab.C + j * 8);
#pragma unroll
    for (int l = 0; l < 8; l++) x[l] = s[j][l];
    fr_add(x, x, k);
#pragma unroll
    for (int l = 0; l < 8; l++) s[j][l] = x[l];
  }
#pragma unroll 1
  for (int fr = 0; fr < 8; fr++) {
    if (fr == 4) {
      // Partial rounds in PAIRS (the schedule of poseidon_permute_const, for every t; t - 2 units of 64 wide multiplies
      // saved per two rounds: 13 % of the partial rounds at t = 13).  Round A: x0 = sbox(s0), s0 = <S_A[0..t), state>, the
      // rank-1 update is NOT applied.  Round B: x1 = sbox(s0), s0 = <S_B[0..t), state> + d x0 with the derived constant
      // d = sum_k S_B[k] c_A[k] (tab.D, one per pair), then s_k += c_A[k] x0 + c_B[k] x1 as one two-term product with one
      // reduction.  An odd RP starts with one ordinary round.  x0 is kept canonical (the extra dot term's bound above;
      // the two-term product is < 3 r^2, its reduction < 1.57 r, added to s_k < 2r by fr_add).
      int r = 0;
      if (rp & 1) {
#pragma unroll
        for (int l = 0; l < 8; l++) x[l] = s[0][l];
        generic_sigma_ark(x, tab.C + (5 * t) * 8, canonical);
#pragma unroll
        for (int l = 0; l < 8; l++) s[0][l] = x[l];
        const u32* srow = tab.S;
        u32 n0[8];
        generic_dot(n0, s, srow, 1, t);
#pragma unroll 1
        for (int kk = 1; kk < t; kk++) {
          u32 prod[8], y[8];
          load_global_const(k, srow + (t + kk - 1) * 8);
          fr_mul(prod, x, k);
#pragma unroll
          for (int l = 0; l < 8; l++) y[l] = s[kk][l];
          fr_add(y, y, prod);
          if (canonical) cond_sub(y, P1);
#pragma unroll
          for (int l = 0; l < 8; l++) s[kk][l] = y[l];
        }
#pragma unroll
        for (int l = 0; l < 8; l++) s[0][l] = n0[l];
        r = 1;
      }
      const u32* dpair = tab.D;
#pragma unroll 1
      for (; r < rp; r += 2, dpair += 8) {
        u32 x0[8], n0[8];
        const u32* srow_a = tab.S + (size_t)(2 * t - 1) * r * 8;
        const u32* srow_b = srow_a + (2 * t - 1) * 8;
        // round A
#pragma unroll
        for (int l = 0; l < 8; l++) x[l] = s[0][l];
        generic_sigma_ark(x, tab.C + (5 * t + r) * 8, canonical);
#pragma unroll
        for (int l = 0; l < 8; l++) s[0][l] = x[l];
        generic_dot(n0, s, srow_a, 1, t);
#pragma unroll
        for (int l = 0; l < 8; l++) {
          x0[l] = x[l];
          s[0][l] = n0[l];
        }
        if (!canonical) cond_sub(x0, P1);  // x < 2r -> canonical
        // round B
#pragma unroll
        for (int l = 0; l < 8; l++) x[l] = s[0][l];
        generic_sigma_ark(x, tab.C + (5 * t + r + 1) * 8, canonical);
#pragma unroll
        for (int l = 0; l < 8; l++) s[0][l] = x[l];
        generic_dot(n0, s, srow_b, 1, t, x0, dpair);
#pragma unroll 1
        for (int kk = 1; kk < t; kk++) {
          u32 prod[8], y[8], ca[8], cb[8];
          load_global_const(ca, srow_a + (t + kk - 1) * 8);
          load_global_const(cb, srow_b + (t + kk - 1) * 8);
          Wide w;
          wide_zero(w);
          wide_mac(w, x0, ca);
          wide_mac(w, x, cb);
          wide_redc(w, prod);
#pragma unroll
          for (int l = 0; l < 8; l++) y[l] = s[kk][l];
          fr_add(y, y, prod);
          if (canonical) cond_sub(y, P1);
#pragma unroll
          for (int l = 0; l < 8; l++) s[kk][l] = y[l];
        }
#pragma unroll
        for (int l = 0; l < 8; l++) s[0][l] = n0[l];
      }
    }
    // sigma on every element; rounds 0..6 are followed by ark + mix, round 7 by the column-0 output
    const u32* crow = (fr < 4) ? tab.C + (fr + 1) * t * 8 : tab.C + ((fr + 1) * t + rp) * 8;
#pragma unroll 1
    for (int j = 0; j < t; j++) {
#pragma unroll
      for (int l = 0; l < 8; l++) x[l] = s[j][l];
      if (fr < 7) {
        generic_sigma_ark(x, crow + j * 8, canonical);
      } else {
        sigma(x);
        if (canonical) cond_sub(x, P1);
      }
#pragma unroll
      for (int l = 0; l < 8; l++) s[j][l] = x[l];
    }
    if (fr == 7) break;
    const u32* mat = (fr == 3) ? tab.P : tab.M;
#pragma unroll 1
    for (int i = 0; i < t; i++) {
      u32 y[8];
      generic_dot(y, s, mat + i * 8, t, t);  // sum_j mat[j*t+i] * s[j]
#pragma unroll
      for (int l = 0; l < 8; l++) n[i][l] = y[l];
    }
#pragma unroll 1
    for (int i = 0; i < t; i++)
#pragma unroll
      for (int l = 0; l < 8; l++) s[i][l] = n[i][l];
  }
  generic_dot(out, s, tab.M, t, t);  // column 0: sum_j M[j*t+0] * s[j]
}

}  // namespace gcp
