// BN254 scalar field Fr on sm_100a: 8 x 32-bit limbs, Montgomery form (R = 2^256), values kept
// "lazy" in [0, 2r) between operations and made canonical only at the engine boundary.
//
// Design ("Fr arithmetic"):
//  * Every 32x32->64 product is one IMAD.WIDE.U32 whose 64-bit accumulator is an aligned register
//    pair.  A product a[j]*b[i] lands at limb i+j; pairs that start on an even limb live in the
//    E accumulator, pairs that start on an odd limb in the O accumulator, so no pair ever needs
//    re-alignment.  ptxas fuses each `mad.lo.cc / madc.hi.cc` pair below into
//    IMAD.WIDE.U32[.X] with predicate carry-in/out (checked with cuobjdump).
//  * A row (one multiplier word times 4 same-parity multiplicand words) is one 4-deep carry chain;
//    the carry that leaves a chain is counted into a small K limb instead of rippling upwards.
//  * The 512-bit value is   sum e[p] * 2^(64p)  +  sum o[p] * 2^(64p+32)  +  sum k[q] * 2^(32(q+8)).
//    Several products may be accumulated before ONE Montgomery reduction (lazy dot products:
//    Poseidon's matrix rows cost t*64 + 64 wide multiplies (+ 8 plain IMAD) instead of t*128 + 8t).
//  * Product rows and reduction rows are issued in CIOS order (row i of the product, then row i of the
//    reduction) so that the reduction's serial chain overlaps with later product rows.
//
// Value semantics follow gnark's test engine: every api.Add/Mul/Sub is exact arithmetic mod r
// (/root/reference SURVEY appendix A); r literal at hash/emulated/bn254/mimc7/constants.go:18.
#pragma once
#include <cstdint>

namespace gcp {

typedef uint32_t u32;
typedef uint64_t u64;

struct Fr {
  u32 v[8];
};

// r, 2r, -r^-1 mod 2^32, R mod r, R^2 mod r (little-endian 32-bit limbs)
#define GCP_P0 0xf0000001u
#define GCP_P1 0x43e1f593u
#define GCP_P2 0x79b97091u
#define GCP_P3 0x2833e848u
#define GCP_P4 0x8181585du
#define GCP_P5 0xb85045b6u
#define GCP_P6 0xe131a029u
#define GCP_P7 0x30644e72u
#define GCP_NP 0xefffffffu

__device__ __constant__ const u32 FR_P[8] = {GCP_P0, GCP_P1, GCP_P2, GCP_P3, GCP_P4, GCP_P5, GCP_P6, GCP_P7};
__device__ __constant__ const u32 FR_2P[8] = {0xe0000002u, 0x87c3eb27u, 0xf372e122u, 0x5067d090u,
                                              0x0302b0bau, 0x70a08b6du, 0xc2634053u, 0x60c89ce5u};
__device__ __constant__ const u32 FR_ONE[8] = {0x4ffffffbu, 0xac96341cu, 0x9f60cd29u, 0x36fc7695u,
                                               0x7879462eu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u};  // R mod r
__device__ __constant__ const u32 FR_R2[8] = {0xae216da7u, 0x1bb8e645u, 0xe35c59e3u, 0x53fe3ab1u,
                                              0x53bb8085u, 0x8c49833du, 0x7f4e44a5u, 0x0216d0b1u};   // R^2 mod r

// ------------------------------------------------------------------------------------------
// 512-bit lazy accumulator
// ------------------------------------------------------------------------------------------
struct Wide {
  u64 e[8];  // e[p]: limbs 2p, 2p+1
  u64 o[7];  // o[p]: limbs 2p+1, 2p+2
  u32 k[8];  // k[q]: carry count at limb 8+q
};

__device__ __forceinline__ void wide_zero(Wide& w) {
#pragma unroll
  for (int i = 0; i < 8; i++) w.e[i] = 0;
#pragma unroll
  for (int i = 0; i < 7; i++) w.o[i] = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) w.k[i] = 0;
}

__device__ __forceinline__ u32 lo32(u64 x) { return (u32)x; }
__device__ __forceinline__ u32 hi32(u64 x) { return (u32)(x >> 32); }

// (p0..p3) += {x0,x1,x2,x3} * y as one carry chain over four consecutive 64-bit pairs; carry -> kc.
__device__ __forceinline__ void chain4(u64& p0, u64& p1, u64& p2, u64& p3, u32& kc, u32 x0, u32 x1, u32 x2, u32 x3,
                                       u32 y) {
  asm("{\n\t"
      ".reg .u32 l0, h0, l1, h1, l2, h2, l3, h3;\n\t"
      "mov.b64 {l0, h0}, %0;\n\t"
      "mov.b64 {l1, h1}, %1;\n\t"
      "mov.b64 {l2, h2}, %2;\n\t"
      "mov.b64 {l3, h3}, %3;\n\t"
      "mad.lo.cc.u32 l0, %5, %9, l0;\n\t"
      "madc.hi.cc.u32 h0, %5, %9, h0;\n\t"
      "madc.lo.cc.u32 l1, %6, %9, l1;\n\t"
      "madc.hi.cc.u32 h1, %6, %9, h1;\n\t"
      "madc.lo.cc.u32 l2, %7, %9, l2;\n\t"
      "madc.hi.cc.u32 h2, %7, %9, h2;\n\t"
      "madc.lo.cc.u32 l3, %8, %9, l3;\n\t"
      "madc.hi.cc.u32 h3, %8, %9, h3;\n\t"
      "addc.u32 %4, %4, 0;\n\t"
      "mov.b64 %0, {l0, h0};\n\t"
      "mov.b64 %1, {l1, h1};\n\t"
      "mov.b64 %2, {l2, h2};\n\t"
      "mov.b64 %3, {l3, h3};\n\t"
      "}"
      : "+l"(p0), "+l"(p1), "+l"(p2), "+l"(p3), "+r"(kc)
      : "r"(x0), "r"(x1), "r"(x2), "r"(x3), "r"(y));
}

// Same chain, carry-out discarded (used where the carry is provably zero: it would land at limb 16).
__device__ __forceinline__ void chain4_nc(u64& p0, u64& p1, u64& p2, u64& p3, u32 x0, u32 x1, u32 x2, u32 x3, u32 y) {
  asm("{\n\t"
      ".reg .u32 l0, h0, l1, h1, l2, h2, l3, h3;\n\t"
      "mov.b64 {l0, h0}, %0;\n\t"
      "mov.b64 {l1, h1}, %1;\n\t"
      "mov.b64 {l2, h2}, %2;\n\t"
      "mov.b64 {l3, h3}, %3;\n\t"
      "mad.lo.cc.u32 l0, %4, %8, l0;\n\t"
      "madc.hi.cc.u32 h0, %4, %8, h0;\n\t"
      "madc.lo.cc.u32 l1, %5, %8, l1;\n\t"
      "madc.hi.cc.u32 h1, %5, %8, h1;\n\t"
      "madc.lo.cc.u32 l2, %6, %8, l2;\n\t"
      "madc.hi.cc.u32 h2, %6, %8, h2;\n\t"
      "madc.lo.cc.u32 l3, %7, %8, l3;\n\t"
      "madc.hi.u32 h3, %7, %8, h3;\n\t"
      "mov.b64 %0, {l0, h0};\n\t"
      "mov.b64 %1, {l1, h1};\n\t"
      "mov.b64 %2, {l2, h2};\n\t"
      "mov.b64 %3, {l3, h3};\n\t"
      "}"
      : "+l"(p0), "+l"(p1), "+l"(p2), "+l"(p3)
      : "r"(x0), "r"(x1), "r"(x2), "r"(x3), "r"(y));
}

// Shorter chains for the squaring rows (same structure as chain4).
__device__ __forceinline__ void chain3(u64& p0, u64& p1, u64& p2, u32& kc, u32 x0, u32 x1, u32 x2, u32 y) {
  asm("{\n\t"
      ".reg .u32 l0, h0, l1, h1, l2, h2;\n\t"
      "mov.b64 {l0, h0}, %0;\n\t"
      "mov.b64 {l1, h1}, %1;\n\t"
      "mov.b64 {l2, h2}, %2;\n\t"
      "mad.lo.cc.u32 l0, %4, %7, l0;\n\t"
      "madc.hi.cc.u32 h0, %4, %7, h0;\n\t"
      "madc.lo.cc.u32 l1, %5, %7, l1;\n\t"
      "madc.hi.cc.u32 h1, %5, %7, h1;\n\t"
      "madc.lo.cc.u32 l2, %6, %7, l2;\n\t"
      "madc.hi.cc.u32 h2, %6, %7, h2;\n\t"
      "addc.u32 %3, %3, 0;\n\t"
      "mov.b64 %0, {l0, h0};\n\t"
      "mov.b64 %1, {l1, h1};\n\t"
      "mov.b64 %2, {l2, h2};\n\t"
      "}"
      : "+l"(p0), "+l"(p1), "+l"(p2), "+r"(kc)
      : "r"(x0), "r"(x1), "r"(x2), "r"(y));
}

__device__ __forceinline__ void chain2(u64& p0, u64& p1, u32& kc, u32 x0, u32 x1, u32 y) {
  asm("{\n\t"
      ".reg .u32 l0, h0, l1, h1;\n\t"
      "mov.b64 {l0, h0}, %0;\n\t"
      "mov.b64 {l1, h1}, %1;\n\t"
      "mad.lo.cc.u32 l0, %3, %5, l0;\n\t"
      "madc.hi.cc.u32 h0, %3, %5, h0;\n\t"
      "madc.lo.cc.u32 l1, %4, %5, l1;\n\t"
      "madc.hi.cc.u32 h1, %4, %5, h1;\n\t"
      "addc.u32 %2, %2, 0;\n\t"
      "mov.b64 %0, {l0, h0};\n\t"
      "mov.b64 %1, {l1, h1};\n\t"
      "}"
      : "+l"(p0), "+l"(p1), "+r"(kc)
      : "r"(x0), "r"(x1), "r"(y));
}

__device__ __forceinline__ void chain1(u64& p0, u32& kc, u32 x0, u32 y) {
  asm("{\n\t"
      ".reg .u32 l0, h0;\n\t"
      "mov.b64 {l0, h0}, %0;\n\t"
      "mad.lo.cc.u32 l0, %2, %3, l0;\n\t"
      "madc.hi.cc.u32 h0, %2, %3, h0;\n\t"
      "addc.u32 %1, %1, 0;\n\t"
      "mov.b64 %0, {l0, h0};\n\t"
      "}"
      : "+l"(p0), "+r"(kc)
      : "r"(x0), "r"(y));
}

__device__ __forceinline__ void chain1_nc(u64& p0, u32 x0, u32 y) {
  asm("{\n\t"
      ".reg .u32 l0, h0;\n\t"
      "mov.b64 {l0, h0}, %0;\n\t"
      "mad.lo.cc.u32 l0, %1, %2, l0;\n\t"
      "madc.hi.u32 h0, %1, %2, h0;\n\t"
      "mov.b64 %0, {l0, h0};\n\t"
      "}"
      : "+l"(p0)
      : "r"(x0), "r"(y));
}

// w += x * y * 2^(32*I)   (x: 8 limbs, y: one limb).  Products x[j]*y land at limb I+j.
template <int I>
__device__ __forceinline__ void mac_row(Wide& w, const u32 (&x)[8], u32 y) {
  if constexpr (I % 2 == 0) {
    constexpr int P = I / 2;
    // even j -> even limb I+j -> E pairs P..P+3, carry at limb I+8
    chain4(w.e[P], w.e[P + 1], w.e[P + 2], w.e[P + 3], w.k[I], x[0], x[2], x[4], x[6], y);
    // odd j -> odd limb I+j -> O pairs P..P+3, carry at limb I+9
    if constexpr (I + 1 < 8)
      chain4(w.o[P], w.o[P + 1], w.o[P + 2], w.o[P + 3], w.k[I + 1], x[1], x[3], x[5], x[7], y);
    else
      chain4_nc(w.o[P], w.o[P + 1], w.o[P + 2], w.o[P + 3], x[1], x[3], x[5], x[7], y);
  } else {
    constexpr int PE = (I + 1) / 2, PO = (I - 1) / 2;
    // even j -> odd limb I+j -> O pairs PO..PO+3, carry at limb I+8
    chain4(w.o[PO], w.o[PO + 1], w.o[PO + 2], w.o[PO + 3], w.k[I], x[0], x[2], x[4], x[6], y);
    // odd j -> even limb I+j -> E pairs PE..PE+3, carry at limb I+9
    if constexpr (I + 1 < 8)
      chain4(w.e[PE], w.e[PE + 1], w.e[PE + 2], w.e[PE + 3], w.k[I + 1], x[1], x[3], x[5], x[7], y);
    else
      chain4_nc(w.e[PE], w.e[PE + 1], w.e[PE + 2], w.e[PE + 3], x[1], x[3], x[5], x[7], y);
  }
}

template <int I>
__device__ __forceinline__ u32 wide_limb_e(const Wide& w) {  // limb I of the E accumulator
  return (I % 2 == 0) ? lo32(w.e[I / 2]) : hi32(w.e[I / 2]);
}
template <int I>
__device__ __forceinline__ u32 wide_limb_o(const Wide& w) {  // limb I (>= 1) of the O accumulator
  return (I % 2 == 1) ? lo32(w.o[(I - 1) / 2]) : hi32(w.o[(I - 1) / 2]);
}

// w += a * b (full 8x8 schoolbook, 64 wide multiplies)
__device__ __forceinline__ void wide_mac(Wide& w, const u32 (&a)[8], const u32 (&b)[8]) {
  mac_row<0>(w, a, b[0]);
  mac_row<1>(w, a, b[1]);
  mac_row<2>(w, a, b[2]);
  mac_row<3>(w, a, b[3]);
  mac_row<4>(w, a, b[4]);
  mac_row<5>(w, a, b[5]);
  mac_row<6>(w, a, b[6]);
  mac_row<7>(w, a, b[7]);
}

// w += a^2 for a < 2^255 (every lazy value is < 2r < 2^255): 36 wide multiplies instead of 64.
// a^2 = sum_i a_i * M_i * 2^(64 i) with M_i = a_i + 2 * (a >> 32(i+1)) * 2^32, whose limbs are
// (a_i, a_{i+1} << 1, (2a)_{i+2}, ..., (2a)_7): the doubling of the cross terms is folded into the multiplicand.
// Row I completes limbs 2I and 2I+1.
struct SqrOperand {
  u32 a[8], d[8], e[8];  // a, limbs of 2a, a_j << 1
};

__device__ __forceinline__ void sqr_prepare(SqrOperand& q, const u32 (&a)[8]) {
#pragma unroll
  for (int j = 0; j < 8; j++) {
    q.a[j] = a[j];
    q.e[j] = a[j] << 1;
    q.d[j] = (j == 0) ? q.e[0] : __funnelshift_l(a[j - 1], a[j], 1);
  }
}

template <int I>
__device__ __forceinline__ void sqr_row(Wide& w, const SqrOperand& q) {
  const u32 (&a)[8] = q.a;
  const u32 (&d)[8] = q.d;
  const u32 (&e)[8] = q.e;
  if constexpr (I == 0) {         // M = (a0, e1, d2, d3, d4, d5, d6, d7)
    chain4(w.e[0], w.e[1], w.e[2], w.e[3], w.k[0], a[0], d[2], d[4], d[6], a[0]);
    chain4(w.o[0], w.o[1], w.o[2], w.o[3], w.k[1], e[1], d[3], d[5], d[7], a[0]);
  } else if constexpr (I == 1) {  // offset 2: M = (a1, e2, d3, d4, d5, d6, d7)
    chain4(w.e[1], w.e[2], w.e[3], w.e[4], w.k[2], a[1], d[3], d[5], d[7], a[1]);
    chain3(w.o[1], w.o[2], w.o[3], w.k[1], e[2], d[4], d[6], a[1]);
  } else if constexpr (I == 2) {  // offset 4: M = (a2, e3, d4, d5, d6, d7)
    chain3(w.e[2], w.e[3], w.e[4], w.k[2], a[2], d[4], d[6], a[2]);
    chain3(w.o[2], w.o[3], w.o[4], w.k[3], e[3], d[5], d[7], a[2]);
  } else if constexpr (I == 3) {  // offset 6: M = (a3, e4, d5, d6, d7)
    chain3(w.e[3], w.e[4], w.e[5], w.k[4], a[3], d[5], d[7], a[3]);
    chain2(w.o[3], w.o[4], w.k[3], e[4], d[6], a[3]);
  } else if constexpr (I == 4) {  // offset 8: M = (a4, e5, d6, d7)
    chain2(w.e[4], w.e[5], w.k[4], a[4], d[6], a[4]);
    chain2(w.o[4], w.o[5], w.k[5], e[5], d[7], a[4]);
  } else if constexpr (I == 5) {  // offset 10: M = (a5, e6, d7)
    chain2(w.e[5], w.e[6], w.k[6], a[5], d[7], a[5]);
    chain1(w.o[5], w.k[5], e[6], a[5]);
  } else if constexpr (I == 6) {  // offset 12: M = (a6, e7)
    chain1(w.e[6], w.k[6], a[6], a[6]);
    chain1(w.o[6], w.k[7], e[7], a[6]);
  } else {                        // offset 14: M = (a7)
    chain1_nc(w.e[7], a[7], a[7]);
  }
}

__device__ __forceinline__ void wide_sqr(Wide& w, const u32 (&a)[8]) {
  SqrOperand q;
  sqr_prepare(q, a);
  sqr_row<0>(w, q);
  sqr_row<1>(w, q);
  sqr_row<2>(w, q);
  sqr_row<3>(w, q);
  sqr_row<4>(w, q);
  sqr_row<5>(w, q);
  sqr_row<6>(w, q);
  sqr_row<7>(w, q);
}

// w += a * R  (places a at limbs 8..15; adds a Montgomery-form constant to a pending dot product)
__device__ __forceinline__ void wide_add_hi(Wide& w, const u32 (&a)[8]) {
  // the E pairs 4..7 hold limbs 8..15; add as four 64-bit values with carries counted into k
  asm("{\n\t"
      ".reg .u32 l0, h0, l1, h1, l2, h2, l3, h3;\n\t"
      "mov.b64 {l0, h0}, %0;\n\t"
      "mov.b64 {l1, h1}, %1;\n\t"
      "mov.b64 {l2, h2}, %2;\n\t"
      "mov.b64 {l3, h3}, %3;\n\t"
      "add.cc.u32 l0, l0, %4;\n\t"
      "addc.cc.u32 h0, h0, %5;\n\t"
      "addc.cc.u32 l1, l1, %6;\n\t"
      "addc.cc.u32 h1, h1, %7;\n\t"
      "addc.cc.u32 l2, l2, %8;\n\t"
      "addc.cc.u32 h2, h2, %9;\n\t"
      "addc.cc.u32 l3, l3, %10;\n\t"
      "addc.u32 h3, h3, %11;\n\t"
      "mov.b64 %0, {l0, h0};\n\t"
      "mov.b64 %1, {l1, h1};\n\t"
      "mov.b64 %2, {l2, h2};\n\t"
      "mov.b64 %3, {l3, h3};\n\t"
      "}"
      : "+l"(w.e[4]), "+l"(w.e[5]), "+l"(w.e[6]), "+l"(w.e[7])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]));
}

template <int I>
__device__ __forceinline__ void redc_row(Wide& w, u32& c) {
  const u32 P[8] = {GCP_P0, GCP_P1, GCP_P2, GCP_P3, GCP_P4, GCP_P5, GCP_P6, GCP_P7};
  u32 s;
  if constexpr (I == 0)
    s = wide_limb_e<0>(w);
  else
    s = wide_limb_e<I>(w) + wide_limb_o<I>(w) + c;
  // (an ALU-pipe shift-add form of this multiply was measured 1.7 % slower: it lengthens the row-to-row critical path;
  //  skipping the m * P[0] product - P[0] = 2^32 - 2^28 + 1, its low word cancels s and the word carried into limb I+1
  //  is m - (m >> 4) + [(m << 28) < m] + floor(S / 2^32), exact and parity-green - makes a reduction 56 wide multiplies
  //  but was measured 17 % slower on Hash2 (135 -> 112 M/s): the extra compares exhaust the 7 predicate registers that
  //  carry the chains, ptxas spills predicates into a GPR with LOP3s and the path kernel's registers go 106 -> 124.)
  u32 m = s * GCP_NP;
  mac_row<I>(w, P, m);
  // limb I of E + O + c is now 0 mod 2^32; it is either 0 or exactly 2^32
  u32 t;
  if constexpr (I == 0)
    t = wide_limb_e<0>(w);
  else
    t = wide_limb_e<I>(w) | wide_limb_o<I>(w) | c;
  c = (t != 0) ? 1u : 0u;
}

// Montgomery reduction: r = w / 2^256 mod r, not fully reduced.
// Bound: r < w / 2^256 + r_mod.  The register-resident kernels keep w < 6.2 r^2 (result < 3.2 r_mod); the generic
// Poseidon rows go up to 22 r^2 (result < 5.2 r_mod < 2^256, see poseidon.cuh).
__device__ __forceinline__ void wide_redc_finish(Wide& w, u32 c, u32 (&r)[8]);

__device__ __forceinline__ void wide_redc(Wide& w, u32 (&r)[8]) {
  u32 c = 0;
  redc_row<0>(w, c);
  redc_row<1>(w, c);
  redc_row<2>(w, c);
  redc_row<3>(w, c);
  redc_row<4>(w, c);
  redc_row<5>(w, c);
  redc_row<6>(w, c);
  redc_row<7>(w, c);
  wide_redc_finish(w, c, r);
}

// after the eight reduction rows: r = E[8..15] + O[8..15] + K[8..15] + c
__device__ __forceinline__ void wide_redc_finish(Wide& w, u32 c, u32 (&r)[8]) {
  u32 e8 = wide_limb_e<8>(w), e9 = wide_limb_e<9>(w), e10 = wide_limb_e<10>(w), e11 = wide_limb_e<11>(w);
  u32 e12 = wide_limb_e<12>(w), e13 = wide_limb_e<13>(w), e14 = wide_limb_e<14>(w), e15 = wide_limb_e<15>(w);
  u32 o8 = wide_limb_o<8>(w), o9 = wide_limb_o<9>(w), o10 = wide_limb_o<10>(w), o11 = wide_limb_o<11>(w);
  u32 o12 = wide_limb_o<12>(w), o13 = wide_limb_o<13>(w), o14 = wide_limb_o<14>(w);
  asm("{\n\t"
      ".reg .u32 t;\n\t"
      "add.cc.u32 t, %8, 0xffffffff;\n\t"  // carry flag := c
      "addc.cc.u32 %0, %9, %17;\n\t"
      "addc.cc.u32 %1, %10, %18;\n\t"
      "addc.cc.u32 %2, %11, %19;\n\t"
      "addc.cc.u32 %3, %12, %20;\n\t"
      "addc.cc.u32 %4, %13, %21;\n\t"
      "addc.cc.u32 %5, %14, %22;\n\t"
      "addc.cc.u32 %6, %15, %23;\n\t"
      "addc.u32 %7, %16, 0;\n\t"
      "}"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
      : "r"(c), "r"(e8), "r"(e9), "r"(e10), "r"(e11), "r"(e12), "r"(e13), "r"(e14), "r"(e15), "r"(o8), "r"(o9),
        "r"(o10), "r"(o11), "r"(o12), "r"(o13), "r"(o14));
  asm("add.cc.u32 %0, %0, %8;\n\t"
      "addc.cc.u32 %1, %1, %9;\n\t"
      "addc.cc.u32 %2, %2, %10;\n\t"
      "addc.cc.u32 %3, %3, %11;\n\t"
      "addc.cc.u32 %4, %4, %12;\n\t"
      "addc.cc.u32 %5, %5, %13;\n\t"
      "addc.cc.u32 %6, %6, %14;\n\t"
      "addc.u32 %7, %7, %15;\n\t"
      : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7])
      : "r"(w.k[0]), "r"(w.k[1]), "r"(w.k[2]), "r"(w.k[3]), "r"(w.k[4]), "r"(w.k[5]), "r"(w.k[6]), "r"(w.k[7]));
}

// ------------------------------------------------------------------------------------------
// 256-bit helpers
// ------------------------------------------------------------------------------------------
// r = a - b, returns borrow (1 if a < b)
__device__ __forceinline__ u32 sub256(u32 (&r)[8], const u32 (&a)[8], const u32 (&b)[8]) {
  u32 borrow;
  asm("sub.cc.u32 %0, %9, %17;\n\t"
      "subc.cc.u32 %1, %10, %18;\n\t"
      "subc.cc.u32 %2, %11, %19;\n\t"
      "subc.cc.u32 %3, %12, %20;\n\t"
      "subc.cc.u32 %4, %13, %21;\n\t"
      "subc.cc.u32 %5, %14, %22;\n\t"
      "subc.cc.u32 %6, %15, %23;\n\t"
      "subc.cc.u32 %7, %16, %24;\n\t"
      "subc.u32 %8, 0, 0;\n\t"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(borrow)
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]), "r"(b[0]), "r"(b[1]),
        "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]));
  return borrow & 1u;
}

__device__ __forceinline__ void add256(u32 (&r)[8], const u32 (&a)[8], const u32 (&b)[8]) {
  asm("add.cc.u32 %0, %8, %16;\n\t"
      "addc.cc.u32 %1, %9, %17;\n\t"
      "addc.cc.u32 %2, %10, %18;\n\t"
      "addc.cc.u32 %3, %11, %19;\n\t"
      "addc.cc.u32 %4, %12, %20;\n\t"
      "addc.cc.u32 %5, %13, %21;\n\t"
      "addc.cc.u32 %6, %14, %22;\n\t"
      "addc.u32 %7, %15, %23;\n\t"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]), "r"(b[0]), "r"(b[1]),
        "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]));
}

// a := a - m if a >= m  (a < 2m on entry)
__device__ __forceinline__ void cond_sub(u32 (&a)[8], const u32 (&m)[8]) {
  u32 t[8];
  u32 borrow = sub256(t, a, m);
#pragma unroll
  for (int i = 0; i < 8; i++) a[i] = borrow ? a[i] : t[i];
}

#define GCP_2P_LIMBS {0xe0000002u, 0x87c3eb27u, 0xf372e122u, 0x5067d090u, 0x0302b0bau, 0x70a08b6du, 0xc2634053u, 0x60c89ce5u}
#define GCP_P_LIMBS {GCP_P0, GCP_P1, GCP_P2, GCP_P3, GCP_P4, GCP_P5, GCP_P6, GCP_P7}

// ------------------------------------------------------------------------------------------
// Fr operations on lazy Montgomery values (all inputs < 2r unless stated, all outputs < 2r)
// ------------------------------------------------------------------------------------------
// r = a * b / R.  inputs < 2r  =>  output < 1.76 r  (4 r^2 / 2^256 + r)
__device__ __forceinline__ void fr_mul(u32 (&r)[8], const u32 (&a)[8], const u32 (&b)[8]) {
  Wide w;
  wide_zero(w);
  // product row i completes limb i, so reduction row i can follow it immediately: the serial m -> m*r -> next m chain
  // of the reduction then overlaps with the independent product rows that come after it
  u32 c = 0;
  mac_row<0>(w, a, b[0]); redc_row<0>(w, c);
  mac_row<1>(w, a, b[1]); redc_row<1>(w, c);
  mac_row<2>(w, a, b[2]); redc_row<2>(w, c);
  mac_row<3>(w, a, b[3]); redc_row<3>(w, c);
  mac_row<4>(w, a, b[4]); redc_row<4>(w, c);
  mac_row<5>(w, a, b[5]); redc_row<5>(w, c);
  mac_row<6>(w, a, b[6]); redc_row<6>(w, c);
  mac_row<7>(w, a, b[7]); redc_row<7>(w, c);
  wide_redc_finish(w, c, r);
}

// r = a * a / R, a < 2r
__device__ __forceinline__ void fr_sqr(u32 (&r)[8], const u32 (&a)[8]) {
  // (interleaving the squaring rows with reduction rows, as fr_mul does, measured 1.5 % slower)
  Wide w;
  wide_zero(w);
  wide_sqr(w, a);
  wide_redc(w, r);
}

// two independent squarings with their reductions advancing together (two chains in flight)
__device__ __forceinline__ void fr_sqr2(u32 (&r1)[8], const u32 (&a1)[8], u32 (&r2)[8], const u32 (&a2)[8]) {
  Wide w1, w2;
  wide_zero(w1);
  wide_zero(w2);
  wide_sqr(w1, a1);
  wide_sqr(w2, a2);
  u32 c1 = 0, c2 = 0;
  redc_row<0>(w1, c1); redc_row<0>(w2, c2);
  redc_row<1>(w1, c1); redc_row<1>(w2, c2);
  redc_row<2>(w1, c1); redc_row<2>(w2, c2);
  redc_row<3>(w1, c1); redc_row<3>(w2, c2);
  redc_row<4>(w1, c1); redc_row<4>(w2, c2);
  redc_row<5>(w1, c1); redc_row<5>(w2, c2);
  redc_row<6>(w1, c1); redc_row<6>(w2, c2);
  redc_row<7>(w1, c1); redc_row<7>(w2, c2);
  wide_redc_finish(w1, c1, r1);
  wide_redc_finish(w2, c2, r2);
}

// Several independent multiplications advanced row by row together (CIOS order each), so that a thread has that
// many carry / reduction chains in flight.  Used where the operands are all ready at once (curve addition formulas).
#define GCP_MULN_ROW2(I)                                \
  mac_row<I>(w1, a1, b1[I]); redc_row<I>(w1, c1);       \
  mac_row<I>(w2, a2, b2[I]); redc_row<I>(w2, c2);
__device__ __forceinline__ void fr_mul2(u32 (&r1)[8], const u32 (&a1)[8], const u32 (&b1)[8], u32 (&r2)[8], const u32 (&a2)[8],
                                        const u32 (&b2)[8]) {
  Wide w1, w2;
  wide_zero(w1);
  wide_zero(w2);
  u32 c1 = 0, c2 = 0;
  GCP_MULN_ROW2(0) GCP_MULN_ROW2(1) GCP_MULN_ROW2(2) GCP_MULN_ROW2(3) GCP_MULN_ROW2(4) GCP_MULN_ROW2(5) GCP_MULN_ROW2(6) GCP_MULN_ROW2(7)
  wide_redc_finish(w1, c1, r1);
  wide_redc_finish(w2, c2, r2);
}
#undef GCP_MULN_ROW2

#define GCP_MULN_ROW3(I)                                \
  mac_row<I>(w1, a1, b1[I]); redc_row<I>(w1, c1);       \
  mac_row<I>(w2, a2, b2[I]); redc_row<I>(w2, c2);       \
  mac_row<I>(w3, a3, b3[I]); redc_row<I>(w3, c3);
__device__ __forceinline__ void fr_mul3(u32 (&r1)[8], const u32 (&a1)[8], const u32 (&b1)[8], u32 (&r2)[8], const u32 (&a2)[8],
                                        const u32 (&b2)[8], u32 (&r3)[8], const u32 (&a3)[8], const u32 (&b3)[8]) {
  Wide w1, w2, w3;
  wide_zero(w1);
  wide_zero(w2);
  wide_zero(w3);
  u32 c1 = 0, c2 = 0, c3 = 0;
  GCP_MULN_ROW3(0) GCP_MULN_ROW3(1) GCP_MULN_ROW3(2) GCP_MULN_ROW3(3) GCP_MULN_ROW3(4) GCP_MULN_ROW3(5) GCP_MULN_ROW3(6) GCP_MULN_ROW3(7)
  wide_redc_finish(w1, c1, r1);
  wide_redc_finish(w2, c2, r2);
  wide_redc_finish(w3, c3, r3);
}
#undef GCP_MULN_ROW3

// r = a + b mod 2r
__device__ __forceinline__ void fr_add(u32 (&r)[8], const u32 (&a)[8], const u32 (&b)[8]) {
  const u32 P2[8] = GCP_2P_LIMBS;
  add256(r, a, b);  // < 4r < 2^256
  cond_sub(r, P2);
}

// r = a - b mod 2r
__device__ __forceinline__ void fr_sub(u32 (&r)[8], const u32 (&a)[8], const u32 (&b)[8]) {
  const u32 P2[8] = GCP_2P_LIMBS;
  u32 t[8], u[8];
  u32 borrow = sub256(t, a, b);
  add256(u, t, P2);
#pragma unroll
  for (int i = 0; i < 8; i++) r[i] = borrow ? u[i] : t[i];
}

// r = -a mod 2r   (a < 2r; 0 stays 0... as 2r - a which is congruent)
__device__ __forceinline__ void fr_neg(u32 (&r)[8], const u32 (&a)[8]) {
  const u32 P2[8] = GCP_2P_LIMBS;
  sub256(r, P2, a);  // in (0, 2r]; 2r only when a == 0
  cond_sub(r, P2);
}

__device__ __forceinline__ void fr_double(u32 (&r)[8], const u32 (&a)[8]) { fr_add(r, a, a); }

// lazy (< 4r) -> canonical [0, r)
__device__ __forceinline__ void fr_canon(u32 (&a)[8]) {
  const u32 P2[8] = GCP_2P_LIMBS;
  const u32 P1[8] = GCP_P_LIMBS;
  cond_sub(a, P2);
  cond_sub(a, P1);
}

// standard canonical integer -> lazy Montgomery
__device__ __forceinline__ void fr_to_mont(u32 (&r)[8], const u32 (&a)[8]) {
  const u32 R2[8] = {0xae216da7u, 0x1bb8e645u, 0xe35c59e3u, 0x53fe3ab1u, 0x53bb8085u, 0x8c49833du, 0x7f4e44a5u, 0x0216d0b1u};
  fr_mul(r, a, R2);
}

// lazy Montgomery -> canonical standard integer
__device__ __forceinline__ void fr_from_mont(u32 (&r)[8], const u32 (&a)[8]) {
  Wide w;
  wide_zero(w);
#pragma unroll
  for (int p = 0; p < 4; p++) w.e[p] = ((u64)a[2 * p + 1] << 32) | a[2 * p];
  wide_redc(w, r);  // < a / 2^256 + r <= r  (a < 2^256)
  const u32 P1[8] = GCP_P_LIMBS;
  cond_sub(r, P1);
}

// r = a / R mod r for any a < 2^256 (one Montgomery reduction, no product); output <= r
__device__ __forceinline__ void fr_redc(u32 (&r)[8], const u32 (&a)[8]) {
  Wide w;
  wide_zero(w);
#pragma unroll
  for (int p = 0; p < 4; p++) w.e[p] = ((u64)a[2 * p + 1] << 32) | a[2 * p];
  wide_redc(w, r);
}

__device__ __forceinline__ bool fr_is_canonical(const u32 (&a)[8]) {  // a < r
  const u32 P1[8] = GCP_P_LIMBS;
  u32 t[8];
  return sub256(t, a, P1) != 0;
}

__device__ __forceinline__ bool is_zero256(const u32 (&a)[8]) {
  return (a[0] | a[1] | a[2] | a[3] | a[4] | a[5] | a[6] | a[7]) == 0;
}

__device__ __forceinline__ bool eq256(const u32 (&a)[8], const u32 (&b)[8]) {
  u32 d = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) d |= a[i] ^ b[i];
  return d == 0;
}

__device__ __forceinline__ void fr_copy(u32 (&r)[8], const u32 (&a)[8]) {
#pragma unroll
  for (int l = 0; l < 8; l++) r[l] = a[l];
}
__device__ __forceinline__ void fr_set_zero(u32 (&r)[8]) {
#pragma unroll
  for (int l = 0; l < 8; l++) r[l] = 0;
}
__device__ __forceinline__ void fr_set_one(u32 (&r)[8]) {
  const u32 one[8] = {0x4ffffffbu, 0xac96341cu, 0x9f60cd29u, 0x36fc7695u, 0x7879462eu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u};  // R mod r
#pragma unroll
  for (int l = 0; l < 8; l++) r[l] = one[l];
}

// 128-bit vectorised global access of one 32-byte element (address must be 16-byte aligned)
__device__ __forceinline__ void load_fr(u32 (&r)[8], const void* p) {
  const uint4* q = reinterpret_cast<const uint4*>(p);
  uint4 x = __ldg(q), y = __ldg(q + 1);
  r[0] = x.x; r[1] = x.y; r[2] = x.z; r[3] = x.w;
  r[4] = y.x; r[5] = y.y; r[6] = y.z; r[7] = y.w;
}

// the same through the coherent path (for data this kernel wrote itself)
__device__ __forceinline__ void load_fr_plain(u32 (&r)[8], const void* p) {
  const uint4* q = reinterpret_cast<const uint4*>(p);
  uint4 x = q[0], y = q[1];
  r[0] = x.x; r[1] = x.y; r[2] = x.z; r[3] = x.w;
  r[4] = y.x; r[5] = y.y; r[6] = y.z; r[7] = y.w;
}

__device__ __forceinline__ void store_fr(void* p, const u32 (&a)[8]) {
  uint4* q = reinterpret_cast<uint4*>(p);
  q[0] = make_uint4(a[0], a[1], a[2], a[3]);
  q[1] = make_uint4(a[4], a[5], a[6], a[7]);
}

}  // namespace gcp
