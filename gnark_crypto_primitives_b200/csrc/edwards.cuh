// Twisted Edwards arithmetic for gnark / gnark-crypto's BN254 embedded curve (a = -1, "reduced" BabyJubJub),
// device side.  The reference calls it through gnark's std/algebra/native/twistededwards (un-vendored):
// curve.Add / Neg / ScalarMul / AssertIsOnCurve at /root/reference/elgamal/ciphertext.go:24-46,
// elgamal/encrypt.go:42-94, elgamal/mul.go:76-166.  Affine law:
//     x3 = (x1 y2 + y1 x2) / (1 + d x1 x2 y1 y2),   y3 = (y1 y2 + x1 x2) / (1 - d x1 x2 y1 y2).
// Kernels keep points in extended coordinates (X:Y:Z:T), T = XY/Z, and use the unified a = -1 addition
// (Hisil-Wong-Carter-Dawson 2008, "add-2008-hwcd-3"), which is the SAME rational function as the affine law
// (it never uses the curve equation), so results agree with the reference for any inputs with non-zero
// denominators; Z3 = 0 exactly when an affine denominator is 0.  Doubling (dbl-2008-hwcd) assumes the point is
// on the curve and is used only on validated points.
#pragma once
#include "fr.cuh"

namespace gcp {

#define GCP_ED_D_MONT {0x504f718du, 0x5c3b8876u, 0x984346b4u, 0x50be2c72u, 0x59126675u, 0x4783751fu, 0xa7a1c091u, 0x305ff669u}
#define GCP_ED_2D_MONT {0xb09ee319u, 0x74951b58u, 0xb6cd1cd7u, 0x7948709cu, 0x30a3748du, 0xd6b6a488u, 0x6e11e0f8u, 0x305b9e60u}
#define GCP_ED_2D_R2 {0x9d9efa30u, 0xa064f144u, 0x19d4c427u, 0x338939eeu, 0x072e9f21u, 0x2483f1a5u, 0x6271a8c1u, 0x06e5885du}
#define GCP_ED_2D_R3 {0x05326973u, 0x900e846eu, 0x05aa9436u, 0xed233c6eu, 0xb6f3823du, 0x391d5278u, 0xf9451009u, 0x1c0c9c01u}
#define GCP_FR_ONE_MONT {0x4ffffffbu, 0xac96341cu, 0x9f60cd29u, 0x36fc7695u, 0x7879462eu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u}
#define GCP_FR_R2 {0xae216da7u, 0x1bb8e645u, 0xe35c59e3u, 0x53fe3ab1u, 0x53bb8085u, 0x8c49833du, 0x7f4e44a5u, 0x0216d0b1u}
// base point of gnark-crypto's bn254 twisted Edwards curve, standard form
#define GCP_ED_GX {0xfe553f9fu, 0xf1f9195au, 0xe6f2a277u, 0x377c749au, 0xc199e94cu, 0x8a4eb7a4u, 0x6ce19d35u, 0x1561ff83u}
#define GCP_ED_GY {0x872d7d8bu, 0x4b3c257au, 0xb9e13377u, 0xfce0051fu, 0xd16bf9edu, 0x25572e1cu, 0xf7a0b249u, 0x25797203u}

struct ExtPoint {  // lazy Montgomery coordinates
  u32 X[8], Y[8], Z[8], T[8];
};

struct NielsPoint {  // affine point prepared for mixed addition: (y - x, y + x, 2 d x y), Montgomery, canonical
  u32 ymx[8], ypx[8], t2d[8];
};

__device__ __forceinline__ void ext_identity(ExtPoint& p) {
  fr_set_zero(p.X);
  fr_set_one(p.Y);
  fr_set_one(p.Z);
  fr_set_zero(p.T);
}

// P += N (mixed addition, 7 multiplies).  CHAINS = true advances the independent products row by row together
// (three, then two and two): +14 % in the tally kernel, whose loop has registers to spare; the window-table kernels
// are register-bound and measured 1 % slower with it, so they keep one chain at a time.
// z_over_r: the Niels point is that of a projectively rescaled Q = (x/R : y/R : 1/R : xy/R) (standard-form coordinates
// read AS Montgomery representations, see tally_partial_kernel), so D = 2 Z1 Z2 is 2 Z1 / R: one reduction, no product.
template <bool CHAINS = false>
__device__ __forceinline__ void ext_add_niels(ExtPoint& p, const NielsPoint& n, bool z_over_r = false) {
  u32 a[8], b[8], c[8], d[8], e[8], f[8], g[8], h[8], t[8], u[8];
  fr_sub(t, p.Y, p.X);
  fr_add(u, p.Y, p.X);
  if constexpr (CHAINS) {
    fr_mul3(a, t, n.ymx, b, u, n.ypx, c, p.T, n.t2d);
  } else {
    fr_mul(a, t, n.ymx);
    fr_mul(b, u, n.ypx);
    fr_mul(c, p.T, n.t2d);
  }
  fr_add(d, p.Z, p.Z);
  if (z_over_r) fr_redc(d, d);
  fr_sub(e, b, a);
  fr_sub(f, d, c);
  fr_add(g, d, c);
  fr_add(h, b, a);
  if constexpr (CHAINS) {
    fr_mul2(p.X, e, f, p.Y, g, h);
    fr_mul2(p.T, e, h, p.Z, f, g);
  } else {
    fr_mul(p.X, e, f);
    fr_mul(p.Y, g, h);
    fr_mul(p.T, e, h);
    fr_mul(p.Z, f, g);
  }
}

// P += Q (both extended, 9 multiplies)
__device__ __forceinline__ void ext_add(ExtPoint& p, const ExtPoint& q) {
  const u32 d2[8] = GCP_ED_2D_MONT;
  u32 a[8], b[8], c[8], d[8], e[8], f[8], g[8], h[8], t[8], u[8];
  fr_sub(t, p.Y, p.X);
  fr_sub(u, q.Y, q.X);
  fr_mul(a, t, u);
  fr_add(t, p.Y, p.X);
  fr_add(u, q.Y, q.X);
  fr_mul(b, t, u);
  fr_mul(t, p.T, q.T);
  fr_mul(c, t, d2);
  fr_mul(t, p.Z, q.Z);
  fr_add(d, t, t);
  fr_sub(e, b, a);
  fr_sub(f, d, c);
  fr_add(g, d, c);
  fr_add(h, b, a);
  fr_mul(p.X, e, f);
  fr_mul(p.Y, g, h);
  fr_mul(p.T, e, h);
  fr_mul(p.Z, f, g);
}

// P = 2P (on-curve points only; 4 multiplies + 4 squarings).  WITH_T = false leaves T stale (3 multiplies): valid
// when the next operation is another doubling, which does not read T.
template <bool WITH_T = true>
__device__ __forceinline__ void ext_double(ExtPoint& p) {
  u32 a[8], b[8], c[8], e[8], f[8], g[8], h[8], t[8];
  fr_sqr(a, p.X);
  fr_sqr(b, p.Y);
  fr_sqr(t, p.Z);
  fr_add(c, t, t);
  fr_add(t, p.X, p.Y);
  fr_sqr(e, t);
  fr_sub(e, e, a);
  fr_sub(e, e, b);   // E = 2XY
  fr_sub(g, b, a);   // G = -A + B  (a = -1)
  fr_sub(f, g, c);   // F = G - C
  fr_add(t, a, b);
  fr_neg(h, t);      // H = -A - B
  fr_mul(p.X, e, f);
  fr_mul(p.Y, g, h);
  if constexpr (WITH_T) fr_mul(p.T, e, h);
  fr_mul(p.Z, f, g);
}

// affine (x, y) in lazy Montgomery form -> extended
__device__ __forceinline__ void ext_from_affine(ExtPoint& p, const u32 (&x)[8], const u32 (&y)[8]) {
  fr_copy(p.X, x);
  fr_copy(p.Y, y);
  fr_set_one(p.Z);
  fr_mul(p.T, x, y);
}

// Out-of-line multiplier bodies with operands and results as structs BY VALUE (the CUDA ABI hands them over in
// registers: no local memory).  Used where one shared body beats inlined copies (varbase_reg.cuh; measured per kernel).
struct E8 {
  u32 v[8];
};
struct E16 {
  u32 a[8], b[8];
};

__device__ __noinline__ E16 fr_mul2_ool(E16 x, E16 y) {  // (x.a * y.a, x.b * y.b)
  E16 r;
  fr_mul2(r.a, x.a, y.a, r.b, x.b, y.b);
  return r;
}
__device__ __noinline__ E16 fr_sqr2_ool(E16 x) {  // (x.a^2, x.b^2)
  E16 r;
  fr_sqr2(r.a, x.a, r.b, x.b);
  return r;
}
__device__ __noinline__ E8 fr_mul_ool(E8 a, E8 b) {
  E8 r;
  fr_mul(r.v, a.v, b.v);
  return r;
}

// a^(r-2): Fermat inversion by square-and-multiply over the fixed exponent (0 -> 0).
__device__ __noinline__ void fr_inv(u32 (&r)[8], const u32 (&a)[8]) {
  const u32 e[8] = {0xefffffffu, 0x43e1f593u, 0x79b97091u, 0x2833e848u, 0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
  u32 acc[8], base[8];
  fr_set_one(acc);
  fr_copy(base, a);
#pragma unroll 1
  for (int i = 0; i < 254; i++) {
    if ((e[i >> 5] >> (i & 31)) & 1u) fr_mul(acc, acc, base);
    fr_sqr(base, base);
  }
  fr_copy(r, acc);
}

// -x^2 + y^2 == 1 + d x^2 y^2   (curve.AssertIsOnCurve; inputs lazy Montgomery)
__device__ __forceinline__ bool ed_is_on_curve(const u32 (&x)[8], const u32 (&y)[8]) {
  const u32 dm[8] = GCP_ED_D_MONT;
  u32 x2[8], y2[8], l[8], r[8], one[8];
  fr_sqr(x2, x);
  fr_sqr(y2, y);
  fr_sub(l, y2, x2);
  fr_mul(r, x2, y2);
  fr_mul(r, r, dm);
  fr_set_one(one);
  fr_add(r, r, one);
  fr_canon(l);
  fr_canon(r);
  return eq256(l, r);
}

// window w of a 256-bit little-endian scalar, WBITS bits wide (WBITS <= 16)
template <int WBITS>
__device__ __forceinline__ u32 scalar_window(const u32 (&k)[8], int w) {
  int bit = w * WBITS;
  int limb = bit >> 5, sh = bit & 31;
  u32 v = k[limb] >> sh;
  if (sh + WBITS > 32 && limb + 1 < 8) v |= k[limb + 1] << (32 - sh);
  return v & ((1u << WBITS) - 1u);
}

__device__ __forceinline__ void load_niels(NielsPoint& n, const u32* p) {
  load_fr(n.ymx, p);
  load_fr(n.ypx, p + 8);
  load_fr(n.t2d, p + 16);
}

}  // namespace gcp
