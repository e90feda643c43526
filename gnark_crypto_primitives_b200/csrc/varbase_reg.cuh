// Variable-base window kernel, register-resident form.
//
// Same work as varbase_window_kernel (varbase.cuh): signed 4-bit windows, four doublings and one cached addition per
// base and window, the per-item table [1..8]P in global scratch.  The interpreter keeps a thread's working set in a
// shared-memory register file so that ONE multiplier body serves every formula (the formulas inlined are ~55 KB of SASS
// and fetch-bound); its price is the round trip of every operand through shared memory and the decode / dispatch of 26
// micro-ops per window.  Here the point formulas are ordinary code on registers and the multiplier bodies are
// out-of-line FUNCTIONS whose operands and results are structs passed BY VALUE: the CUDA ABI then hands them over in
// registers (no local memory: 0 LDL / STL in the SASS), which round 1's out-of-line bodies - arrays by pointer - could
// not do.  Code: one two-chain multiply, one two-chain squaring, one single multiply + the formulas' add / sub glue.
#pragma once
#include "varbase.cuh"

namespace gcp {

// P = 2P (dbl-2008-hwcd, a = -1); with_t = false leaves T stale (the next operation is another doubling)
__device__ __forceinline__ void vr_double(ExtPoint& p, bool with_t) {
  E16 in, sq1, sq2;
  fr_copy(in.a, p.X);
  fr_copy(in.b, p.Y);
  sq1 = fr_sqr2_ool(in);                 // A = X^2, B = Y^2
  fr_copy(in.a, p.Z);
  fr_add(in.b, p.X, p.Y);
  sq2 = fr_sqr2_ool(in);                 // Z^2, (X + Y)^2
  u32 e[8], f[8], g[8], h[8], t[8];
  fr_sub(t, sq2.b, sq1.a);
  fr_sub(e, t, sq1.b);                   // E = (X+Y)^2 - A - B
  fr_sub(g, sq1.b, sq1.a);               // G = B - A
  fr_add(t, sq2.a, sq2.a);
  fr_sub(f, g, t);                       // F = G - 2 Z^2
  fr_add(t, sq1.a, sq1.b);
  fr_neg(h, t);                          // H = -(A + B)
  E16 l, r, o;
  fr_copy(l.a, e);
  fr_copy(r.a, f);
  fr_copy(l.b, g);
  fr_copy(r.b, h);
  o = fr_mul2_ool(l, r);                 // X = E F, Y = G H
  fr_copy(p.X, o.a);
  fr_copy(p.Y, o.b);
  if (with_t) {
    fr_copy(l.a, f);
    fr_copy(r.a, g);
    fr_copy(l.b, e);
    fr_copy(r.b, h);
    o = fr_mul2_ool(l, r);               // Z = F G, T = E H
    fr_copy(p.Z, o.a);
    fr_copy(p.T, o.b);
  } else {
    E8 x, y, z;
    fr_copy(x.v, f);
    fr_copy(y.v, g);
    z = fr_mul_ool(x, y);
    fr_copy(p.Z, z.v);
  }
}

// P += Q with Q cached as (Y-X, Y+X, 2dT, 2Z) at q (32 words); neg: add -Q = (-x, y): the first two swap, 2dT changes sign
__device__ __forceinline__ void vr_add_cached(ExtPoint& p, const u32* __restrict__ q, bool neg) {
  E16 l, r, ab, cd;
  fr_sub(l.a, p.Y, p.X);
  fr_add(l.b, p.Y, p.X);
  load_fr_plain(r.a, q + (neg ? 8 : 0));
  load_fr_plain(r.b, q + (neg ? 0 : 8));
  ab = fr_mul2_ool(l, r);                // A = (Y-X) q0, B = (Y+X) q1
  fr_copy(l.a, p.T);
  fr_copy(l.b, p.Z);
  load_fr_plain(r.a, q + 16);
  load_fr_plain(r.b, q + 24);
  cd = fr_mul2_ool(l, r);                // C = T q2, D = Z q3
  if (neg) {
    u32 t[8];
    fr_neg(t, cd.a);
    fr_copy(cd.a, t);
  }
  u32 e[8], f[8], g[8], h[8];
  fr_sub(e, ab.b, ab.a);
  fr_add(h, ab.b, ab.a);
  fr_sub(f, cd.b, cd.a);
  fr_add(g, cd.b, cd.a);
  E16 o;
  fr_copy(l.a, e);
  fr_copy(r.a, f);
  fr_copy(l.b, g);
  fr_copy(r.b, h);
  o = fr_mul2_ool(l, r);                 // X = E F, Y = G H
  fr_copy(p.X, o.a);
  fr_copy(p.Y, o.b);
  fr_copy(l.a, e);
  fr_copy(r.a, h);
  fr_copy(l.b, f);
  fr_copy(r.b, g);
  o = fr_mul2_ool(l, r);                 // T = E H, Z = F G
  fr_copy(p.T, o.a);
  fr_copy(p.Z, o.b);
}

// table entry <- cached(P) = (Y-X, Y+X, 2d T, 2Z)
__device__ __forceinline__ void vr_store_cached(u32* __restrict__ q, const ExtPoint& p) {
  const u32 d2[8] = GCP_ED_2D_MONT;
  u32 t[8];
  fr_sub(t, p.Y, p.X);
  store_fr(q, t);
  fr_add(t, p.Y, p.X);
  store_fr(q + 8, t);
  E8 x, y, z;
  fr_copy(x.v, p.T);
  fr_copy(y.v, d2);
  z = fr_mul_ool(x, y);
  store_fr(q + 16, z.v);
  fr_add(t, p.Z, p.Z);
  store_fr(q + 24, t);
}

template <int MIN_BLOCKS>
__global__ void __launch_bounds__(VB_THREADS, MIN_BLOCKS) varbase_window_reg_kernel(VarbaseArgs a) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= a.n) return;
  const bool live = !a.status || a.status[idx] == GCP_STATUS_OK;
  const int nb = a.n_bases;
  u32* o = a.out + idx * 32;
  ExtPoint p;
  ext_identity(p);
  if (live) {
    u32 mag0[8], sgn0[2], mag1[8], sgn1[2];
    {
      u32 k[8];
      load_fr(k, a.scalars[0] + idx * 8);
      vb_recode(mag0, sgn0, k);
      if (nb > 1) load_fr(k, a.scalars[1] + idx * 8);
      vb_recode(mag1, sgn1, k);
    }
    u32* const tab = a.table + idx * (size_t)nb * VB_TABLE_WORDS;
    // per base: table[e] = cached([e + 1] P), e = 0..7: P, 2P = dbl(P), then (e+1)P = eP + P
#pragma unroll 1
    for (int base = 0; base < nb; base++) {
      u32* const tb = tab + (size_t)base * 8 * 32;
      const u32* bp = a.bases + (idx * (size_t)nb + base) * 32;
      load_fr(p.X, bp);
      load_fr(p.Y, bp + 8);
      load_fr(p.Z, bp + 16);
      load_fr(p.T, bp + 24);
      vr_store_cached(tb, p);
#pragma unroll 1
      for (int e = 1; e < 8; e++) {
        if (e == 1)
          vr_double(p, true);
        else
          vr_add_cached(p, tb, false);
        vr_store_cached(tb + e * 32, p);
      }
    }
    ext_identity(p);
#pragma unroll 1
    for (int win = 63; win >= 0; win--) {
#pragma unroll 1
      for (int d = 0; d < 4; d++) vr_double(p, d == 3);
#pragma unroll 1
      for (int base = 0; base < nb; base++) {
        const u32 mw = base ? vb_pick8(mag1, win >> 3) : vb_pick8(mag0, win >> 3);
        const u32 sw = base ? ((win >> 5) ? sgn1[1] : sgn1[0]) : ((win >> 5) ? sgn0[1] : sgn0[0]);
        const u32 mg = (mw >> ((win & 7) * 4)) & 15u;
        const bool neg = ((sw >> (win & 31)) & 1u) != 0;
        if (mg != 0) vr_add_cached(p, tab + ((size_t)base * 8 + (mg - 1)) * 32, neg);
      }
    }
  }
  store_fr(o, p.X);
  store_fr(o + 8, p.Y);
  store_fr(o + 16, p.Z);
  store_fr(o + 24, p.T);
}

}  // namespace gcp
