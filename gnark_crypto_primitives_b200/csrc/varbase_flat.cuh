// Variable-base window kernel, flat form (the default): the same algorithm, register file and multiplier bodies as
// varbase_window_kernel (varbase.cuh: signed 4-bit windows, four doublings and one cached addition per base and window,
// per-item table [1..8]P in global scratch, Straus for two bases) with the interpreter's overheads taken out.
//
// What the SASS of the step-structured interpreter showed (static count per window, one base: 10 484 instructions of
// which 4 288 IMAD.WIDE = 40.9 %, ncu: 39.5 % executed): the FMA-heavy pipe needs 2 issue slots per IMAD.WIDE, so with
// fewer than half of the instructions wide the kernel is ISSUE-bound (9.3 k pipe cycles against 10.5 k issue slots per
// window; fmaheavy 76 %).  Of the 6 200 other instructions ~1 400 were the sequencer (a ~125-instruction step prologue -
// build / main phase, digit pick out of 20 digit registers, table pointer - five times per window, and ~30 per micro-op
// for fetch, field extraction and a 14-way switch), ~850 the doubling's seven separately normalised add-type operations.
// Here:
//   * ONE op stream: [table build, base 1][table build, base 0][P = O][window program] with the window program repeated
//     64 times; per op: two LDC.128 (operand byte offsets stored as whole words: no field extraction), an if-chain that
//     reaches the three multiplier bodies in 1-3 compares, a 4-instruction program counter.
//   * control flow is uniform over the block: a zero digit adds the cached identity (1, 1, 0, 2) instead of skipping, an
//     item with a status computes on whatever its slots hold and is overwritten with the identity at the end.
//   * the recoded digits live in shared memory (10 words per base and thread, word-interleaved), not in 20 registers.
//   * the add-type work is the tail of the op that produces its last operand (operands still in registers): the doubling
//     tail after the second squaring pair, the addition tail after the second product pair.
//   * doubling with all four middle terms negated (E' = A + B - S, F' = 2 Z^2 + A - B, G' = A - B, H' = A + B: every
//     output is a product of two of them, so the signs cancel): five normalised add-type operations instead of seven;
//     a negative digit swaps the destinations of D - C and D + C instead of negating C.
// Bounds are those of varbase.cuh: every add-type result is normalised to [0, 2r), every product of two such values is
// below 1.76 r, squarings see values below 2r < 2^255.
#pragma once
#include "varbase.cuh"

namespace gcp {

constexpr int VF_DIG_WORDS = 10;  // per base and thread: 8 words of digit magnitudes, 2 words of digit signs
constexpr size_t VF_RF_BYTES = (size_t)VB_SLOTS * 2 * VB_THREADS * 16;
constexpr size_t VF_SMEM_BYTES = VF_RF_BYTES + (size_t)2 * VF_DIG_WORDS * VB_THREADS * 4;

enum : u32 { VF_MUL2 = 0, VF_SQR2 = 1, VF_MUL = 2, VF_LDQB = 3, VF_STQ = 4, VF_BFLY = 5, VF_ADD = 6, VF_LOADBASE = 7, VF_SETID = 8 };
constexpr u32 VF_OPMASK = 0xffu;
constexpr u32 VF_MID = 0x100u;    // MUL2: then the addition's middle terms; SQR2: then the doubling's middle terms
constexpr u32 VF_SUM = 0x200u;    // SQR2: store f2 + f5 to f3 first (the doubling's X + Y)
constexpr u32 VF_FIXED = 0x400u;  // LDQB: table entry [1]P, not the window's digit (table build)
constexpr u32 VF_BASE1 = 0x800u;  // LDQB / STQ / LOADBASE: the second base

// 32 bytes per op: byte offsets of the six operand slots inside the register file as whole words
struct alignas(16) VfOp {
  u32 op, o1, o2, o3, o4, o5, o6, aux;
};

constexpr int VF_DBL_LEN = 4, VF_ADD_LEN = 5, VF_CACHE_LEN = 4;
constexpr int VF_BUILD_LEN = 1 + VF_CACHE_LEN + VF_DBL_LEN + VF_CACHE_LEN + 6 * (VF_ADD_LEN + VF_CACHE_LEN);  // 67
constexpr int VF_PC_BUILD1 = 0, VF_PC_BUILD0 = VF_BUILD_LEN, VF_PC_SETID = 2 * VF_BUILD_LEN;
constexpr int VF_PC_MAIN1 = VF_PC_SETID + 1, VF_MAIN1_LEN = 4 * VF_DBL_LEN + VF_ADD_LEN;
constexpr int VF_PC_MAIN2 = VF_PC_MAIN1 + VF_MAIN1_LEN, VF_MAIN2_LEN = 4 * VF_DBL_LEN + 2 * VF_ADD_LEN;
constexpr int VF_PROG_LEN = VF_PC_MAIN2 + VF_MAIN2_LEN;

struct alignas(16) VfProgram {
  VfOp ops[VF_PROG_LEN];
};

constexpr VfOp vf_op(u32 op, int f1 = 0, int f2 = 0, int f3 = 0, int f4 = 0, int f5 = 0, int f6 = 0, u32 aux = 0) {
  return VfOp{op, VB_OFF(f1), VB_OFF(f2), VB_OFF(f3), VB_OFF(f4), VB_OFF(f5), VB_OFF(f6), aux};
}
// slots: 0 X, 1 Y, 2 Z, 3 T, 4..9 temporaries, 10 the constant 2d
// P = 2P (dbl-2008-hwcd, a = -1) with negated middle terms: A = X^2 (4), B = Y^2 (5), X + Y (3: T is dead),
// Z^2, (X+Y)^2 -> H' = A + B (4), E' = H' - (X+Y)^2 (7), G' = A - B (8), F' = G' + 2 Z^2 (6);
// X = E' F', Y = G' H', Z = F' G', T = E' H'
constexpr int vf_emit_dbl(VfOp* p, int n, bool with_t) {
  p[n++] = vf_op(VF_SQR2 | VF_SUM, 4, 0, 3, 5, 1);
  p[n++] = vf_op(VF_SQR2 | VF_MID, 6, 2, 0, 7, 3);
  p[n++] = vf_op(VF_MUL2, 0, 7, 6, 1, 8, 4);
  p[n++] = with_t ? vf_op(VF_MUL2, 2, 6, 8, 3, 7, 4) : vf_op(VF_MUL, 2, 6, 8);
  return n;
}
// P += Q (add-2008-hwcd-3, Q = (Y-X, Y+X, 2dT, 2Z) cached): LDQB: Y - X (4), Y + X (5), Q (6..9); A = 4 * 6, B = 5 * 7;
// C = T * 8, D = Z * 9 -> E = B - A (8), H = B + A (4), F = D - C (9), G = D + C (6) (F and G exchanged for -Q);
// X = E F, Y = G H, T = E H, Z = F G
constexpr int vf_emit_add(VfOp* p, int n, u32 ldq_flags) {
  p[n++] = vf_op(VF_LDQB | ldq_flags);
  p[n++] = vf_op(VF_MUL2, 4, 4, 6, 5, 5, 7);
  p[n++] = vf_op(VF_MUL2 | VF_MID, 6, 3, 8, 7, 2, 9);
  p[n++] = vf_op(VF_MUL2, 0, 8, 9, 1, 6, 4);
  p[n++] = vf_op(VF_MUL2, 3, 8, 4, 2, 9, 6);
  return n;
}
// table entry <- cached(P) = (Y-X, Y+X, 2d T, 2Z)
constexpr int vf_emit_cache(VfOp* p, int n, u32 base_flag, u32 entry) {
  p[n++] = vf_op(VF_BFLY, 4, 1, 0, 5);
  p[n++] = vf_op(VF_MUL, 6, 3, VB_SLOT_2D);
  p[n++] = vf_op(VF_ADD, 7, 2, 2);
  p[n++] = vf_op(VF_STQ | base_flag, 0, 0, 0, 0, 0, 0, entry);
  return n;
}
constexpr int vf_emit_build(VfOp* p, int n, u32 base_flag) {
  p[n++] = vf_op(VF_LOADBASE | base_flag);
  n = vf_emit_cache(p, n, base_flag, 0);
  n = vf_emit_dbl(p, n, true);
  n = vf_emit_cache(p, n, base_flag, 1);
  for (u32 e = 2; e < 8; e++) {
    n = vf_emit_add(p, n, VF_FIXED | base_flag);
    n = vf_emit_cache(p, n, base_flag, e);
  }
  return n;
}
constexpr VfProgram vf_build_program() {
  VfProgram g{};
  int n = 0;
  n = vf_emit_build(g.ops, n, VF_BASE1);
  n = vf_emit_build(g.ops, n, 0);
  g.ops[n++] = vf_op(VF_SETID);
  for (int nb = 1; nb <= 2; nb++) {
    for (int d = 0; d < 4; d++) n = vf_emit_dbl(g.ops, n, d == 3);
    for (int b = 0; b < nb; b++) n = vf_emit_add(g.ops, n, b ? VF_BASE1 : 0);
  }
  return g;
}
static_assert(VF_PROG_LEN == 2 * VF_BUILD_LEN + 1 + VF_MAIN1_LEN + VF_MAIN2_LEN, "program layout");
__device__ __constant__ VfProgram c_vf_prog = vf_build_program();

// cached form of the identity (0 : 1 : 1 : 0): (Y - X, Y + X, 2dT, 2Z) = (1, 1, 0, 2) in Montgomery form
__device__ const u32 c_vf_identity_entry[32] = {
    0x4ffffffbu, 0xac96341cu, 0x9f60cd29u, 0x36fc7695u, 0x7879462eu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u,
    0x4ffffffbu, 0xac96341cu, 0x9f60cd29u, 0x36fc7695u, 0x7879462eu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u,
    0, 0, 0, 0, 0, 0, 0, 0,
    0x9ffffff6u, 0x592c6838u, 0x3ec19a53u, 0x6df8ed2bu, 0xf0f28c5cu, 0xccdd46deu, 0x340fbe5eu, 0x1c14ef83u};

__device__ __forceinline__ u32 vf_lds(u32 sa) {
  u32 v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(sa) : "memory");
  return v;
}
__device__ __forceinline__ void vf_sts(u32 sa, u32 v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(sa), "r"(v) : "memory"); }

template <int MIN_BLOCKS>
__global__ void __launch_bounds__(VB_THREADS, MIN_BLOCKS) varbase_flat_kernel(VarbaseArgs a) {
  extern __shared__ uint4 vf_smem[];  // VB_SLOTS x 2 x VB_THREADS uint4, then 2 x VF_DIG_WORDS x VB_THREADS words
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= a.n) return;
  const u32 rf_sa = (u32)__cvta_generic_to_shared(vf_smem + threadIdx.x);
  const u32 dig_sa = (u32)__cvta_generic_to_shared(vf_smem) + (u32)VF_RF_BYTES + threadIdx.x * 4u;
  const bool live = !a.status || a.status[idx] == GCP_STATUS_OK;
  const int nb = a.n_bases;
#pragma unroll 1
  for (int b = 0; b < nb; b++) {
    u32 k[8], mag[8], sgn[2];
    load_fr(k, a.scalars[b] + idx * 8);
    vb_recode(mag, sgn, k);
#pragma unroll
    for (int w = 0; w < 8; w++) vf_sts(dig_sa + (u32)(b * VF_DIG_WORDS + w) * (VB_THREADS * 4u), mag[w]);
    vf_sts(dig_sa + (u32)(b * VF_DIG_WORDS + 8) * (VB_THREADS * 4u), sgn[0]);
    vf_sts(dig_sa + (u32)(b * VF_DIG_WORDS + 9) * (VB_THREADS * 4u), sgn[1]);
  }
  {
    const u32 d2[8] = GCP_ED_2D_MONT;
    vb_st(rf_sa + VB_OFF(VB_SLOT_2D), d2);
  }
  const u32* const bases = a.bases + idx * (size_t)nb * 32;
  u32* const tab = a.table + idx * (size_t)nb * VB_TABLE_WORDS;

  const uint4* const prog = reinterpret_cast<const uint4*>(c_vf_prog.ops);
  const int main_start = nb == 2 ? VF_PC_MAIN2 : VF_PC_MAIN1;
  const int main_end = main_start + (nb == 2 ? VF_MAIN2_LEN : VF_MAIN1_LEN);
  int pc = nb == 2 ? VF_PC_BUILD1 : VF_PC_BUILD0, pc_end = VF_PC_MAIN1;
  int win = 64;  // 64: table build; 63..0: the windows, most significant first
  bool neg = false;
  uint4 lo = prog[2 * pc], hi = prog[2 * pc + 1];
#pragma unroll 1
  for (;;) {
    const uint4 c0 = lo, c1 = hi;
    int npc = pc + 1, nwin = win;
    if (npc == pc_end) {
      nwin = win - 1;
      npc = main_start;
      pc_end = main_end;
    }
    lo = prog[2 * npc];  // fetched one op ahead
    hi = prog[2 * npc + 1];
    const u32 op = c0.x & VF_OPMASK;
    if (op == VF_MUL2) {
      u32 x1[8], y1[8], x2[8], y2[8], r1[8], r2[8];
      vb_ld(x1, rf_sa + c0.z);
      vb_ld(y1, rf_sa + c0.w);
      vb_ld(x2, rf_sa + c1.y);
      vb_ld(y2, rf_sa + c1.z);
      fr_mul2(r1, x1, y1, r2, x2, y2);
      if (c0.x & VF_MID) {  // r1 = C = T * 2dT_Q, r2 = D = Z * 2Z_Q
        u32 t[8];
        vb_ld(x1, rf_sa + VB_OFF(4));  // A
        vb_ld(y1, rf_sa + VB_OFF(5));  // B
        fr_sub(t, y1, x1);
        vb_st(rf_sa + VB_OFF(8), t);  // E = B - A
        fr_add(t, y1, x1);
        vb_st(rf_sa + VB_OFF(4), t);  // H = B + A
        fr_sub(t, r2, r1);
        vb_st(rf_sa + (neg ? VB_OFF(6) : VB_OFF(9)), t);  // F = D - C (G for -Q)
        fr_add(t, r2, r1);
        vb_st(rf_sa + (neg ? VB_OFF(9) : VB_OFF(6)), t);  // G = D + C (F for -Q)
      } else {
        vb_st(rf_sa + c0.y, r1);
        vb_st(rf_sa + c1.x, r2);
      }
    } else if (op == VF_SQR2) {
      u32 x1[8], x2[8], r1[8], r2[8];
      vb_ld(x1, rf_sa + c0.z);
      vb_ld(x2, rf_sa + c1.y);
      if (c0.x & VF_SUM) {
        fr_add(r1, x1, x2);
        vb_st(rf_sa + c0.w, r1);
      }
      fr_sqr2(r1, x1, r2, x2);
      if (c0.x & VF_MID) {  // r1 = Z^2, r2 = (X + Y)^2
        u32 t[8], g[8];
        vb_ld(x1, rf_sa + VB_OFF(4));  // A
        vb_ld(x2, rf_sa + VB_OFF(5));  // B
        fr_add(t, x1, x2);
        vb_st(rf_sa + VB_OFF(4), t);  // H' = A + B
        fr_sub(g, t, r2);
        vb_st(rf_sa + VB_OFF(7), g);  // E' = A + B - (X + Y)^2
        fr_sub(g, x1, x2);
        vb_st(rf_sa + VB_OFF(8), g);  // G' = A - B
        fr_add(t, r1, r1);
        fr_add(x1, g, t);
        vb_st(rf_sa + VB_OFF(6), x1);  // F' = G' + 2 Z^2
      } else {
        vb_st(rf_sa + c0.y, r1);
        vb_st(rf_sa + c1.x, r2);
      }
    } else if (op == VF_MUL) {
      u32 x[8], y[8], r[8];
      vb_ld(x, rf_sa + c0.z);
      vb_ld(y, rf_sa + c0.w);
      fr_mul(r, x, y);
      vb_st(rf_sa + c0.y, r);
    } else if (op == VF_LDQB) {
      const u32 b = (c0.x & VF_BASE1) ? 1u : 0u;
      u32 mg = 1;
      neg = false;
      if (!(c0.x & VF_FIXED)) {
        const u32 mw = vf_lds(dig_sa + (b * VF_DIG_WORDS + ((u32)win >> 3)) * (VB_THREADS * 4u));
        const u32 sw = vf_lds(dig_sa + (b * VF_DIG_WORDS + 8u + ((u32)win >> 5)) * (VB_THREADS * 4u));
        mg = (mw >> (((u32)win & 7u) * 4u)) & 15u;
        neg = ((sw >> ((u32)win & 31u)) & 1u) != 0;
      }
      const u32* q = mg ? tab + (b * 8u + mg - 1u) * 32u : c_vf_identity_entry;
      u32 x[8], y[8], t[8];
      vb_ld(x, rf_sa + VB_OFF(0));
      vb_ld(y, rf_sa + VB_OFF(1));
      fr_sub(t, y, x);
      vb_st(rf_sa + VB_OFF(4), t);
      fr_add(t, y, x);
      vb_st(rf_sa + VB_OFF(5), t);
      {
        u32 v[8];
        load_fr_plain(v, q);
        vb_st(rf_sa + (neg ? VB_OFF(7) : VB_OFF(6)), v);
        load_fr_plain(v, q + 8);
        vb_st(rf_sa + (neg ? VB_OFF(6) : VB_OFF(7)), v);
        load_fr_plain(v, q + 16);
        vb_st(rf_sa + VB_OFF(8), v);
        load_fr_plain(v, q + 24);
        vb_st(rf_sa + VB_OFF(9), v);
      }
    } else if (op == VF_STQ) {
      u32* q = tab + (((c0.x & VF_BASE1) ? 8u : 0u) + c1.w) * 32u;
#pragma unroll
      for (int c = 0; c < 4; c++) {
        u32 v[8];
        vb_ld(v, rf_sa + VB_OFF(4 + c));
        store_fr(q + c * 8, v);
      }
    } else if (op == VF_BFLY) {
      u32 x[8], y[8], r[8];
      vb_ld(x, rf_sa + c0.z);
      vb_ld(y, rf_sa + c0.w);
      fr_sub(r, x, y);
      vb_st(rf_sa + c0.y, r);
      fr_add(r, x, y);
      vb_st(rf_sa + c1.x, r);
    } else if (op == VF_ADD) {
      u32 x[8], y[8], r[8];
      vb_ld(x, rf_sa + c0.z);
      vb_ld(y, rf_sa + c0.w);
      fr_add(r, x, y);
      vb_st(rf_sa + c0.y, r);
    } else if (op == VF_LOADBASE) {
      const u32* bp = bases + ((c0.x & VF_BASE1) ? 32 : 0);
#pragma unroll
      for (int c = 0; c < 4; c++) {
        u32 v[8];
        load_fr(v, bp + c * 8);
        vb_st(rf_sa + c * VB_OFF(1), v);
      }
    } else {  // VF_SETID: P = (0 : 1 : 1 : 0)
      u32 zero[8], one[8];
      fr_set_zero(zero);
      fr_set_one(one);
      vb_st(rf_sa + VB_OFF(0), zero);
      vb_st(rf_sa + VB_OFF(1), one);
      vb_st(rf_sa + VB_OFF(2), one);
      vb_st(rf_sa + VB_OFF(3), zero);
    }
    if (nwin < 0) break;
    pc = npc;
    win = nwin;
  }
  u32* o = a.out + idx * 32;
#pragma unroll
  for (int c = 0; c < 4; c++) {
    u32 v[8];
    if (live) {
      vb_ld(v, rf_sa + c * VB_OFF(1));
    } else if (c == 1 || c == 2) {
      fr_set_one(v);
    } else {
      fr_set_zero(v);
    }
    store_fr(o + c * 8, v);
  }
}

}  // namespace gcp
