// Variable-base scalar multiplication, curve.ScalarMul(P, k) of the reference (gnark std/algebra/native/twistededwards,
// un-vendored; call sites /root/reference/elgamal/encrypt.go:55, elgamal/ciphertext.go:58,147-160,
// ecc/bn254/eddsa/verifier.go:71-80): [k]P for the integer k in [0, r), and the gadgets built on it:
//   (*Ciphertext).Encrypt with a key per item   /root/reference/elgamal/encrypt.go:42-64
//   (*Ciphertext).AssertDecrypt                 /root/reference/elgamal/ciphertext.go:50-67
//   DecryptionProof.Verify                      /root/reference/elgamal/ciphertext.go:124-168 (hashPointsToScalar :173-184)
//   EdDSA-Poseidon Verifier.IsValid             /root/reference/ecc/bn254/eddsa/verifier.go:55-88
//
// Shape of the work.  Each gadget is split into a PRE pass (validation, Montgomery conversion, the fixed-base parts,
// the inputs of the Fiat-Shamir / EdDSA hash), the Poseidon batch kernel where a challenge is needed, the WINDOW kernel
// below (all of the variable-base work: ~85 % of the multiplies), and a POST pass (a few additions and the projective
// comparison).  The round-1 kernels did all of this per thread in one kernel and were bound by instruction fetch: a
// doubling and an addition inlined once each are ~55 KB of SASS against a 32 KB instruction cache (ncu: icc hit rate
// 74-78 %, 1.6-2.2 no-instruction stalls per issue), and out-of-line multipliers passed their operands through local
// memory.
//
// The window kernel is a field-operation interpreter.  A thread's working set (the accumulator X, Y, Z, T and six
// temporaries: ten 32-byte slots) lives in a shared-memory register file laid out [slot][half][thread], so a slot is two
// conflict-free LDS.128 / STS.128; the point formulas are micro-programs in constant memory (op | dst | a | b), uniform
// over the block, and the interpreter holds exactly ONE inlined multiplier, ONE squaring and one add / sub / neg:
// ~10 KB of SASS for the whole kernel, no register-file rotation, ~90 registers.  Signed 4-bit windows (digits in
// [-7, 8], 64 of them, MSB first): per window four doublings (dbl-2008-hwcd, the first three without T) and one addition
// per base (add-2008-hwcd-3 against a "cached" table entry (Y-X, Y+X, 2dT, 2Z), 8 multiplies).  The per-task table
// [1..8]P is built by the same micro-programs and kept in GLOBAL scratch as [task][base][entry][32 words]: a lookup is
// one whole 128-byte line per lane (the round-1 per-thread local array interleaves threads word by word, so 32 lanes
// with 32 different digits touched 1024 lines for 4 KB of data).  Two bases share the doublings (Straus):
// DecryptionProof.Verify's  [z]C1 - [e]D  is one pass.
#pragma once
#include "edwards.cuh"
#include "elgamal.cuh"
#include "poseidon.cuh"
#include "smt.cuh"

namespace gcp {

constexpr int VB_THREADS = 128;
constexpr int VB_SLOTS = 11;
constexpr int VB_TABLE_WORDS = 8 * 32;  // per (task, base): entries [1..8]P x (ymx, ypx, t2d, z2)

// micro-ops on the shared-memory register file: op | f1 << 4 | f2 << 8 | f3 << 12 | f4 << 16 | f5 << 20 | f6 << 24 (slot numbers)
//   MUL  f1 = f2 * f3                      MUL2  f1 = f2 * f3 and f4 = f5 * f6 (two reduction chains in flight)
//   SQR2 f1 = f2^2 and f4 = f5^2           BFLY  f1 = f2 - f3 and f4 = f2 + f3
//   ADD  f1 = f2 + f3;  SUB f1 = f2 - f3;  SUB3 f1 = f2 - f3 - f4;  NEG f1 = -f2;  CNEG f1 = -f2 where the digit is negative
//   LDQ  slots f1..f1+3 <- the thread's table entry (Y-X and Y+X exchanged where the digit is negative: -(x, y) = (-x, y))
//   STQ  the thread's table entry <- slots f2..f2+3
// Every op reads all of its operands before it writes, so destinations may alias sources.
// Fused ops with fixed slots (round 2: 29 of the 49 ops of a window were add-type ops, each a round trip of its operands
// through shared memory; now 26 ops per window):
//   SQR2 with f3 != 0 also stores f2 + f5 to f3 before squaring (the doubling's X + Y)
//   DBLMID  A 4, B 5, C' 6, E' 7  ->  E 7 = E' - A - B, F 6 = (B - A) - 2 C', G 8 = B - A, H 4 = -(A + B)
//   ADDMID  A 4, B 5, C 6, D 7 (C negated where the digit is negative)  ->  E 8 = B - A, H 4 = B + A, F 9 = D - C, G 6 = D + C
//   LDQB    LDQ into 6..9 and (Y - X, Y + X) into 4, 5
enum : u32 { VB_MUL = 0, VB_MUL2 = 1, VB_SQR2 = 2, VB_BFLY = 3, VB_ADD = 4, VB_SUB = 5, VB_SUB3 = 6, VB_NEG = 7, VB_CNEG = 8,
             VB_LDQ = 9, VB_STQ = 10, VB_DBLMID = 11, VB_ADDMID = 12, VB_LDQB = 13 };
// stored pre-decoded, 16 bytes per op (one LDC.128): {op, off1 | off2 << 16, off3 | off4 << 16, off5 | off6 << 16} with
// off = slot * 4096, the byte offset of the slot inside the register file (2 halves x 128 threads x 16 bytes)
#define VB_OFF(slot) ((u32)(slot) * (2u * VB_THREADS * 16u))
#define VB_OP2(op, f1, f2, f3, f4, f5, f6)                                                                    \
  {(u32)(op), VB_OFF(f1) | (VB_OFF(f2) << 16), VB_OFF(f3) | (VB_OFF(f4) << 16), VB_OFF(f5) | (VB_OFF(f6) << 16)}
#define VB_OP(op, f1, f2, f3) VB_OP2(op, f1, f2, f3, 0, 0, 0)

// slots: 0 X, 1 Y, 2 Z, 3 T, 4..9 temporaries, 10 the constant 2d
constexpr int VB_SLOT_2D = 10;
constexpr int VB_PROG_DBL = 0, VB_PROG_DBL_T = 5, VB_PROG_DBL_LEN = 5;  // without / with T (T is dead before another doubling)
constexpr int VB_PROG_ADD = 10, VB_PROG_ADD_LEN = 6;
constexpr int VB_PROG_CACHE = 16, VB_PROG_CACHE_LEN = 4;
// P = 2P (dbl-2008-hwcd, a = -1): A = X^2, B = Y^2, C = 2 Z^2, E = (X+Y)^2 - A - B, G = B - A, F = G - C, H = -(A+B);
// X = E F, Y = G H, Z = F G, T = E H.   slots: A 4, B 5, C' 6, E' 7, then E 7, F 6, G 8, H 4
#define VB_DBL_HEAD \
  VB_OP2(VB_SQR2, 4, 0, 3, 5, 1, 0), VB_OP2(VB_SQR2, 6, 2, 0, 7, 3, 0), VB_OP(VB_DBLMID, 0, 0, 0), VB_OP2(VB_MUL2, 0, 7, 6, 1, 8, 4)
__device__ __constant__ uint4 c_vb_prog[21] = {
    VB_DBL_HEAD, VB_OP(VB_MUL, 2, 6, 8),
    VB_DBL_HEAD, VB_OP2(VB_MUL2, 2, 6, 8, 3, 7, 4),
    // P += Q (add-2008-hwcd-3, Q cached in slots 6..9): A = (Y-X) q0, B = (Y+X) q1, C = T q2, D = Z q3,
    // E = B - A, H = B + A, F = D - C, G = D + C;  X = E F, Y = G H, T = E H, Z = F G
    VB_OP(VB_LDQB, 0, 0, 0), VB_OP2(VB_MUL2, 4, 4, 6, 5, 5, 7), VB_OP2(VB_MUL2, 6, 3, 8, 7, 2, 9), VB_OP(VB_ADDMID, 0, 0, 0),
    VB_OP2(VB_MUL2, 0, 8, 9, 1, 6, 4), VB_OP2(VB_MUL2, 3, 8, 4, 2, 9, 6),
    // table entry <- cached(P) = (Y-X, Y+X, 2d T, 2Z)
    VB_OP2(VB_BFLY, 4, 1, 0, 5, 0, 0), VB_OP(VB_MUL, 6, 3, VB_SLOT_2D), VB_OP(VB_ADD, 7, 2, 2), VB_OP(VB_STQ, 0, 4, 0),
    VB_OP(VB_NEG, 0, 0, 0) /* never executed: the interpreter fetches one op ahead */};

struct VarbaseArgs {
  const u32* bases;       // n x n_bases x 32 words: extended (X, Y, Z, T), lazy Montgomery, on the curve
  const u32* scalars[2];  // per base: n x 8 words, integers < 2^254
  int n_bases;            // 1, or 2 (Straus: the doublings are shared)
  size_t n;
  const u8* status;       // n, or nullptr: items with status != 0 are skipped (out = identity)
  u32* table;             // n x n_bases x VB_TABLE_WORDS scratch
  u32* out;               // n x 32 words (X, Y, Z, T), lazy Montgomery
};

// a slot is two 16-byte pieces VB_THREADS * 16 bytes apart; `sa` is a 32-bit shared-window address (slot base + this thread)
__device__ __forceinline__ void vb_ld(u32 (&r)[8], u32 sa) {
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(sa) : "memory");
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4+2048];" : "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(sa) : "memory");
}
__device__ __forceinline__ void vb_st(u32 sa, const u32 (&r)[8]) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(sa), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]) : "memory");
  asm volatile("st.shared.v4.u32 [%0+2048], {%1, %2, %3, %4};" ::"r"(sa), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
static_assert(VB_THREADS * 16 == 2048, "vb_ld / vb_st hard-code the distance between the halves of a slot");

// scalar -> 64 signed nibbles d_i in [-7, 8], sum d_i 16^i = k: |d_i| packed 4 bits each in mag[8], signs in sgn[2]
__device__ __forceinline__ void vb_recode(u32 (&mag)[8], u32 (&sgn)[2], const u32 (&k)[8]) {
  sgn[0] = sgn[1] = 0;
  u32 carry = 0;
#pragma unroll
  for (int w = 0; w < 8; w++) {
    u32 mw = 0, sw = 0;
#pragma unroll
    for (int q = 0; q < 8; q++) {
      const u32 v = ((k[w] >> (q * 4)) & 15u) + carry;
      carry = v > 8u ? 1u : 0u;
      mw |= (carry ? 16u - v : v) << (q * 4);
      sw |= carry << q;
    }
    mag[w] = mw;
    sgn[w >> 2] |= sw << ((w & 3) * 8);
  }
}

__device__ __forceinline__ u32 vb_pick8(const u32 (&a)[8], int i) {
  u32 v = a[0];
#pragma unroll
  for (int w = 1; w < 8; w++) v = (i == w) ? a[w] : v;
  return v;
}

template <int MIN_BLOCKS>
__global__ void __launch_bounds__(VB_THREADS, MIN_BLOCKS) varbase_window_kernel(VarbaseArgs a) {
  extern __shared__ uint4 vb_smem[];  // VB_SLOTS x 2 x VB_THREADS
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= a.n) return;
  const u32 rf_sa = (u32)__cvta_generic_to_shared(vb_smem + threadIdx.x);
  const bool live = !a.status || a.status[idx] == GCP_STATUS_OK;
  const int nb = a.n_bases;
  u32 mag0[8], sgn0[2], mag1[8], sgn1[2];
  {
    u32 k[8];
    load_fr(k, a.scalars[0] + idx * 8);
    vb_recode(mag0, sgn0, k);
    if (nb > 1) load_fr(k, a.scalars[1] + idx * 8);
    vb_recode(mag1, sgn1, k);
  }
  u32* const tab = a.table + idx * (size_t)nb * VB_TABLE_WORDS;
  {
    const u32 d2[8] = GCP_ED_2D_MONT;
    vb_st(rf_sa + VB_OFF(VB_SLOT_2D), d2);
  }

  // step sequence, uniform over the block: per base 15 table steps (cache e0, dbl, cache e1, then 6 x (add e0, cache)),
  // then 64 windows of (dbl, dbl, dbl, dbl+T, add per base)
  constexpr int BUILD = 15;
  const int n_build = nb * BUILD, per_window = 4 + nb, n_steps = n_build + 64 * per_window;
  int win = 63, ws = 0;  // window and step inside it, advanced once per main-phase step
#pragma unroll 1
  for (int seq = 0; seq < n_steps; seq++) {
    int pc, pc_end, base = 0, entry = 0;
    bool neg = false, active = live;
    if (seq < n_build) {
      base = seq >= BUILD ? 1 : 0;
      const int s = seq - base * BUILD;
      if (s == 0 && live) {  // P = base point
        const u32* bp = a.bases + (idx * (size_t)nb + base) * 32;
#pragma unroll
        for (int c = 0; c < 4; c++) {
          u32 v[8];
          load_fr(v, bp + c * 8);
          vb_st(rf_sa + c * VB_OFF(1), v);
        }
      }
      if (s == 1) {
        pc = VB_PROG_DBL_T;
        pc_end = pc + VB_PROG_DBL_LEN;
      } else if (s & 1) {
        pc = VB_PROG_ADD;
        pc_end = pc + VB_PROG_ADD_LEN;
      } else {
        pc = VB_PROG_CACHE;
        pc_end = pc + VB_PROG_CACHE_LEN;
        entry = s >> 1;
      }
    } else {
      const int s = ws;
      if (seq == n_build) {  // P = identity (0 : 1 : 1 : 0)
        u32 zero[8], one[8];
        fr_set_zero(zero);
        fr_set_one(one);
        vb_st(rf_sa + VB_OFF(0), zero);
        vb_st(rf_sa + VB_OFF(1), one);
        vb_st(rf_sa + VB_OFF(2), one);
        vb_st(rf_sa + VB_OFF(3), zero);
      }
      if (s < 4) {
        pc = s == 3 ? VB_PROG_DBL_T : VB_PROG_DBL;
        pc_end = pc + VB_PROG_DBL_LEN;
      } else {
        base = s - 4;
        const u32 mw = base ? vb_pick8(mag1, win >> 3) : vb_pick8(mag0, win >> 3);
        const u32 sw = base ? ((win >> 5) ? sgn1[1] : sgn1[0]) : ((win >> 5) ? sgn0[1] : sgn0[0]);
        const u32 mg = (mw >> ((win & 7) * 4)) & 15u;
        neg = ((sw >> (win & 31)) & 1u) != 0;
        active = live && mg != 0;
        entry = (int)mg - 1;
        pc = VB_PROG_ADD;
        pc_end = pc + VB_PROG_ADD_LEN;
      }
    }
    u32* const q = tab + ((size_t)base * 8 + (entry < 0 ? 0 : entry)) * 32;
    if (seq >= n_build && ++ws == per_window) {
      ws = 0;
      win--;
    }
    if (!active) continue;
    uint4 ins = c_vb_prog[pc];
#pragma unroll 1
    for (; pc < pc_end; pc++) {
      const uint4 cur = ins;
      ins = c_vb_prog[pc + 1];  // fetched one op ahead: the constant-cache latency hides behind this op's arithmetic
      const u32 s1 = rf_sa + (cur.y & 0xffffu), s2 = rf_sa + (cur.y >> 16), s3 = rf_sa + (cur.z & 0xffffu);
      const u32 s4 = rf_sa + (cur.z >> 16);
      // every case loads its operands, computes and stores on its own: no register merging between the bodies
      switch (cur.x) {
        case VB_MUL2: {
          const u32 s5 = rf_sa + (cur.w & 0xffffu), s6 = rf_sa + (cur.w >> 16);
          u32 x1[8], y1[8], x2[8], y2[8], r1[8], r2[8];
          vb_ld(x1, s2);
          vb_ld(y1, s3);
          vb_ld(x2, s5);
          vb_ld(y2, s6);
          fr_mul2(r1, x1, y1, r2, x2, y2);
          vb_st(s1, r1);
          vb_st(s4, r2);
          break;
        }
        case VB_SQR2: {
          const u32 s5 = rf_sa + (cur.w & 0xffffu);
          u32 x1[8], x2[8], r1[8], r2[8];
          vb_ld(x1, s2);
          vb_ld(x2, s5);
          if (cur.z & 0xffffu) {  // also f3 = f2 + f5 (the doubling's X + Y, squared by the next op)
            fr_add(r1, x1, x2);
            vb_st(s3, r1);
          }
          fr_sqr2(r1, x1, r2, x2);
          vb_st(s1, r1);
          vb_st(s4, r2);
          break;
        }
        case VB_MUL: {
          u32 x[8], y[8], r[8];
          vb_ld(x, s2);
          vb_ld(y, s3);
          fr_mul(r, x, y);
          vb_st(s1, r);
          break;
        }
        case VB_BFLY: {
          u32 x[8], y[8], r[8];
          vb_ld(x, s2);
          vb_ld(y, s3);
          fr_sub(r, x, y);
          vb_st(s1, r);
          fr_add(r, x, y);
          vb_st(s4, r);
          break;
        }
        case VB_ADD: {
          u32 x[8], y[8], r[8];
          vb_ld(x, s2);
          vb_ld(y, s3);
          fr_add(r, x, y);
          vb_st(s1, r);
          break;
        }
        case VB_SUB3: {
          u32 x[8], y[8], r[8];
          vb_ld(x, s2);
          vb_ld(y, s3);
          fr_sub(r, x, y);
          vb_ld(y, s4);
          fr_sub(x, r, y);
          vb_st(s1, x);
          break;
        }
        case VB_SUB: {
          u32 x[8], y[8], r[8];
          vb_ld(x, s2);
          vb_ld(y, s3);
          fr_sub(r, x, y);
          vb_st(s1, r);
          break;
        }
        case VB_CNEG:
          if (!neg) break;
          // fall through
        case VB_NEG: {
          u32 x[8], r[8];
          vb_ld(x, s2);
          fr_neg(r, x);
          vb_st(s1, r);
          break;
        }
        case VB_DBLMID: {
          u32 a[8], b[8], c[8], e[8], t[8];
          vb_ld(a, rf_sa + VB_OFF(4));
          vb_ld(b, rf_sa + VB_OFF(5));
          vb_ld(e, rf_sa + VB_OFF(7));
          fr_sub(t, e, a);
          fr_sub(e, t, b);
          vb_st(rf_sa + VB_OFF(7), e);   // E = (X+Y)^2 - A - B
          fr_sub(e, b, a);
          vb_st(rf_sa + VB_OFF(8), e);   // G = B - A
          vb_ld(c, rf_sa + VB_OFF(6));
          fr_add(t, c, c);
          fr_sub(c, e, t);
          vb_st(rf_sa + VB_OFF(6), c);   // F = G - 2 Z^2
          fr_add(t, a, b);
          fr_neg(a, t);
          vb_st(rf_sa + VB_OFF(4), a);   // H = -(A + B)
          break;
        }
        case VB_ADDMID: {
          u32 a[8], b[8], c[8], d[8], t[8];
          vb_ld(a, rf_sa + VB_OFF(4));
          vb_ld(b, rf_sa + VB_OFF(5));
          fr_sub(t, b, a);
          vb_st(rf_sa + VB_OFF(8), t);   // E = B - A
          fr_add(t, b, a);
          vb_st(rf_sa + VB_OFF(4), t);   // H = B + A
          vb_ld(c, rf_sa + VB_OFF(6));
          vb_ld(d, rf_sa + VB_OFF(7));
          if (neg) {                     // -Q: its 2dT changes sign
            fr_neg(t, c);
            fr_copy(c, t);
          }
          fr_sub(t, d, c);
          vb_st(rf_sa + VB_OFF(9), t);   // F = D - C
          fr_add(t, d, c);
          vb_st(rf_sa + VB_OFF(6), t);   // G = D + C
          break;
        }
        case VB_LDQB: {
          u32 x[8], y[8], t[8];
          vb_ld(x, rf_sa + VB_OFF(0));
          vb_ld(y, rf_sa + VB_OFF(1));
          fr_sub(t, y, x);
          vb_st(rf_sa + VB_OFF(4), t);
          fr_add(t, y, x);
          vb_st(rf_sa + VB_OFF(5), t);
#pragma unroll
          for (int c = 0; c < 4; c++) {
            u32 v[8];
            load_fr_plain(v, q + ((neg && c < 2) ? (c ^ 1) : c) * 8);
            vb_st(rf_sa + VB_OFF(6 + c), v);
          }
          break;
        }
        case VB_LDQ: {
#pragma unroll
          for (int c = 0; c < 4; c++) {
            u32 v[8];
            load_fr_plain(v, q + ((neg && c < 2) ? (c ^ 1) : c) * 8);
            vb_st(s1 + c * VB_OFF(1), v);
          }
          break;
        }
        default: {  // VB_STQ
#pragma unroll
          for (int c = 0; c < 4; c++) {
            u32 v[8];
            vb_ld(v, s2 + c * VB_OFF(1));
            store_fr(q + c * 8, v);
          }
          break;
        }
      }
    }
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

// ---------------------------------------------------------------------------------------------------------------------
// PRE / POST passes.  Cold code (a few tens of multiplies per item around ~2 300 per scalar multiplication).
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void store_ext(u32* o, const ExtPoint& p) {
  store_fr(o, p.X);
  store_fr(o + 8, p.Y);
  store_fr(o + 16, p.Z);
  store_fr(o + 24, p.T);
}
__device__ __forceinline__ void load_ext(ExtPoint& p, const u32* s) {
  load_fr(p.X, s);
  load_fr(p.Y, s + 8);
  load_fr(p.Z, s + 16);
  load_fr(p.T, s + 24);
}

__device__ __forceinline__ void ext_neg(ExtPoint& p) {
  u32 t[8];
  fr_neg(t, p.X);
  fr_copy(p.X, t);
  fr_neg(t, p.T);
  fr_copy(p.T, t);
}

// projective equality of two extended points (Z != 0 on both sides for curve points)
__device__ __noinline__ bool ext_equal(const ExtPoint& p, const ExtPoint& q) {
  u32 a[8], b[8];
  fr_mul(a, p.X, q.Z);
  fr_mul(b, q.X, p.Z);
  fr_canon(a);
  fr_canon(b);
  if (!eq256(a, b)) return false;
  fr_mul(a, p.Y, q.Z);
  fr_mul(b, q.Y, p.Z);
  fr_canon(a);
  fr_canon(b);
  return eq256(a, b);
}

// affine point from memory: canonical check, Montgomery conversion, on-curve check; result extended
__device__ __noinline__ void load_curve_point(ExtPoint& p, bool& canonical, bool& on_curve, const u32* src, int mont, int te = 0) {
  u32 xs[8], ys[8], x[8], y[8];
  load_fr(xs, src);
  load_fr(ys, src + 8);
  canonical = canonical && fr_is_canonical(xs) && fr_is_canonical(ys);
  if (mont) {
    fr_copy(x, xs);
    fr_copy(y, ys);
  } else {
    fr_to_mont(x, xs);
    fr_to_mont(y, ys);
  }
  if (te) te_to_rte_x(x);
  on_curve = on_curve && ed_is_on_curve(x, y);
  ext_from_affine(p, x, y);
}

__device__ __noinline__ void ext_add_ool(ExtPoint& p, const ExtPoint& q) { ext_add(p, q); }

// ---- curve.ScalarMul on its own: out = [s]P (+ [s2]P2) ----------------------------------------------------------------
// pre: canonical and on-curve checks (the reference's callers assert the point first: encrypt.go:49, ciphertext.go:53-54),
// bases and integer scalars for the window kernel; bases: n x n_bases x 32 words.
__global__ void __launch_bounds__(128) scalar_mul_pre_kernel(const u32* __restrict__ points, const u32* __restrict__ scalars,
                                                             const u32* __restrict__ points2, const u32* __restrict__ scalars2,
                                                             size_t n, int mont, int te, u8* __restrict__ status,
                                                             u32* __restrict__ bases, u32* __restrict__ k0, u32* __restrict__ k1) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  const int nb = points2 ? 2 : 1;
  bool canon = true, on_curve = true;
#pragma unroll 1
  for (int b = 0; b < nb; b++) {
    ExtPoint p;
    u32 k[8];
    load_curve_point(p, canon, on_curve, (b ? points2 : points) + idx * 16, mont, te);
    load_scalar(k, canon, (b ? scalars2 : scalars) + idx * 8, mont);
    store_ext(bases + (idx * nb + b) * 32, p);
    store_fr((b ? k1 : k0) + idx * 8, k);
  }
  status[idx] = !canon ? GCP_STATUS_NONCANONICAL : (!on_curve ? GCP_STATUS_OFF_CURVE : GCP_STATUS_OK);
}

// ---- Encrypt with a public key per item (encrypt.go:42-64) ----------------------------------------------------------
// pre: AssertIsOnCurve(pubKey) (:49), k as an integer; window kernel: S = [k]pubKey (:55); finish: C1 = [k]G (:52),
// C2 = [m]G + S (:58-61), one thread per point (whole warps per half), then normalize_kernel.
__global__ void __launch_bounds__(128) encrypt_per_key_pre_kernel(const u32* __restrict__ pks, const u32* __restrict__ ks,
                                                                  const u32* __restrict__ ms, size_t n, int mont, int te,
                                                                  u8* __restrict__ status, u32* __restrict__ bases,
                                                                  u32* __restrict__ kint) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  bool canon = true, on_curve = true;
  u32 k[8], m[8];
  load_scalar(k, canon, ks + idx * 8, mont);
  load_scalar(m, canon, ms + idx * 8, mont);
  ExtPoint pk;
  load_curve_point(pk, canon, on_curve, pks + idx * 16, mont, te);
  status[idx] = !canon ? GCP_STATUS_NONCANONICAL : (!on_curve ? GCP_STATUS_OFF_CURVE : GCP_STATUS_OK);
  store_ext(bases + idx * 32, pk);
  store_fr(kint + idx * 8, k);
}

__global__ void __launch_bounds__(128, 4) encrypt_per_key_finish_kernel(const u32* __restrict__ tabG, const u32* __restrict__ ks,
                                                                     const u32* __restrict__ ms, const u32* __restrict__ kpk,
                                                                     const u8* __restrict__ status, size_t n, int mont,
                                                                     u32* __restrict__ out_xyz) {
  const size_t pidx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (pidx >= 2 * n) return;
  const bool second = pidx >= n;
  const size_t idx = second ? pidx - n : pidx;
  ExtPoint c;
  ext_identity(c);
  if (status[idx] == GCP_STATUS_OK) {
    bool canon = true;
    u32 sc[8];
    load_scalar(sc, canon, (second ? ms : ks) + idx * 8, mont);
    if (second) load_ext(c, kpk + idx * 32);
    fixed_base_accumulate(c, sc, tabG);
  }
  store_ext_xyz(out_xyz + (idx * 2 + (second ? 1 : 0)) * 24, c);
}

// ---- AssertDecrypt: C1, C2 on the curve;  C2 - [priv]C1 == [m]G  (ciphertext.go:50-67) ---------------------------------
// pre: rhs = C2 - [m]G;  window kernel: S = [priv]C1;  post: flag = (S == rhs)
__global__ void __launch_bounds__(128, 4) assert_decrypt_pre_kernel(const u32* __restrict__ tabG, const u32* __restrict__ cts,
                                                                 const u32* __restrict__ privs, const u32* __restrict__ msgs,
                                                                 size_t n, int mont, int te, u8* __restrict__ status,
                                                                 u32* __restrict__ bases, u32* __restrict__ kint,
                                                                 u32* __restrict__ rhs) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  bool canon = true, on_curve = true;
  ExtPoint c1, c2;
  u32 priv[8], msg[8];
  load_curve_point(c1, canon, on_curve, cts + idx * 32, mont, te);
  load_curve_point(c2, canon, on_curve, cts + idx * 32 + 16, mont, te);
  load_scalar(priv, canon, privs + idx * 8, mont);
  load_scalar(msg, canon, msgs + idx * 8, mont);
  const u8 st = !canon ? GCP_STATUS_NONCANONICAL : (!on_curve ? GCP_STATUS_OFF_CURVE : GCP_STATUS_OK);
  status[idx] = st;
  if (st == GCP_STATUS_OK) {
    ExtPoint m;
    ext_identity(m);
    fixed_base_accumulate(m, msg, tabG);  // ciphertext.go:60
    ext_neg(m);
    ext_add_ool(c2, m);
  }
  store_ext(bases + idx * 32, c1);
  store_fr(kint + idx * 8, priv);
  store_ext(rhs + idx * 32, c2);
}

__global__ void __launch_bounds__(128) ext_compare_kernel(const u32* __restrict__ lhs, const u32* __restrict__ rhs,
                                                          const u8* __restrict__ status, size_t n, u8* __restrict__ flags) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  u8 flag = 0;
  if (status[idx] == GCP_STATUS_OK) {
    ExtPoint p, q;
    load_ext(p, lhs + idx * 32);
    load_ext(q, rhs + idx * 32);
    flag = ext_equal(p, q) ? 1 : 0;  // ciphertext.go:64-65
  }
  flags[idx] = flag;
}

// ---- DecryptionProof.Verify (ciphertext.go:124-168) -----------------------------------------------------------------
// pre: D = C2 - [msg]G (:137-139), the 12 hash inputs PK, PK, C1, D, A1, A2 (:141, hashPointsToScalar :173-184; MultiHash
// of 12 inputs is one Hash with t = 13), zG = [z]G;  Poseidon batch kernel: e;  window kernel twice: [e]PK and the
// double-scalar [z]C1 - [e]D;  post: zG == A1 + [e]PK (:143-151) and [z]C1 - [e]D == A2 (:153-166).
__global__ void __launch_bounds__(128, 4) decryption_proof_pre_kernel(const u32* __restrict__ tabG, const u32* __restrict__ pks,
                                                                   const u32* __restrict__ cts, const u32* __restrict__ msgs,
                                                                   const u32* __restrict__ a1s, const u32* __restrict__ a2s,
                                                                   const u32* __restrict__ zs, size_t n, int mont, int te,
                                                                   u8* __restrict__ status, u32* __restrict__ hash_in,
                                                                   u32* __restrict__ zg, u32* __restrict__ base_pk,
                                                                   u32* __restrict__ base_c1_d, u32* __restrict__ zint) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  bool canon = true, on_curve = true;
  u32* hin = hash_in + idx * 96;
  u32 msg[8], z[8];
  ExtPoint c2;
  // each point is loaded, validated and written out at once (one live ExtPoint at a time keeps this pass in registers)
#pragma unroll 1
  for (int which = 0; which < 4; which++) {  // PK, C1, A1, A2
    const u32* src = which == 0 ? pks + idx * 16 : (which == 1 ? cts + idx * 32 : (which == 2 ? a1s + idx * 16 : a2s + idx * 16));
    ExtPoint p;
    load_curve_point(p, canon, on_curve, src, mont, te);
    // hash input positions (elements): PK at 0,1 and 2,3; C1 at 4,5; (D at 6,7); A1 at 8,9; A2 at 10,11
    const int pos = which == 0 ? 0 : (which == 1 ? 4 : (which == 2 ? 8 : 10));
    u32 cx[8], cy[8];
    fr_copy(cx, p.X);
    fr_copy(cy, p.Y);
    fr_canon(cx);
    fr_canon(cy);
    store_fr(hin + pos * 8, cx);
    store_fr(hin + pos * 8 + 8, cy);
    if (which == 0) {
      store_fr(hin + 16, cx);
      store_fr(hin + 24, cy);
      store_ext(base_pk + idx * 32, p);
    } else if (which == 1) {
      store_ext(base_c1_d + idx * 64, p);
    }
  }
  load_curve_point(c2, canon, on_curve, cts + idx * 32 + 16, mont, te);
  load_scalar(msg, canon, msgs + idx * 8, mont);
  load_scalar(z, canon, zs + idx * 8, mont);
  u8 st = !canon ? GCP_STATUS_NONCANONICAL : (!on_curve ? GCP_STATUS_OFF_CURVE : GCP_STATUS_OK);
  if (st == GCP_STATUS_OK) {
    ExtPoint m;
    ext_identity(m);
    fixed_base_accumulate(m, msg, tabG);
    ext_neg(m);
    ext_add_ool(c2, m);  // D
    u32 zc[8];
    fr_copy(zc, c2.Z);
    fr_canon(zc);
    if (is_zero256(zc)) {
      st = GCP_STATUS_ZERO_DENOM;
    } else {
      u32 zi[8], dx[8], dy[8];
      fr_inv(zi, c2.Z);
      fr_mul(dx, c2.X, zi);
      fr_mul(dy, c2.Y, zi);
      fr_canon(dx);
      fr_canon(dy);
      store_fr(hin + 48, dx);
      store_fr(hin + 56, dy);
      ext_neg(c2);  // the window kernel adds: [z]C1 + [e](-D)
      store_ext(base_c1_d + idx * 64 + 32, c2);
      ext_identity(m);
      fixed_base_accumulate(m, z, tabG);
      store_ext(zg + idx * 32, m);
    }
  }
  store_fr(zint + idx * 8, z);
  status[idx] = st;
}

__global__ void __launch_bounds__(128) decryption_proof_post_kernel(const u32* __restrict__ a1s, const u32* __restrict__ a2s,
                                                                    const u32* __restrict__ zg, const u32* __restrict__ epk,
                                                                    const u32* __restrict__ zc1_ed, const u8* __restrict__ status,
                                                                    size_t n, int mont, int te, u8* __restrict__ flags) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  u8 flag = 0;
  if (status[idx] == GCP_STATUS_OK) {
    bool c = true, oc = true;
    ExtPoint a, l, r;
    load_curve_point(a, c, oc, a1s + idx * 16, mont, te);
    load_ext(r, epk + idx * 32);
    ext_add_ool(r, a);  // A1 + e*P
    load_ext(l, zg + idx * 32);
    bool ok = ext_equal(l, r);
    load_curve_point(a, c, oc, a2s + idx * 16, mont, te);
    load_ext(l, zc1_ed + idx * 32);  // z*C1 - e*D
    ok = ok && ext_equal(l, a);
    flag = ok ? 1 : 0;
  }
  flags[idx] = flag;
}

// ---- EdDSA-Poseidon IsValid (/root/reference/ecc/bn254/eddsa/verifier.go:55-88) ----------------------------------------
// A, R in TE (circom/iden3) coordinates; h = Poseidon(R.x, R.y, A.x, A.y, msg) on those coordinates (t = 6);
// A' = RTE(A), R' = RTE(R) asserted on the a = -1 curve; flag = ([S]G == 8*[h]A' + R')  (rteB8 == G, constants.go:11-18).

__global__ void __launch_bounds__(128, 4) eddsa_pre_kernel(const u32* __restrict__ tabG, const u32* __restrict__ pub_a,
                                                        const u32* __restrict__ sig_r, const u32* __restrict__ sig_s,
                                                        const u32* __restrict__ msgs, size_t n, int mont, u8* __restrict__ status,
                                                        u32* __restrict__ hash_in, u32* __restrict__ left, u32* __restrict__ base_a) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  const u32 negf[8] = GCP_NEG_F_MONT;
  bool canon = true;
  u32 e[5][8], s_int[8];
  load_elem(e[0], canon, sig_r + idx * 16, mont);
  load_elem(e[1], canon, sig_r + idx * 16 + 8, mont);
  load_elem(e[2], canon, pub_a + idx * 16, mont);
  load_elem(e[3], canon, pub_a + idx * 16 + 8, mont);
  load_elem(e[4], canon, msgs + idx * 8, mont);
  load_scalar(s_int, canon, sig_s + idx * 8, mont);
  u32* hin = hash_in + idx * 40;
#pragma unroll
  for (int j = 0; j < 5; j++) {
    u32 c[8];
    fr_copy(c, e[j]);
    fr_canon(c);
    store_fr(hin + j * 8, c);
  }
  u32 ax[8], rx[8];
  fr_mul(rx, e[0], negf);  // RTE: x * (-f), y unchanged (ecc/format/twistededwards.go:42-48)
  fr_mul(ax, e[2], negf);
  const bool on_curve = ed_is_on_curve(ax, e[3]) && ed_is_on_curve(rx, e[1]);  // PointToRTE, verifier.go:46
  const u8 st = !canon ? GCP_STATUS_NONCANONICAL : (!on_curve ? GCP_STATUS_OFF_CURVE : GCP_STATUS_OK);
  status[idx] = st;
  ExtPoint a, l;
  ext_from_affine(a, ax, e[3]);
  store_ext(base_a + idx * 32, a);
  ext_identity(l);
  if (st == GCP_STATUS_OK) fixed_base_accumulate(l, s_int, tabG);  // [S] rteB8
  store_ext(left + idx * 32, l);
}

__global__ void __launch_bounds__(128) eddsa_post_kernel(const u32* __restrict__ sig_r, const u32* __restrict__ left,
                                                         const u32* __restrict__ ha, const u8* __restrict__ status, size_t n,
                                                         int mont, u8* __restrict__ flags) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  u8 flag = 0;
  if (status[idx] == GCP_STATUS_OK) {
    const u32 negf[8] = GCP_NEG_F_MONT;
    bool canon = true;
    u32 x[8], y[8], rx[8];
    load_elem(x, canon, sig_r + idx * 16, mont);
    load_elem(y, canon, sig_r + idx * 16 + 8, mont);
    fr_mul(rx, x, negf);
    ExtPoint r, r1, l;
    ext_from_affine(r, rx, y);
    load_ext(r1, ha + idx * 32);
#pragma unroll 1
    for (int d = 0; d < 3; d++) ext_double(r1);  // verifier.go:72-74
    ext_add_ool(r1, r);
    load_ext(l, left + idx * 32);
    flag = ext_equal(l, r1) ? 1 : 0;
  }
  flags[idx] = flag;
}

}  // namespace gcp
