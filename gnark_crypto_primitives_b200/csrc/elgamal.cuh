// ElGamal over the a = -1 BN254 twisted Edwards curve: fixed-base tables, encrypt, ciphertext add, tally.
//
// Values reproduced (file:line under /root/reference):
//   FixedBaseScalarMulBN254  elgamal/mul.go:76-166   == [s]G for s in [0, r)  (64 4-bit windows over table
//                            mul.go:26-72; zero nibbles skipped, window 0 initialises)
//   (*Ciphertext).Encrypt    elgamal/encrypt.go:42-64  C1 = [k]G, C2 = [m]G + [k]PK, AssertIsOnCurve(PK)
//   EncryptedZero            elgamal/encrypt.go:72-94  == Encrypt(k, 0)
//   (*Ciphertext).Add        elgamal/ciphertext.go:24-32 component-wise curve.Add;  Neg :37-46
//   tally                    left fold of Add from NewCiphertext (ciphertext.go:16-19) — caller's loop
// All of these are group elements given by exact field arithmetic, so any addition order / window width yields
// the same canonical affine coordinates.  This file uses signed w-bit windows over precomputed Niels tables
// (tab[i][d-1] = [d * 2^(w*i)] B, d = 1 .. 2^(w-1), w = 20 or 24) for both G and a shared public key, accumulates in extended coordinates,
// and converts to affine with chunked Montgomery batch inversion (one Fermat inversion per BATCH_INV ... BATCH_INV_MAX
// points, the launcher's choice: normalize_kernel).
#pragma once
#include "edwards.cuh"
#include "kernels.h"

namespace gcp {

// Signed fixed windows: a scalar is recoded into digits in [-2^(w-1)+1, 2^(w-1)]; the table holds the positive
// multiples only (negating a Niels point is a swap and one field negation).  The window width w is a property of the
// TABLE (a 128-byte header in front of the entries: {w, windows, entries per window}), so every kernel that takes a
// table pointer works with any width and a context can replace a table by a wider one while it runs:
//   w = 20: 13 windows x 2^19 entries x 96 B = 654 MB per base, 13 mixed additions per scalar multiplication (the
//           reference's 4-bit table, mul.go:26-72, needs up to 63; the first version of this file used w = 14, 19
//           additions out of a 14.9 MB L2-resident table); built in ~2 ms, the width every table starts with;
//   w = 24: 11 windows x 2^23 entries = 8.9 GB per base, 11 additions: fused encrypt + tally 308 -> 351 M enc/s on
//           2^24 ballots x 8 fields (w = 22: 332 M, w = 26 with 32 GB per base: 339 M - its random reads start to
//           cost more than the saved addition, profiles/r02_fixed_base_window_sweep.jsonl); ~45 ms to build, used once
//           a base has served enough multiplications to repay that (capi.cu: fb_note_use).  180 GB of HBM per GPU is
//           what makes a 9 GB table per base a reasonable trade.
// A lookup is one random 96-byte read (23-27 of them per ciphertext, ~0.8 TB/s, an eighth of the HBM bandwidth), staged
// one window ahead through shared memory (cp.async) so that its latency hides behind the 7-multiply addition of the
// current window.
constexpr int FB_MIN_WBITS = 8, FB_MAX_WBITS = 26;
constexpr int FB_HEADER_WORDS = FB_TABLE_HEADER_WORDS;  // kernels.h
__host__ __device__ inline int fb_windows(int wbits) { return (256 + wbits - 1) / wbits; }  // 254-bit scalars + recoding carry
__host__ __device__ inline int fb_lo_bits(int wbits) { return wbits / 2; }
// small tables per window (table construction): [j] B_w for j < 2^lo and [j 2^lo] B_w for j <= 2^(w-1-lo)
__host__ __device__ inline int fb_small_per_window(int wbits) {
  return (1 << fb_lo_bits(wbits)) + (1 << (wbits - 1 - fb_lo_bits(wbits))) + 1;
}
constexpr int BATCH_INV = 32;
constexpr int BATCH_INV_MAX = 128;  // normalize_kernel on large batches

// ---- table construction (one-time per base point) ------------------------------------------------------
// base: affine (x, y), 16 words, standard or Montgomery form.  flag[0] = 1 if B is on the curve and canonical, else 0.
// iden3 / circom twisted-Edwards coordinates at the boundary (ecc/format/twistededwards.go:29-48):
// FromTEtoRTE(x, y) = (x * (-f), y), FromRTEtoTE(x, y) = (x / (-f), y).  With GCP_COORDS_TE set in the format argument
// every point a kernel reads is converted on load and every point it writes on store, one multiply each, so callers that
// hold circom-side points need no separate conversion pass (and no extra PCIe round trip).  The constants are c * R:
// mont_mul(x, c * R) = x * c in whichever representation x is in.
#define GCP_NEG_F_MONT {0xc9603c7bu, 0x5c62c8e0u, 0x8fabc7f1u, 0xf8382911u, 0x6aa07f4du, 0x7d53da81u, 0x6ba06ab6u, 0x1da7c5b3u}
#define GCP_NEG_F_INV_MONT {0xb1b017d8u, 0x61d380bfu, 0x8415d72eu, 0x7f5d8063u, 0x294f7a18u, 0x77e18e30u, 0x305733c2u, 0x10d2ede5u}
__device__ __forceinline__ void te_to_rte_x(u32 (&x)[8]) {
  const u32 negf[8] = GCP_NEG_F_MONT;
  u32 t[8];
  fr_mul(t, x, negf);
  fr_copy(x, t);
}
__device__ __forceinline__ void rte_to_te_x(u32 (&x)[8]) {
  const u32 negf_inv[8] = GCP_NEG_F_INV_MONT;
  u32 t[8];
  fr_mul(t, x, negf_inv);
  fr_copy(x, t);
}

// Table construction, four launches (round 1 built every entry by its own double-and-add and inverted every Z by its own
// Fermat chain: ~70 k wide multiplies per entry, 0.1 s for w = 20; now ~3 k per entry):
//   bases   B_w = [2^(w_bits w)] B per window, the header, flag[0] = 1 iff B is canonical and on the curve
//   small   per window [j] B_w (j < 2^lo) and [j 2^lo] B_w (j <= 2^(w_bits-1-lo)) by double-and-add (a few thousand entries)
//   sum     entry d = hi 2^lo + lo  is  [hi 2^lo] B_w + [lo] B_w: ONE unified addition per entry, (X, Y, Z) written into
//           the entry's own 96 bytes
//   niels   in place: Montgomery batch inversion over 32 strided entries per thread, (y-x, y+x, 2dxy) canonical
// small: windows x 32 words (the B_w) followed by windows x fb_small_per_window x 32 words.
__global__ void fb_table_bases_kernel(const u32* __restrict__ base, int base_mont, int te, u32* __restrict__ small,
                                      u32* __restrict__ header, u32* __restrict__ flag, int wbits) {
  int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= fb_windows(wbits)) return;
  u32 xs[8], ys[8], x[8], y[8];
  load_fr(xs, base);
  load_fr(ys, base + 8);
  bool ok = fr_is_canonical(xs) && fr_is_canonical(ys);
  if (base_mont) {
    fr_copy(x, xs);
    fr_copy(y, ys);
  } else {
    fr_to_mont(x, xs);
    fr_to_mont(y, ys);
  }
  if (te) te_to_rte_x(x);
  ok = ok && ed_is_on_curve(x, y);
  if (w == 0) {
    flag[0] = ok ? 1u : 0u;
    header[0] = (u32)wbits;
    header[1] = (u32)fb_windows(wbits);
    header[2] = 1u << (wbits - 1);
  }
  ExtPoint p;
  if (ok)
    ext_from_affine(p, x, y);
  else
    ext_identity(p);  // keeps every later kernel well defined; results are masked by the flag
#pragma unroll 1
  for (int i = 0; i < w * wbits; i++) ext_double(p);
  u32* o = small + (size_t)w * 32;
  store_fr(o, p.X);
  store_fr(o + 8, p.Y);
  store_fr(o + 16, p.Z);
  store_fr(o + 24, p.T);
}

__global__ void fb_table_small_kernel(u32* __restrict__ small, int wbits) {
  const int windows = fb_windows(wbits), per = fb_small_per_window(wbits), lo_bits = fb_lo_bits(wbits);
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= windows * per) return;
  const int w = idx / per, j = idx % per;
  const u32 scalar = j < (1 << lo_bits) ? (u32)j : (u32)(j - (1 << lo_bits)) << lo_bits;
  ExtPoint b, acc;
  const u32* s = small + (size_t)w * 32;
  load_fr(b.X, s);
  load_fr(b.Y, s + 8);
  load_fr(b.Z, s + 16);
  load_fr(b.T, s + 24);
  ext_identity(acc);
#pragma unroll 1
  for (int bit = wbits - 1; bit >= 0; bit--) {
    ext_double(acc);
    if ((scalar >> bit) & 1u) ext_add(acc, b);
  }
  u32* o = small + ((size_t)windows + idx) * 32;
  store_fr(o, acc.X);
  store_fr(o + 8, acc.Y);
  store_fr(o + 16, acc.Z);
  store_fr(o + 24, acc.T);
}

__global__ void __launch_bounds__(128) fb_table_sum_kernel(const u32* __restrict__ small, u32* __restrict__ tab, int wbits) {
  const int windows = fb_windows(wbits), per = fb_small_per_window(wbits), lo_bits = fb_lo_bits(wbits);
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= ((size_t)windows << (wbits - 1))) return;
  const int w = (int)(idx >> (wbits - 1));
  const u32 d = (u32)(idx & (((size_t)1 << (wbits - 1)) - 1)) + 1u;
  const u32 hi = d >> lo_bits, lo = d & ((1u << lo_bits) - 1u);
  const u32* sw = small + ((size_t)windows + (size_t)w * per) * 32;
  ExtPoint p, q;
  const u32* sp = sw + ((size_t)(1u << lo_bits) + hi) * 32;
  const u32* sq = sw + (size_t)lo * 32;
  load_fr(p.X, sp);
  load_fr(p.Y, sp + 8);
  load_fr(p.Z, sp + 16);
  load_fr(p.T, sp + 24);
  load_fr(q.X, sq);
  load_fr(q.Y, sq + 8);
  load_fr(q.Z, sq + 16);
  load_fr(q.T, sq + 24);
  ext_add(p, q);
  u32* o = tab + idx * 24;
  store_fr(o, p.X);
  store_fr(o + 8, p.Y);
  store_fr(o + 16, p.Z);
}

// (X, Y, Z) -> Niels (y-x, y+x, 2dxy), canonical Montgomery, in place.  Every entry is read and written by one thread only.
__global__ void __launch_bounds__(128) fb_table_niels_kernel(u32* __restrict__ tab, size_t total) {
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  if (tid >= total) return;
  u32 pre[BATCH_INV][8];  // prefix products (local memory)
  u32 acc[8];
  fr_set_one(acc);
  int cnt = 0;
#pragma unroll 1
  for (int j = 0; j < BATCH_INV; j++) {
    const size_t p = tid + (size_t)j * stride;
    if (p >= total) break;
    u32 z[8];
    load_fr_plain(z, tab + p * 24 + 16);
    fr_mul(acc, acc, z);  // Z != 0: the addition law is complete on the curve, and an off-curve base was replaced by O
#pragma unroll
    for (int l = 0; l < 8; l++) pre[j][l] = acc[l];
    cnt++;
  }
  u32 inv[8];
  fr_inv(inv, acc);
  const u32 d2[8] = GCP_ED_2D_MONT;
#pragma unroll 1
  for (int j = cnt - 1; j >= 0; j--) {
    const size_t p = tid + (size_t)j * stride;
    u32* e = tab + p * 24;
    u32 z[8], zi[8], X[8], Y[8], x[8], y[8], t[8];
    if (j > 0) {
      u32 prev[8];
#pragma unroll
      for (int l = 0; l < 8; l++) prev[l] = pre[j - 1][l];
      load_fr_plain(z, e + 16);
      fr_mul(zi, inv, prev);  // 1 / z_j
      fr_mul(inv, inv, z);    // 1 / (z_0 ... z_{j-1})
    } else {
      fr_copy(zi, inv);
    }
    load_fr_plain(X, e);
    load_fr_plain(Y, e + 8);
    fr_mul(x, X, zi);
    fr_mul(y, Y, zi);
    NielsPoint n;
    fr_sub(n.ymx, y, x);
    fr_add(n.ypx, y, x);
    fr_mul(t, x, y);
    fr_mul(n.t2d, t, d2);
    fr_canon(n.ymx);
    fr_canon(n.ypx);
    fr_canon(n.t2d);
    store_fr(e, n.ymx);
    store_fr(e + 8, n.ypx);
    store_fr(e + 16, n.t2d);
  }
}

// window at bit offset `bit`, wbits (<= 26) bits wide, of a 256-bit little-endian scalar.  The limb index is dynamic: the
// scalar lives in 32 bytes of local memory (two L1-resident loads per window), which costs less than the registers or the
// select chains that keeping it in registers would (measured both: +8 registers spill in the encrypt kernels)
__device__ __forceinline__ u32 fb_scalar_window(const u32 (&k)[8], int bit, int wbits) {
  const int limb = bit >> 5, sh = bit & 31;
  const u32 lo = k[limb], hi = (limb + 1 < 8) ? k[limb + 1] : 0u;
  return __funnelshift_r(lo, hi, sh) & ((1u << wbits) - 1u);
}

// acc += [k] B using B's table; k is an integer < 2^254 (a canonical Fr element used as an integer)
__device__ __forceinline__ void fb_digit(const u32 (&k)[8], int w, int wbits, u32& carry, u32& d, bool& neg) {
  u32 raw = fb_scalar_window(k, w * wbits, wbits) + carry;
  neg = raw > (1u << (wbits - 1));
  d = neg ? ((1u << wbits) - raw) : raw;
  carry = neg ? 1u : 0u;
}

// Table entries are staged through shared memory with cp.async, one window ahead: the 96-byte read of window w+1
// (a random HBM access: the 654 MB table does not live in L2, and a `prefetch` hint is dropped on a TLB miss - measured:
// no effect) is in flight while the 7-multiply addition of window w runs, and costs no registers.  Slot layout
// [buffer][16-byte piece][thread]: every cp.async / LDS.128 of a warp touches 32 consecutive 16-byte words.
constexpr int FB_STAGE_THREADS = 128;  // every kernel that calls fixed_base_accumulate has at most 128 threads per block

__device__ __forceinline__ void fb_stage_issue(uint4* stage, int buf, const u32* e) {
#pragma unroll
  for (int c = 0; c < 6; c++) {
    unsigned saddr = (unsigned)__cvta_generic_to_shared(stage + (buf * 6 + c) * FB_STAGE_THREADS + threadIdx.x);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(saddr), "l"(e + c * 4) : "memory");
  }
}

__device__ __forceinline__ void fb_stage_read(u32 (&a)[8], u32 (&b)[8], u32 (&c)[8], const uint4* stage, int buf) {
  uint4 v[6];
#pragma unroll
  for (int q = 0; q < 6; q++) v[q] = stage[(buf * 6 + q) * FB_STAGE_THREADS + threadIdx.x];
  a[0] = v[0].x; a[1] = v[0].y; a[2] = v[0].z; a[3] = v[0].w; a[4] = v[1].x; a[5] = v[1].y; a[6] = v[1].z; a[7] = v[1].w;
  b[0] = v[2].x; b[1] = v[2].y; b[2] = v[2].z; b[3] = v[2].w; b[4] = v[3].x; b[5] = v[3].y; b[6] = v[3].z; b[7] = v[3].w;
  c[0] = v[4].x; c[1] = v[4].y; c[2] = v[4].z; c[3] = v[4].w; c[4] = v[5].x; c[5] = v[5].y; c[6] = v[5].z; c[7] = v[5].w;
}

// How many windows ahead the table reads run (FB_AHEAD + 1 staging buffers of 12 KB per block).  One window ahead hides
// ~7 us of addition behind each read; with 24-bit tables (8.9 GB per base, TLB and DRAM-page misses on every read) ncu
// still showed 0.37 long-scoreboard stalls per issue, so the depth is a build-time number that was measured.
#ifndef GCP_FB_AHEAD
#define GCP_FB_AHEAD 1
#endif
constexpr int FB_AHEAD = GCP_FB_AHEAD;
constexpr int FB_BUFS = FB_AHEAD + 1;
static_assert(FB_AHEAD >= 1 && FB_AHEAD <= 3, "staging depth");

__device__ __forceinline__ void fixed_base_accumulate(ExtPoint& acc, const u32 (&k)[8], const u32* __restrict__ tab) {
  __shared__ uint4 stage[FB_BUFS * 6 * FB_STAGE_THREADS];
  // the table's header gives the width; the only loop state beyond round 1's (w, tab) is that one register: the window
  // count is "while the next window starts below bit 256" and the table pointer advances by one window per iteration
  const int wbits = (int)__ldg(tab - FB_HEADER_WORDS);
  const size_t wstep = (size_t)24 << (wbits - 1);
  u32 carry = 0;
  u32 dq[FB_AHEAD];   // digits of the windows in flight, oldest first
  bool nq[FB_AHEAD];
#pragma unroll
  for (int i = 0; i < FB_AHEAD; i++) {
    dq[i] = 0;
    nq[i] = false;
    if (i * wbits < 256) {
      fb_digit(k, i, wbits, carry, dq[i], nq[i]);
      if (dq[i] != 0) fb_stage_issue(stage, i, tab + (size_t)(dq[i] - 1) * 24);
      tab += wstep;
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  int w = 0, rbuf = 0, wbuf = FB_AHEAD;  // buffer of window w / of window w + FB_AHEAD
#pragma unroll 1
  do {
    u32 dn = 0;
    bool negn = false;
    if ((w + FB_AHEAD) * wbits < 256) {  // a later window's entry is on its way while this window's addition runs
      fb_digit(k, w + FB_AHEAD, wbits, carry, dn, negn);
      if (dn != 0) fb_stage_issue(stage, wbuf, tab + (size_t)(dn - 1) * 24);
      tab += wstep;
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group %0;" ::"n"(FB_AHEAD) : "memory");  // all but the FB_AHEAD newest groups have landed
    const u32 d = dq[0];
    const bool neg = nq[0];
    if (d != 0) {
      NielsPoint n;
      if (neg) {  // -(x, y) = (-x, y): swaps y-x and y+x, negates 2dxy
        u32 t[8];
        fb_stage_read(n.ypx, n.ymx, t, stage, rbuf);
        fr_neg(n.t2d, t);
      } else {
        fb_stage_read(n.ymx, n.ypx, n.t2d, stage, rbuf);
      }
      ext_add_niels(acc, n);
    }
#pragma unroll
    for (int i = 0; i + 1 < FB_AHEAD; i++) {
      dq[i] = dq[i + 1];
      nq[i] = nq[i + 1];
    }
    dq[FB_AHEAD - 1] = dn;
    nq[FB_AHEAD - 1] = negn;
    rbuf = (rbuf + 1 == FB_BUFS) ? 0 : rbuf + 1;
    wbuf = (wbuf + 1 == FB_BUFS) ? 0 : wbuf + 1;
    w++;
  } while (w * wbits < 256);
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}

__device__ __forceinline__ void store_ext_xyz(u32* o, const ExtPoint& p) {
  store_fr(o, p.X);
  store_fr(o + 8, p.Y);
  store_fr(o + 16, p.Z);
}

__device__ __forceinline__ void load_scalar(u32 (&k)[8], bool& canonical, const u32* p, int mont) {
  u32 x[8];
  load_fr(x, p);
  canonical = canonical && fr_is_canonical(x);
  if (mont)
    fr_from_mont(k, x);  // scalars are integers: leave Montgomery form
  else
    fr_copy(k, x);
}

// ---- [s]G -------------------------------------------------------------------------------------------------
// out_xyz: n x 24 words (X, Y, Z).  status: n bytes.
__global__ void __launch_bounds__(128) fixed_base_mul_kernel(const u32* __restrict__ tabG, const u32* __restrict__ scalars,
                                                             size_t n, u32* __restrict__ out_xyz, u8* __restrict__ status,
                                                             int mont) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  bool canon = true;
  u32 k[8];
  load_scalar(k, canon, scalars + idx * 8, mont);
  ExtPoint acc;
  ext_identity(acc);
  if (canon) fixed_base_accumulate(acc, k, tabG);
  store_ext_xyz(out_xyz + idx * 24, acc);
  status[idx] = canon ? GCP_STATUS_OK : GCP_STATUS_NONCANONICAL;
}

// ---- Encrypt with one shared public key (the election key) ----------------------------------------------------
// One thread per POINT (2n threads): threads [0, n) compute C1 = [k]G, threads [n, 2n) compute C2 = [k]PK + [m]G
// (whole warps take the same branch), as (X, Y, Z) into out_xyz (n x 2 x 24 words).  Both halves validate k and m.
__global__ void __launch_bounds__(128, 4) encrypt_shared_kernel(const u32* __restrict__ tabG, const u32* __restrict__ tabPK,
                                                             const u32* __restrict__ pk_flag, const u32* __restrict__ ks,
                                                             const u32* __restrict__ ms, size_t n, u32* __restrict__ out_xyz,
                                                             u8* __restrict__ status, int mont) {
  size_t pidx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (pidx >= 2 * n) return;
  const bool second = pidx >= n;
  size_t idx = second ? pidx - n : pidx;
  bool canon = true;
  u32 k[8], m[8];
  load_scalar(k, canon, ks + idx * 8, mont);
  load_scalar(m, canon, ms + idx * 8, mont);
  bool pk_ok = pk_flag[0] != 0;
  ExtPoint c;
  ext_identity(c);
  if (canon && pk_ok) {
    // C1 = [k]G (encrypt.go:52); C2 = [k]PK (:55) + [m]G (:58,61).  One inlined copy of the window loop / mixed
    // addition serves all three: three copies (~80 KB of SASS) made the kernel instruction-fetch bound.
    const int n_pass = second ? 2 : 1;
#pragma unroll 1
    for (int pass = 0; pass < n_pass; pass++) {
      const u32* tab = (second && pass == 0) ? tabPK : tabG;
      u32 sc[8];
#pragma unroll
      for (int l = 0; l < 8; l++) sc[l] = (pass == 0) ? k[l] : m[l];
      fixed_base_accumulate(c, sc, tab);
    }
  }
  store_ext_xyz(out_xyz + (idx * 2 + (second ? 1 : 0)) * 24, c);
  if (!second) status[idx] = !canon ? GCP_STATUS_NONCANONICAL : (!pk_ok ? GCP_STATUS_OFF_CURVE : GCP_STATUS_OK);
}

// ---- (X, Y, Z) -> canonical affine, Montgomery batch inversion over BATCH_INV points per thread ----------------
// xyz: n_points x xyz_words words (24, or 32 for extended points with T behind Z); out: n_points x 16 words.  status (optional) is indexed by point / pts_per_item.
// `per` (<= BATCH_INV_MAX) points share one inversion: the launcher raises it above BATCH_INV once the batch is large
// enough to fill the machine anyway.  A Fermat inversion is ~325 multiplications: 10 per point at 32 points per thread
// (39 % of a Ciphertext.Add, 9 % of an Encrypt), 2.5 at 128; the prefix products cost 32 bytes of local memory per point.
__global__ void __launch_bounds__(128) normalize_kernel(const u32* __restrict__ xyz, size_t n_points, u32* __restrict__ out,
                                                        u8* __restrict__ status, int pts_per_item, int mont, int xyz_words,
                                                        int te, int per) {
  size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  if (tid >= n_points) return;
  // Both passes are a serial chain of multiplications fed by loads nothing else depends on; left to themselves the loads
  // of iteration j are issued at its top and the chain waits a DRAM latency per point (ncu: 1.9 long-scoreboard stalls per
  // issue).  Forward: the next Z is loaded into registers before the current multiplication.  Backward: the next record
  // (X | Y | Z, 96 contiguous bytes) is staged through shared memory with cp.async, like the fixed-base table entries.
  __shared__ uint4 stage[2 * 6 * FB_STAGE_THREADS];
  u32 pre[BATCH_INV_MAX][8];  // prefix products (local memory)
  u32 acc[8], zn[8];
  fr_set_one(acc);
  int cnt = 0;
  load_fr(zn, xyz + tid * (size_t)xyz_words + 16);
#pragma unroll 1
  for (int j = 0; j < per; j++) {
    size_t p = tid + (size_t)j * stride;
    if (p >= n_points) break;
    u32 z[8];
    fr_copy(z, zn);
    if (j + 1 < per && p + stride < n_points) load_fr(zn, xyz + (p + stride) * (size_t)xyz_words + 16);
    u32 zc[8];
    fr_copy(zc, z);
    fr_canon(zc);
    if (is_zero256(zc)) {  // zero denominator: flag it and keep the batch invertible
      if (status) status[p / pts_per_item] = GCP_STATUS_ZERO_DENOM;
      fr_set_one(z);
    }
    fr_mul(acc, acc, z);
#pragma unroll
    for (int l = 0; l < 8; l++) pre[j][l] = acc[l];
    cnt++;
  }
  u32 inv[8];
  fr_inv(inv, acc);
  fb_stage_issue(stage, (cnt - 1) & 1, xyz + (tid + (size_t)(cnt - 1) * stride) * (size_t)xyz_words);
  asm volatile("cp.async.commit_group;" ::: "memory");
#pragma unroll 1
  for (int j = cnt - 1; j >= 0; j--) {
    size_t p = tid + (size_t)j * stride;
    u32 z[8], zi[8], X[8], Y[8], x[8], y[8];
    if (j > 0) fb_stage_issue(stage, (j - 1) & 1, xyz + (p - stride) * (size_t)xyz_words);
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 1;" ::: "memory");
    fb_stage_read(X, Y, z, stage, j & 1);
    {
      u32 zc[8];
      fr_copy(zc, z);
      fr_canon(zc);
      if (is_zero256(zc)) fr_set_one(z);
    }
    if (j > 0) {
      u32 prev[8];
#pragma unroll
      for (int l = 0; l < 8; l++) prev[l] = pre[j - 1][l];
      fr_mul(zi, inv, prev);   // 1 / z_j
      fr_mul(inv, inv, z);     // 1 / (z_0 ... z_{j-1})
    } else {
      fr_copy(zi, inv);
    }
    if (!mont) fr_from_mont(zi, zi);  // standard-form output: take the factor R off 1 / z once instead of off x and y
    fr_mul(x, X, zi);
    fr_mul(y, Y, zi);
    if (te) rte_to_te_x(x);  // results leave in iden3 coordinates (a product with a Montgomery-form constant keeps the form)
    u32 ox[8], oy[8];
    fr_copy(ox, x);
    fr_copy(oy, y);
    fr_canon(ox);
    fr_canon(oy);
    bool bad = status && status[p / pts_per_item] != GCP_STATUS_OK;
    if (bad) {
      fr_set_zero(ox);
      fr_set_zero(oy);
    }
    store_fr(out + p * 16, ox);
    store_fr(out + p * 16 + 8, oy);
  }
}

// ---- Ciphertext.Add / Neg, element-wise ------------------------------------------------------------------------
// a, b: n x 32 words (ciphertexts); out_xyz: n x 2 x 24 words.  One thread per point pair, 9 multiplications:
// the unified addition on two AFFINE inputs without the T output (A, B, x1 y1, x2 y2, their product, times 2d, X, Y, Z;
// D = 2 Z1 Z2 is a constant) instead of two ext_from_affine + ext_add (11), and no conversion of standard-form inputs
// (4 more): every product below is a Montgomery product of the values AS READ, so with standard-form inputs each of
// A, B, C, D carries the same factor 1 / R (C through the constant 2d R^3, D as the constant 2 / R) and X : Y : Z is the
// same projective point; normalize_kernel divides the factor out.  Same rational function as the reference's affine law
// (ciphertext.go:29-30): it never uses the curve equation, and Z = 0 exactly when an affine denominator is 0.
#define GCP_FR_TWO_MONT {0x9ffffff6u, 0x592c6838u, 0x3ec19a53u, 0x6df8ed2bu, 0xf0f28c5cu, 0xccdd46deu, 0x340fbe5eu, 0x1c14ef83u}
#define GCP_FR_TWO_OVER_R {0xdb62329cu, 0xb8b7400au, 0xc223d90fu, 0x121deb53u, 0x5d70babau, 0x904c1bc9u, 0x058aaa39u, 0x2bd7f2a3u}
__global__ void __launch_bounds__(128) ct_add_kernel(const u32* __restrict__ a, const u32* __restrict__ b, size_t n,
                                                     u32* __restrict__ out_xyz, u8* __restrict__ status, int mont, int te) {
  // One point per thread.  A per-thread loop over several points (so that the resident warps share what the instruction
  // cache holds of the ~30 KB of inlined multiplier bodies: ncu shows 2.8 no-instruction stalls per issue) was measured:
  // 1.55 against 1.44 ms per 2^23 points - the body is too large for a loop to keep it cached (the tally loop's lesson);
  // the out-of-line multiplier bodies of edwards.cuh (fr_mul2_ool / fr_mul_ool, ~10 KB of code instead of 30): 1.42 ms, the
  // argument moves cost what the fetch stalls did.  The inlined form stays.
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= 2 * n) return;
  u32 x1[8], y1[8], x2[8], y2[8];
  load_fr(x1, a + idx * 16);
  load_fr(y1, a + idx * 16 + 8);
  load_fr(x2, b + idx * 16);
  load_fr(y2, b + idx * 16 + 8);
  const bool canon = fr_is_canonical(x1) && fr_is_canonical(y1) && fr_is_canonical(x2) && fr_is_canonical(y2);
  if (te) {  // iden3 coordinates: x_RTE = x_TE * (-f), in whichever form the values are
    te_to_rte_x(x1);
    te_to_rte_x(x2);
  }
  ExtPoint p;
  {
    const u32 k_m[8] = GCP_ED_2D_MONT, k_s[8] = GCP_ED_2D_R3, d_m[8] = GCP_FR_TWO_MONT, d_s[8] = GCP_FR_TWO_OVER_R;
    u32 kc[8], d[8];
#pragma unroll
    for (int l = 0; l < 8; l++) {
      kc[l] = mont ? k_m[l] : k_s[l];
      d[l] = mont ? d_m[l] : d_s[l];
    }
    u32 t[8], u[8], v[8], w[8], A[8], B[8], c[8], e[8], f[8], g[8], h[8];
    fr_sub(t, y1, x1);
    fr_sub(u, y2, x2);
    fr_add(v, y1, x1);
    fr_add(w, y2, x2);
    fr_mul2(A, t, u, B, v, w);
    fr_mul2(t, x1, y1, u, x2, y2);
    fr_mul(c, t, u);
    fr_mul(c, c, kc);
    fr_sub(e, B, A);
    fr_sub(f, d, c);
    fr_add(g, d, c);
    fr_add(h, B, A);
    fr_mul2(p.X, e, f, p.Y, g, h);
    fr_mul(p.Z, f, g);
  }
  if (!canon) {
    fr_set_zero(p.X);
    fr_set_one(p.Y);
    fr_set_one(p.Z);
    status[idx / 2] = GCP_STATUS_NONCANONICAL;  // benign race: both halves write the same value
  }
  store_ext_xyz(out_xyz + idx * 24, p);
}

// Neg: (x, y) -> (-x, y) (ciphertext.go:37-46); pure element-wise, canonical in/out
__global__ void ct_neg_kernel(const u32* __restrict__ a, size_t n_points, u32* __restrict__ out, u8* __restrict__ status) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_points) return;
  const u32 P1[8] = GCP_P_LIMBS;
  u32 x[8], y[8], nx[8];
  load_fr(x, a + idx * 16);
  load_fr(y, a + idx * 16 + 8);
  bool canon = fr_is_canonical(x) && fr_is_canonical(y);
  if (is_zero256(x))
    fr_set_zero(nx);
  else
    sub256(nx, P1, x);  // valid in both element formats: negation commutes with the Montgomery map
  if (!canon) {
    fr_set_zero(nx);
    fr_set_zero(y);
    status[idx / 2] = GCP_STATUS_NONCANONICAL;
  }
  store_fr(out + idx * 16, nx);
  store_fr(out + idx * 16 + 8, y);
}

// IsEqual (ciphertext.go:79-87): 1 iff the four coordinates agree as field elements.  One thread per ciphertext.
__global__ void ct_is_equal_kernel(const u32* __restrict__ a, const u32* __restrict__ b, size_t n, u8* __restrict__ flags,
                                   u8* __restrict__ status) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  bool canon = true, eq = true;
#pragma unroll
  for (int c = 0; c < 4; c++) {
    u32 x[8], y[8];
    load_fr(x, a + idx * 32 + c * 8);
    load_fr(y, b + idx * 32 + c * 8);
    canon = canon && fr_is_canonical(x) && fr_is_canonical(y);
    eq = eq && eq256(x, y);  // canonical representatives (in either element format) are equal iff the elements are
  }
  flags[idx] = (canon && eq) ? 1 : 0;
  status[idx] = canon ? GCP_STATUS_OK : GCP_STATUS_NONCANONICAL;
}

// Select (ciphertext.go:90-96): z = b ? i1 : i2 per coordinate; api.Select asserts that b is boolean.
__global__ void ct_select_kernel(const u8* __restrict__ sel, const u32* __restrict__ i1, const u32* __restrict__ i2, size_t n,
                                 u32* __restrict__ out, u8* __restrict__ status) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;  // one thread per 16-byte piece: 8 per ciphertext
  if (idx >= n * 8) return;
  const size_t item = idx >> 3;
  const u8 b = sel[item];
  const uint4* src = reinterpret_cast<const uint4*>(b == 1 ? i1 : i2) + idx;
  uint4 v = (b <= 1) ? __ldg(src) : make_uint4(0, 0, 0, 0);
  reinterpret_cast<uint4*>(out)[idx] = v;
  if ((idx & 7) == 0) status[item] = (b <= 1) ? GCP_STATUS_OK : GCP_STATUS_NOT_BOOLEAN;
}

// ---- Tally: sum over ballots of ct[ballot][field], per field ----------------------------------------------------
// Each thread owns one (field, point-half) column and strides over ballots; block tree-reduction in shared
// memory; one partial (X, Y, Z, T) per (block, field, half) to `partials` [gridDim.x][n_fields*2][32 words].
constexpr int TALLY_THREADS = 128;

// Tree reduction over the rows of each column in shared memory (TALLY_THREADS x 32 words); result in row 0's acc.
// A thread's partner `h` rows further down is thread threadIdx.x + h * cols.
__device__ __forceinline__ void block_reduce_columns(ExtPoint& acc, u32* smem, int cols, int rows_per_block, int row, bool active) {
  u32* mine = smem + threadIdx.x * 32;
#pragma unroll
  for (int l = 0; l < 8; l++) {
    mine[l] = acc.X[l];
    mine[8 + l] = acc.Y[l];
    mine[16 + l] = acc.Z[l];
    mine[24 + l] = acc.T[l];
  }
  __syncthreads();
  int live = rows_per_block;
#pragma unroll 1
  while (live > 1) {
    int half = (live + 1) / 2;
    if (active && row < live / 2) {
      const u32* other = smem + (threadIdx.x + half * cols) * 32;
      ExtPoint q;
#pragma unroll
      for (int l = 0; l < 8; l++) {
        q.X[l] = other[l];
        q.Y[l] = other[8 + l];
        q.Z[l] = other[16 + l];
        q.T[l] = other[24 + l];
      }
      ext_add(acc, q);
    }
    __syncthreads();
    if (active && row < live / 2) {
#pragma unroll
      for (int l = 0; l < 8; l++) {
        mine[l] = acc.X[l];
        mine[8 + l] = acc.Y[l];
        mine[16 + l] = acc.Z[l];
        mine[24 + l] = acc.T[l];
      }
    }
    __syncthreads();
    live = half;
  }
}

__global__ void __launch_bounds__(TALLY_THREADS, 4) tally_partial_kernel(const u32* __restrict__ ct, size_t n_ballots, int n_fields,
                                                                      u32* __restrict__ partials, u32* __restrict__ bad_count,
                                                                      int mont, int te) {
  extern __shared__ u32 smem[];  // TALLY_THREADS x 32 words
  const int cols = n_fields * 2;                       // point columns per ballot
  const int rows_per_block = TALLY_THREADS / cols;     // ballots processed concurrently by one block
  const int col = threadIdx.x % cols, row = threadIdx.x / cols;
  const bool active = row < rows_per_block;
  ExtPoint acc;
  ext_identity(acc);
  u32 bad = 0;
  if (active) {
    size_t b = (size_t)blockIdx.x * rows_per_block + row;
    const size_t bstride = (size_t)gridDim.x * rows_per_block;
    // software pipeline: the next ballot's point is in flight while the current one is added
    u32 nx[8], ny[8];
    if (b < n_ballots) {
      const u32* src = ct + (b * cols + col) * 16;
      load_fr(nx, src);
      load_fr(ny, src + 8);
    }
#pragma unroll 1
    for (; b < n_ballots; b += bstride) {
      u32 xs[8], ys[8];
      fr_copy(xs, nx);
      fr_copy(ys, ny);
      if (b + bstride < n_ballots) {
        const u32* src = ct + ((b + bstride) * cols + col) * 16;
        load_fr(nx, src);
        load_fr(ny, src + 8);
      }
      if (!(fr_is_canonical(xs) && fr_is_canonical(ys))) {
        bad = 1;
        continue;
      }
      // Niels form (y - x, y + x, 2d T) of the input point, two multiplies and no conversion.  Montgomery inputs are
      // the point itself (Z = 1).  Standard-form inputs are read AS Montgomery representations: (x, y, 1, xy) then
      // stands for the projectively rescaled point (x/R : y/R : 1/R : xy/R) - the same affine point - whose sums and
      // differences need no product, whose T is x*y with the lost factor restored by the constant (2d R^2 instead of
      // 2d R), and whose Z = 1/R turns D = 2 Z1 Z2 into one Montgomery reduction of 2 Z1 (ext_add_niels, z_over_r).
      // 9 (Montgomery) / 9.5 (standard) multiplies per point instead of 11 with the inputs converted first.
      if (te) te_to_rte_x(xs);  // iden3 coordinates: x * (-f) in either representation (the constant carries the R)
      const u32 c_t_std[8] = GCP_ED_2D_R2, c_t_mont[8] = GCP_ED_2D_MONT;
      u32 c_t[8];
#pragma unroll
      for (int l = 0; l < 8; l++) c_t[l] = mont ? c_t_mont[l] : c_t_std[l];
      NielsPoint n;
      u32 t[8];
      fr_sub(n.ymx, ys, xs);
      fr_add(n.ypx, ys, xs);
      fr_mul(t, xs, ys);
      fr_mul(n.t2d, t, c_t);
      ext_add_niels<true>(acc, n, !mont);
    }
  }
  if (bad) atomicAdd(bad_count + col / 2, 1u);
  block_reduce_columns(acc, smem, cols, rows_per_block, row, active);
  if (active && row == 0) {
    u32* o = partials + ((size_t)blockIdx.x * cols + col) * 32;
    store_fr(o, acc.X);
    store_fr(o + 8, acc.Y);
    store_fr(o + 16, acc.Z);
    store_fr(o + 24, acc.T);
  }
}

// Fused Encrypt + tally (the ciphertexts are never materialised): thread (row, col) strides over ballots and adds
// [k]G into the C1 column or [k]PK + [m]G into the C2 column of its field; same reduction as tally_partial_kernel.
// ks / ms: n_ballots x n_fields scalars.  bad_count[f] counts non-canonical scalars of field f.
// mask (optional): n_ballots bytes, a ballot is summed only where mask[b] != 0.
template <int M_WORDS>
__global__ void __launch_bounds__(TALLY_THREADS, 4) encrypt_tally_partial_kernel(const u32* __restrict__ tabG, const u32* __restrict__ tabPK,
                                                                              const u32* __restrict__ ks, const u32* __restrict__ ms,
                                                                              const u8* __restrict__ mask, size_t n_ballots, int n_fields,
                                                                              u32* __restrict__ partials, u32* __restrict__ bad_count, int mont) {
  extern __shared__ u32 smem[];
  // M_WORDS: 8 = the messages are field elements like k; 2 = little-endian uint64 integers (GCP_MSG_U64).  A template
  // parameter: as a run-time argument it cost the field-element form 7 % (the kernel sits at its 128-register cap)
  // thread layout: the first half of the block (whole warps) computes C1 columns, the second half C2 columns, so a
  // warp never mixes the two branches; inside a half, thread = row * n_fields + field
  constexpr int HALF_THREADS = TALLY_THREADS / 2;
  const int cols = n_fields * 2;
  const int half = threadIdx.x / HALF_THREADS;
  const int th = threadIdx.x % HALF_THREADS;
  const int rows_per_block = HALF_THREADS / n_fields;
  const int field = th % n_fields, row = th / n_fields;
  const bool active = row < rows_per_block;
  ExtPoint acc;
  ext_identity(acc);
  u32 bad = 0;
  if (active) {
    size_t b = (size_t)blockIdx.x * rows_per_block + row;
    size_t bstride = (size_t)gridDim.x * rows_per_block;
#pragma unroll 1
    for (; b < n_ballots; b += bstride) {
      if (mask && mask[b] == 0) continue;  // ballot not admitted (e.g. census proof flag 0)
      const u32* kp = ks + (b * n_fields + field) * 8;
      const u32* mp = ms + (b * n_fields + field) * (size_t)M_WORDS;
      {  // both scalars are checked before anything is added; each is loaded again by the pass that multiplies by it (an
         // L1 hit), so that neither is live across the other's window loop (the kernel sits at its 128-register cap)
        u32 t[8];
        load_fr(t, kp);
        bool canon = fr_is_canonical(t);
        if constexpr (M_WORDS == 8) {
          load_fr(t, mp);
          canon = canon && fr_is_canonical(t);
        }
        if (!canon) {
          bad = 1;
          continue;
        }
      }
      const int n_pass = half ? 2 : 1;       // one inlined copy of the window loop (see encrypt_shared_kernel)
#pragma unroll 1
      for (int pass = 0; pass < n_pass; pass++) {
        const u32* tab = (half && pass == 0) ? tabPK : tabG;
        u32 sc[8];
        bool canon = true;
        if constexpr (M_WORDS == 8) {
          load_scalar(sc, canon, pass == 0 ? kp : mp, mont);
        } else {
          if (pass == 0) {
            load_scalar(sc, canon, kp, mont);
          } else {
            const uint2 v = __ldg(reinterpret_cast<const uint2*>(mp));
            fr_set_zero(sc);
            sc[0] = v.x;
            sc[1] = v.y;
          }
        }
        fixed_base_accumulate(acc, sc, tab);
      }
    }
  }
  if (bad) atomicAdd(bad_count + field, 1u);
  block_reduce_columns(acc, smem, n_fields, rows_per_block, row, active);
  if (active && row == 0) {
    u32* o = partials + ((size_t)blockIdx.x * cols + field * 2 + half) * 32;
    store_fr(o, acc.X);
    store_fr(o + 8, acc.Y);
    store_fr(o + 16, acc.Z);
    store_fr(o + 24, acc.T);
  }
}

// Second stage: one block per column; threads stride over the per-block partials, then a shared-memory tree.
// Writes (X, Y, Z) for normalize_kernel.
__global__ void __launch_bounds__(TALLY_THREADS) tally_final_kernel(const u32* __restrict__ partials, int n_blocks, int cols,
                                                                    u32* __restrict__ out_xyz, const u32* __restrict__ bad_count,
                                                                    u8* __restrict__ status) {
  __shared__ u32 smem[TALLY_THREADS * 32];
  const int col = blockIdx.x;
  ExtPoint acc;
  ext_identity(acc);
#pragma unroll 1
  for (int b = threadIdx.x; b < n_blocks; b += TALLY_THREADS) {
    const u32* s = partials + ((size_t)b * cols + col) * 32;
    ExtPoint q;
    load_fr(q.X, s);
    load_fr(q.Y, s + 8);
    load_fr(q.Z, s + 16);
    load_fr(q.T, s + 24);
    ext_add(acc, q);
  }
  u32* mine = smem + threadIdx.x * 32;
  int live = min(n_blocks, TALLY_THREADS);
  if (live < 1) live = 1;
#pragma unroll 1
  while (true) {
#pragma unroll
    for (int l = 0; l < 8; l++) {
      mine[l] = acc.X[l];
      mine[8 + l] = acc.Y[l];
      mine[16 + l] = acc.Z[l];
      mine[24 + l] = acc.T[l];
    }
    __syncthreads();
    if (live <= 1) break;
    int half = (live + 1) / 2;
    if ((int)threadIdx.x < live / 2) {
      const u32* other = smem + (threadIdx.x + half) * 32;
      ExtPoint q;
#pragma unroll
      for (int l = 0; l < 8; l++) {
        q.X[l] = other[l];
        q.Y[l] = other[8 + l];
        q.Z[l] = other[16 + l];
        q.T[l] = other[24 + l];
      }
      ext_add(acc, q);
    }
    __syncthreads();
    live = half;
  }
  if (threadIdx.x == 0) {
    store_ext_xyz(out_xyz + (size_t)col * 24, acc);
    if ((col & 1) == 0) status[col / 2] = bad_count[col / 2] ? GCP_STATUS_NONCANONICAL : GCP_STATUS_OK;
  }
}

}  // namespace gcp
