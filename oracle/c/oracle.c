/* oracle.c — CPU restatement (plain C, 4x64-bit Montgomery limbs, the representation gnark-crypto's
 * fr.Element uses) of the reference's hot path.  TEST INFRASTRUCTURE ONLY: linked/loaded by tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs, never by the product.
 *
 * Restates (file:line under /root/reference):
 *   Poseidon      hash/native/bn254/poseidon/poseidon.go:116-233 (== the plain-field referenceHash of
 *                 hash/emulated/bn254/poseidon/poseidon_test.go:112-273), MultiHash :54-91
 *   SMT verifier  tree/smt/verifier.go:102-242, lev_ins.go:43-77, verifier_sm.go:5-14,
 *                 verifier_level.go:8-17, hash.go:10-27, utils.go:11-56   (literal: all n levels hashed)
 *   ElGamal       elgamal/encrypt.go:42-64, elgamal/mul.go:26-166, elgamal/ciphertext.go:24-46
 *                 over the un-vendored gnark-crypto bn254 twisted Edwards curve (a = -1)
 *   Address       ecc/secp256k1/ecdsa/address.go:14-40 (legacy Keccak-256)
 * It is itself checked against the Python oracle (oracle/ *.py) (which is pinned to the golden vectors) in tests/test_oracle_c.py.
 * Build: make -C oracle/c   ->  oracle/c/liboracle.so
 */
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned __int128 u128;
typedef uint64_t u64;
typedef struct { u64 v[4]; } fe; /* Montgomery form, canonical (< r) */

static const u64 P[4] = {0x43e1f593f0000001ULL, 0x2833e84879b97091ULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL};
static const u64 NP = 0xc2e1f593efffffffULL; /* -r^-1 mod 2^64 */
static const fe R2 = {{0x1bb8e645ae216da7ULL, 0x53fe3ab1e35c59e3ULL, 0x8c49833d53bb8085ULL, 0x0216d0b17f4e44a5ULL}};
static const fe ONE = {{0xac96341c4ffffffbULL, 0x36fc76959f60cd29ULL, 0x666ea36f7879462eULL, 0x0e0a77c19a07df2fULL}};
static const fe ZERO = {{0, 0, 0, 0}};

static inline int geq_p(const u64 a[4]) {
  for (int i = 3; i >= 0; i--) {
    if (a[i] > P[i]) return 1;
    if (a[i] < P[i]) return 0;
  }
  return 1;
}
static inline void sub_p(u64 a[4]) {
  u128 b = 0;
  for (int i = 0; i < 4; i++) {
    u128 d = (u128)a[i] - P[i] - (u64)b;
    a[i] = (u64)d;
    b = (d >> 64) & 1;
  }
}
static inline void fe_add(fe* r, const fe* a, const fe* b) {
  u128 c = 0;
  for (int i = 0; i < 4; i++) {
    c += (u128)a->v[i] + b->v[i];
    r->v[i] = (u64)c;
    c >>= 64;
  }
  if (c || geq_p(r->v)) sub_p(r->v);
}
static inline void fe_sub(fe* r, const fe* a, const fe* b) {
  u64 t[4];
  u128 br = 0;
  for (int i = 0; i < 4; i++) {
    u128 d = (u128)a->v[i] - b->v[i] - (u64)br;
    t[i] = (u64)d;
    br = (d >> 64) & 1;
  }
  if (br) {
    u128 c = 0;
    for (int i = 0; i < 4; i++) {
      c += (u128)t[i] + P[i];
      t[i] = (u64)c;
      c >>= 64;
    }
  }
  memcpy(r->v, t, 32);
}
/* CIOS Montgomery multiplication */
static inline void fe_mul(fe* r, const fe* a, const fe* b) {
  u64 t[6] = {0, 0, 0, 0, 0, 0};
  for (int i = 0; i < 4; i++) {
    u128 c = 0;
    for (int j = 0; j < 4; j++) {
      c += (u128)a->v[j] * b->v[i] + t[j];
      t[j] = (u64)c;
      c >>= 64;
    }
    c += t[4];
    t[4] = (u64)c;
    t[5] = (u64)(c >> 64);
    u64 m = t[0] * NP;
    c = ((u128)m * P[0] + t[0]) >> 64;
    for (int j = 1; j < 4; j++) {
      c += (u128)m * P[j] + t[j];
      t[j - 1] = (u64)c;
      c >>= 64;
    }
    c += t[4];
    t[3] = (u64)c;
    t[4] = t[5] + (u64)(c >> 64);
  }
  if (t[4] || geq_p(t)) sub_p(t);
  memcpy(r->v, t, 32);
}
static inline int fe_is_zero(const fe* a) { return (a->v[0] | a->v[1] | a->v[2] | a->v[3]) == 0; }
static inline int fe_eq(const fe* a, const fe* b) { return memcmp(a, b, 32) == 0; }
static inline void fe_neg(fe* r, const fe* a) { fe_sub(r, &ZERO, a); }
/* bytes (LE canonical integer) <-> Montgomery; returns 0 when the integer is >= r */
static inline int fe_from_bytes(fe* r, const uint8_t* b) {
  fe t;
  memcpy(t.v, b, 32);
  if (geq_p(t.v)) return 0;
  fe_mul(r, &t, &R2);
  return 1;
}
static inline void fe_to_bytes(uint8_t* b, const fe* a) {
  fe one_std = {{1, 0, 0, 0}}, t;
  fe_mul(&t, a, &one_std);
  memcpy(b, t.v, 32);
}
static void fe_from_u64(fe* r, u64 x) {
  fe t = {{x, 0, 0, 0}};
  fe_mul(r, &t, &R2);
}
static void fe_pow(fe* r, const fe* a, const u64 e[4]) {
  fe acc = ONE, base = *a;
  for (int i = 0; i < 256; i++) {
    if ((e[i / 64] >> (i % 64)) & 1) fe_mul(&acc, &acc, &base);
    fe_mul(&base, &base, &base);
  }
  *r = acc;
}
static void fe_inv(fe* r, const fe* a) { /* a^(r-2); 0 -> 0 */
  u64 e[4] = {P[0] - 2, P[1], P[2], P[3]};
  fe_pow(r, a, e);
}

/* ------------------------------------------------------------------------------------------------ */
/* Poseidon tables                                                                                  */
/* ------------------------------------------------------------------------------------------------ */
typedef struct { int t, rp; fe *C, *S, *M, *Pm; } ptab;
static ptab TAB[18];
static fe* ALL = NULL;
static int INITED = 0;

int oracle_init(const char* blob_path) {
  if (INITED) return 0;
  FILE* f = fopen(blob_path, "rb");
  if (!f) return -1;
  fseek(f, 0, SEEK_END);
  long sz = ftell(f);
  fseek(f, 0, SEEK_SET);
  uint8_t* raw = (uint8_t*)malloc(sz);
  if (fread(raw, 1, sz, f) != (size_t)sz) { fclose(f); free(raw); return -2; }
  fclose(f);
  uint32_t hdr[4];
  memcpy(hdr, raw, 16);
  if (hdr[0] != 0x32425350u || hdr[1] != 1 || hdr[2] != 16) { free(raw); return -3; }
  size_t base = 16 + 16 * 40, n = (sz - base) / 32;
  ALL = (fe*)malloc(n * sizeof(fe));
  for (size_t i = 0; i < n; i++)
    if (!fe_from_bytes(&ALL[i], raw + base + 32 * i)) { free(raw); return -4; }
  for (int i = 0; i < 16; i++) {
    uint32_t d[10];
    memcpy(d, raw + 16 + 40 * i, 40);
    ptab* p = &TAB[d[0]];
    p->t = d[0]; p->rp = d[1];
    p->C = ALL + d[2]; p->S = ALL + d[4]; p->M = ALL + d[6]; p->Pm = ALL + d[8];
  }
  free(raw);
  INITED = 1;
  return 0;
}

static inline void sigma(fe* x) { /* poseidon.go:199-203 */
  fe x2, x4;
  fe_mul(&x2, x, x);
  fe_mul(&x4, &x2, &x2);
  fe_mul(x, &x4, x);
}
static void mix(fe* out, const fe* in, const fe* m, int t) { /* poseidon.go:213-224: out[i] = sum_j m[j][i] in[j] */
  for (int i = 0; i < t; i++) {
    fe acc = ZERO, pr;
    for (int j = 0; j < t; j++) {
      fe_mul(&pr, &m[j * t + i], &in[j]);
      fe_add(&acc, &acc, &pr);
    }
    out[i] = acc;
  }
}
/* poseidon.go:116-183; in: n_in Montgomery elements */
static void poseidon_sum(fe* out, const fe* in, int n_in) {
  int t = n_in + 1;
  const ptab* p = &TAB[t];
  fe st[17], nw[17], pr;
  st[0] = ZERO;
  for (int j = 1; j < t; j++) st[j] = in[j - 1];
  for (int j = 0; j < t; j++) fe_add(&st[j], &st[j], &p->C[j]);
  for (int r = 0; r < 3; r++) {
    for (int j = 0; j < t; j++) { sigma(&st[j]); fe_add(&st[j], &st[j], &p->C[(r + 1) * t + j]); }
    mix(nw, st, p->M, t);
    memcpy(st, nw, sizeof(fe) * t);
  }
  for (int j = 0; j < t; j++) { sigma(&st[j]); fe_add(&st[j], &st[j], &p->C[4 * t + j]); }
  mix(nw, st, p->Pm, t);
  memcpy(st, nw, sizeof(fe) * t);
  for (int r = 0; r < p->rp; r++) {
    sigma(&st[0]);
    fe_add(&st[0], &st[0], &p->C[5 * t + r]);
    const fe* s = p->S + (2 * t - 1) * r;
    fe n0 = ZERO;
    for (int j = 0; j < t; j++) { fe_mul(&pr, &s[j], &st[j]); fe_add(&n0, &n0, &pr); }
    for (int k = 1; k < t; k++) { fe_mul(&pr, &st[0], &s[t + k - 1]); fe_add(&st[k], &st[k], &pr); }
    st[0] = n0;
  }
  for (int r = 0; r < 3; r++) {
    for (int j = 0; j < t; j++) { sigma(&st[j]); fe_add(&st[j], &st[j], &p->C[5 * t + p->rp + r * t + j]); }
    mix(nw, st, p->M, t);
    memcpy(st, nw, sizeof(fe) * t);
  }
  fe acc = ZERO;
  for (int j = 0; j < t; j++) { sigma(&st[j]); fe_mul(&pr, &p->M[j * t], &st[j]); fe_add(&acc, &acc, &pr); }
  *out = acc;
}
/* poseidon.go:54-91 */
static void poseidon_multihash_fe(fe* out, const fe* in, int len) {
  if (len <= 16) { poseidon_sum(out, in, len); return; }
  int nch = (len + 15) / 16;
  fe* h = (fe*)malloc(sizeof(fe) * nch);
  for (int c = 0; c < nch; c++) {
    int a = len - 16 * c; if (a > 16) a = 16;
    poseidon_sum(&h[c], in + 16 * c, a);
  }
  poseidon_multihash_fe(out, h, nch);
  free(h);
}

/* ------------------------------------------------------------------------------------------------ */
/* thread pool: static index ranges                                                                  */
/* ------------------------------------------------------------------------------------------------ */
typedef void (*range_fn)(void* arg, size_t lo, size_t hi);
typedef struct { range_fn fn; void* arg; size_t lo, hi; } job;
static void* job_main(void* p) { job* j = (job*)p; j->fn(j->arg, j->lo, j->hi); return NULL; }
static void parallel_for(range_fn fn, void* arg, size_t n, int threads) {
  if (threads <= 1 || n < 2) { fn(arg, 0, n); return; }
  if ((size_t)threads > n) threads = (int)n;
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * threads);
  job* jb = (job*)malloc(sizeof(job) * threads);
  for (int i = 0; i < threads; i++) {
    jb[i].fn = fn; jb[i].arg = arg; jb[i].lo = n * i / threads; jb[i].hi = n * (i + 1) / threads;
    pthread_create(&th[i], NULL, job_main, &jb[i]);
  }
  for (int i = 0; i < threads; i++) pthread_join(th[i], NULL);
  free(th); free(jb);
}

/* ---- batch Poseidon ------------------------------------------------------------------------------ */
typedef struct { const uint8_t* in; int len; uint8_t* out; uint8_t* status; int multi; } pos_args;
static void pos_range(void* vp, size_t lo, size_t hi) {
  pos_args* a = (pos_args*)vp;
  fe* buf = (fe*)malloc(sizeof(fe) * a->len);
  for (size_t i = lo; i < hi; i++) {
    int ok = 1;
    for (int j = 0; j < a->len; j++) ok &= fe_from_bytes(&buf[j], a->in + (i * a->len + j) * 32);
    fe h = ZERO;
    if (ok) { if (a->multi) poseidon_multihash_fe(&h, buf, a->len); else poseidon_sum(&h, buf, a->len); }
    fe_to_bytes(a->out + i * 32, &h);
    if (a->status) a->status[i] = ok ? 0 : 1;
  }
  free(buf);
}
int oracle_poseidon_hash(const uint8_t* in, int arity, size_t n, uint8_t* out, uint8_t* status, int threads) {
  if (!INITED || arity < 1 || arity > 16) return -1;
  pos_args a = {in, arity, out, status, 0};
  parallel_for(pos_range, &a, n, threads);
  return 0;
}
int oracle_poseidon_multihash(const uint8_t* in, int len, size_t n, uint8_t* out, uint8_t* status, int threads) {
  if (!INITED || len < 1 || len > 4096) return -1;
  pos_args a = {in, len, out, status, 1};
  parallel_for(pos_range, &a, n, threads);
  return 0;
}

/* ---- SMT verifier, literal field-arithmetic state machine ---------------------------------------- */
typedef struct {
  int n_levels; const uint8_t *roots; int shared_root; const uint8_t* sib; const uint8_t *okeys, *ovals, *is0;
  const uint8_t *keys, *vals, *fnc, *en; uint8_t *flags, *status, *oroots; int literal;
} smt_args;
static void hash1(fe* out, const fe* k, const fe* v) { fe in[3] = {*k, *v, ONE}; poseidon_sum(out, in, 3); }
static void hash2(fe* out, const fe* l, const fe* r) { fe in[2] = {*l, *r}; poseidon_sum(out, in, 2); }

static void smt_one(const smt_args* a, size_t i, fe* sib, fe* lev_ins, fe (*st)[5], fe* levels) {
  const int n = a->n_levels;
  uint8_t flag = 0, status = 0;
  fe root_out = ZERO;
  fe root, key, val, okey, oval;
  int ok = 1;
  ok &= fe_from_bytes(&root, a->roots + (a->shared_root ? 0 : i * 32));
  ok &= fe_from_bytes(&key, a->keys + i * 32);
  ok &= fe_from_bytes(&val, a->vals + i * 32);
  if (a->okeys) { ok &= fe_from_bytes(&okey, a->okeys + i * 32); ok &= fe_from_bytes(&oval, a->ovals + i * 32); }
  else { okey = key; oval = val; }
  for (int l = 0; l < n; l++) ok &= fe_from_bytes(&sib[l], a->sib + (i * n + l) * 32);
  unsigned en = a->en ? a->en[i] : 1, fn = a->fnc ? a->fnc[i] : 0, i0 = a->is0 ? a->is0[i] : 0;
  if (!ok) { status = 1; goto done; }
  if (en > 1 || fn > 1 || i0 > 1) { status = 3; goto done; }
  { /* lowBits: key < 2^n (utils.go:11-13) */
    const uint8_t* kb = a->keys + i * 32;
    for (int b = n; b < 256; b++) if ((kb[b / 8] >> (b % 8)) & 1) { status = 2; goto done; }
  }
  {
    fe enabled, fnc, is0, one_minus_is0, h1old, h1new;
    fe_from_u64(&enabled, en); fe_from_u64(&fnc, fn); fe_from_u64(&is0, i0);
    fe_sub(&one_minus_is0, &ONE, &is0);
    hash1(&h1old, &okey, &oval);                      /* verifier.go:112 */
    hash1(&h1new, &key, &val);                        /* verifier.go:113 */
    /* LevInsFlag, lev_ins.go:43-77 */
    fe done_acc, tmp, tmp2;
    int leaf_zero = fe_is_zero(&sib[n - 1]);
    fe nz;
    if (fe_is_zero(&sib[n - 2])) nz = ZERO; else nz = ONE;
    lev_ins[n - 1] = nz;
    done_acc = lev_ins[n - 1];
    for (int l = n - 2; l > 0; l--) {
      fe_sub(&tmp, &ONE, &done_acc);
      if (fe_is_zero(&sib[l - 1])) tmp2 = ZERO; else tmp2 = ONE;
      fe_mul(&lev_ins[l], &tmp, &tmp2);
      fe_add(&done_acc, &lev_ins[l], &done_acc);
    }
    fe_sub(&lev_ins[0], &ONE, &done_acc);
    fe sum = ZERO;
    for (int l = 0; l < n; l++) fe_add(&sum, &sum, &lev_ins[l]);
    int one_hot = fe_eq(&sum, &ONE);
    int flag_lev_ins = en ? (leaf_zero && one_hot) : 1;
    /* VerifierSM chain, verifier.go:194-202, verifier_sm.go:5-14: st = {top, i0, iold, inew, na} */
    fe prev[5];
    prev[0] = enabled; prev[1] = ZERO; prev[2] = ZERO; prev[3] = ZERO; fe_sub(&prev[4], &ONE, &enabled);
    for (int l = 0; l < n; l++) {
      fe aux1, aux2;
      fe_mul(&aux1, &prev[0], &lev_ins[l]);
      fe_mul(&aux2, &aux1, &fnc);
      fe_sub(&st[l][0], &prev[0], &aux1);
      fe_sub(&st[l][3], &aux1, &aux2);
      fe_mul(&st[l][2], &aux2, &one_minus_is0);
      fe_mul(&st[l][1], &aux1, &is0);
      fe_add(&tmp, &prev[4], &prev[3]); fe_add(&tmp, &tmp, &prev[2]); fe_add(&st[l][4], &tmp, &prev[1]);
      memcpy(prev, st[l], sizeof(prev));
    }
    fe_add(&tmp, &st[n - 1][4], &st[n - 1][2]); fe_add(&tmp, &tmp, &st[n - 1][3]); fe_add(&tmp, &tmp, &st[n - 1][1]);
    int flag_states = fe_eq(&tmp, &ONE);
    /* levels, verifier.go:211-220, verifier_level.go:8-17 */
    const uint8_t* kb = a->keys + i * 32;
    for (int l = n - 1; l >= 0; l--) {
      fe child = (l < n - 1) ? levels[l + 1] : ZERO;
      int bit = (kb[l / 8] >> (l % 8)) & 1;
      fe ph = ZERO;
      if (a->literal || !fe_is_zero(&st[l][0])) {     /* the gadget always hashes; stTop==0 masks the result */
        if (bit) hash2(&ph, &sib[l], &child); else hash2(&ph, &child, &sib[l]);
      }
      fe t1, t2, t3;
      fe_mul(&t1, &ph, &st[l][0]); fe_mul(&t2, &h1old, &st[l][2]); fe_mul(&t3, &h1new, &st[l][3]);
      fe_add(&t1, &t1, &t2); fe_add(&levels[l], &t1, &t3);
    }
    int keys_eq = fe_eq(&okey, &key);
    int key_reuse = (fn && !i0 && keys_eq && en);
    int flag_root = en ? fe_eq(&levels[0], &root) : 1;
    flag = (uint8_t)(flag_states && !key_reuse && flag_root && flag_lev_ins);
    root_out = levels[0];
  }
done:
  a->flags[i] = status ? 0 : flag;
  a->status[i] = status;
  if (a->oroots) fe_to_bytes(a->oroots + i * 32, &root_out);
}
static void smt_range(void* vp, size_t lo, size_t hi) {
  smt_args* a = (smt_args*)vp;
  int n = a->n_levels;
  fe* sib = (fe*)malloc(sizeof(fe) * n);
  fe* li = (fe*)malloc(sizeof(fe) * n);
  fe(*st)[5] = (fe(*)[5])malloc(sizeof(fe) * 5 * n);
  fe* lv = (fe*)malloc(sizeof(fe) * n);
  for (size_t i = lo; i < hi; i++) smt_one(a, i, sib, li, st, lv);
  free(sib); free(li); free(st); free(lv);
}
int oracle_smt_verify(int n_levels, size_t n, const uint8_t* roots, int shared_root, const uint8_t* siblings,
                      const uint8_t* old_keys, const uint8_t* old_values, const uint8_t* is_old0, const uint8_t* keys,
                      const uint8_t* values, const uint8_t* fnc, const uint8_t* enabled, uint8_t* flags,
                      uint8_t* status, uint8_t* out_roots, int literal, int threads) {
  if (!INITED || n_levels < 2 || n_levels > 253) return -1;
  smt_args a = {n_levels, roots, shared_root, siblings, old_keys, old_values, is_old0, keys, values, fnc, enabled,
                flags, status, out_roots, literal};
  parallel_for(smt_range, &a, n, threads);
  return 0;
}

/* ---- twisted Edwards (a = -1) / ElGamal ------------------------------------------------------------ */
typedef struct { fe x, y; } aff;
static fe ED_D; static aff ED_G; static int ED_INIT = 0;
static aff FB_TABLE[64][16];
static const uint8_t D_BYTES[32] = {0x8e,0xeb,0xd7,0xf4,0x8c,0xca,0x75,0xd0,0x67,0xc8,0xb7,0xeb,0x59,0x29,0x9b,0x03,
                                    0xfc,0x11,0xfd,0x99,0xd7,0x72,0xf0,0x3d,0x69,0x89,0x21,0x5f,0xf1,0x90,0xee,0x1a};
static const uint8_t GX_BYTES[32] = {0x9f,0x3f,0x55,0xfe,0x5a,0x19,0xf9,0xf1,0x77,0xa2,0xf2,0xe6,0x9a,0x74,0x7c,0x37,
                                    0x4c,0xe9,0x99,0xc1,0xa4,0xb7,0x4e,0x8a,0x35,0x9d,0xe1,0x6c,0x83,0xff,0x61,0x15};
static const uint8_t GY_BYTES[32] = {0x8b,0x7d,0x2d,0x87,0x7a,0x25,0x3c,0x4b,0x77,0x33,0xe1,0xb9,0x1f,0x05,0xe0,0xfc,
                                    0xed,0xf9,0x6b,0xd1,0x1c,0x2e,0x57,0x25,0x49,0xb2,0xa0,0xf7,0x03,0x72,0x79,0x25};
/* complete affine addition; returns 0 on a zero denominator */
static int ed_add(aff* r, const aff* p, const aff* q) {
  fe x1x2, y1y2, k, dx, dy, nx, ny, t1, t2, ix, iy;
  fe_mul(&x1x2, &p->x, &q->x); fe_mul(&y1y2, &p->y, &q->y);
  fe_mul(&k, &x1x2, &y1y2); fe_mul(&k, &k, &ED_D);
  fe_add(&dx, &ONE, &k); fe_sub(&dy, &ONE, &k);
  if (fe_is_zero(&dx) || fe_is_zero(&dy)) return 0;
  fe_mul(&t1, &p->x, &q->y); fe_mul(&t2, &p->y, &q->x); fe_add(&nx, &t1, &t2);
  fe_add(&ny, &y1y2, &x1x2); /* y1y2 - a x1x2, a = -1 */
  fe_inv(&ix, &dx); fe_inv(&iy, &dy);
  fe_mul(&r->x, &nx, &ix); fe_mul(&r->y, &ny, &iy);
  return 1;
}
static int ed_on_curve(const aff* p) { /* -x^2 + y^2 == 1 + d x^2 y^2 */
  fe x2, y2, l, r;
  fe_mul(&x2, &p->x, &p->x); fe_mul(&y2, &p->y, &p->y);
  fe_sub(&l, &y2, &x2); fe_mul(&r, &x2, &y2); fe_mul(&r, &r, &ED_D); fe_add(&r, &r, &ONE);
  return fe_eq(&l, &r);
}
/* projective version of the same law for the long double-and-add chains (one inversion at the end) */
typedef struct { fe x, y, z; } prj;
static void prj_add(prj* r, const prj* p, const prj* q) {
  fe A, B, C, Dd, E, F, G, t1, t2, t3;
  fe_mul(&A, &p->z, &q->z); fe_mul(&B, &A, &A); fe_mul(&C, &p->x, &q->x); fe_mul(&Dd, &p->y, &q->y);
  fe_mul(&E, &C, &Dd); fe_mul(&E, &E, &ED_D); fe_sub(&F, &B, &E); fe_add(&G, &B, &E);
  fe_add(&t1, &p->x, &p->y); fe_add(&t2, &q->x, &q->y); fe_mul(&t3, &t1, &t2); fe_sub(&t3, &t3, &C); fe_sub(&t3, &t3, &Dd);
  fe_mul(&t1, &A, &F); fe_mul(&r->x, &t1, &t3);
  fe_add(&t2, &Dd, &C); fe_mul(&t1, &A, &G); fe_mul(&r->y, &t1, &t2);
  fe_mul(&r->z, &F, &G);
}
static int ed_scalar_mul(aff* r, const aff* p, const uint8_t* s_le) { /* [s]P, s a 256-bit LE integer */
  prj acc = {ZERO, ONE, ONE}, base = {p->x, p->y, ONE};
  for (int i = 0; i < 256; i++) {
    if ((s_le[i / 8] >> (i % 8)) & 1) prj_add(&acc, &acc, &base);
    prj_add(&base, &base, &base);
  }
  if (fe_is_zero(&acc.z)) return 0;
  fe zi; fe_inv(&zi, &acc.z);
  fe_mul(&r->x, &acc.x, &zi); fe_mul(&r->y, &acc.y, &zi);
  return 1;
}
static void ed_init(void) {
  if (ED_INIT) return;
  fe_from_bytes(&ED_D, D_BYTES); fe_from_bytes(&ED_G.x, GX_BYTES); fe_from_bytes(&ED_G.y, GY_BYTES);
  /* mul.go:26-72: table[i][j] = [j * 2^(4i)] G ; entry 0 = identity */
  aff base = ED_G;
  for (int i = 0; i < 64; i++) {
    int entries = (i == 63) ? 4 : 16;
    FB_TABLE[i][0].x = ZERO; FB_TABLE[i][0].y = ONE;
    for (int j = 1; j < entries; j++) ed_add(&FB_TABLE[i][j], &FB_TABLE[i][j - 1], &base);
    aff b2, b4, b8;
    ed_add(&b2, &base, &base); ed_add(&b4, &b2, &b2); ed_add(&b8, &b4, &b4); ed_add(&base, &b8, &b8);
  }
  ED_INIT = 1;
}
/* mul.go:76-166 */
static int fixed_base_mul(aff* res, const uint8_t* s_le) {
  for (int i = 0; i < 64; i++) {
    int nib = (s_le[i / 2] >> ((i & 1) * 4)) & 0xF;
    if (i == 63) nib &= 3;
    if (i == 0) { *res = FB_TABLE[0][nib]; continue; }
    if (nib) { aff t; if (!ed_add(&t, res, &FB_TABLE[i][nib])) return 0; *res = t; }
  }
  return 1;
}
static int pt_from_bytes(aff* p, const uint8_t* b) { return fe_from_bytes(&p->x, b) & fe_from_bytes(&p->y, b + 32); }
static void pt_to_bytes(uint8_t* b, const aff* p) { fe_to_bytes(b, &p->x); fe_to_bytes(b + 32, &p->y); }

typedef struct { const uint8_t* pk; int pk_per_item; const uint8_t *k, *m; uint8_t* out; uint8_t* status; } enc_args;
static void enc_range(void* vp, size_t lo, size_t hi) {
  enc_args* a = (enc_args*)vp;
  for (size_t i = lo; i < hi; i++) {
    aff pk, c1, c2, s, mp;
    fe tk, tm;
    uint8_t st = 0;
    memset(a->out + i * 128, 0, 128);
    if (!pt_from_bytes(&pk, a->pk + (a->pk_per_item ? i * 64 : 0)) || !fe_from_bytes(&tk, a->k + i * 32) ||
        !fe_from_bytes(&tm, a->m + i * 32)) st = 1;
    else if (!ed_on_curve(&pk)) st = 4;                                   /* encrypt.go:49 */
    else if (!fixed_base_mul(&c1, a->k + i * 32) ||                       /* encrypt.go:52 */
             !ed_scalar_mul(&s, &pk, a->k + i * 32) ||                    /* encrypt.go:55 */
             !fixed_base_mul(&mp, a->m + i * 32) ||                       /* encrypt.go:58 */
             !ed_add(&c2, &mp, &s)) st = 5;                               /* encrypt.go:61 */
    if (!st) { pt_to_bytes(a->out + i * 128, &c1); pt_to_bytes(a->out + i * 128 + 64, &c2); }
    if (a->status) a->status[i] = st;
  }
}
int oracle_elgamal_encrypt(const uint8_t* pk, int pk_per_item, const uint8_t* k, const uint8_t* m, size_t n,
                           uint8_t* out, uint8_t* status, int threads) {
  ed_init();
  enc_args a = {pk, pk_per_item, k, m, out, status};
  parallel_for(enc_range, &a, n, threads);
  return 0;
}
typedef struct { const uint8_t *a, *b; uint8_t *out, *status; } add_args;
static void add_range(void* vp, size_t lo, size_t hi) {
  add_args* a = (add_args*)vp;
  for (size_t i = lo; i < hi; i++) {
    uint8_t st = 0;
    memset(a->out + i * 128, 0, 128);
    for (int h = 0; h < 2 && !st; h++) {
      aff p, q, r;
      if (!pt_from_bytes(&p, a->a + i * 128 + 64 * h) || !pt_from_bytes(&q, a->b + i * 128 + 64 * h)) { st = 1; break; }
      if (!ed_add(&r, &p, &q)) { st = 5; break; }                          /* ciphertext.go:29-30 */
      pt_to_bytes(a->out + i * 128 + 64 * h, &r);
    }
    if (st) memset(a->out + i * 128, 0, 128);
    if (a->status) a->status[i] = st;
  }
}
int oracle_elgamal_add(const uint8_t* x, const uint8_t* y, size_t n, uint8_t* out, uint8_t* status, int threads) {
  ed_init();
  add_args a = {x, y, out, status};
  parallel_for(add_range, &a, n, threads);
  return 0;
}
/* left fold of Ciphertext.Add from NewCiphertext (ciphertext.go:16-32) per field: ct[ballot][field] */
typedef struct { const uint8_t* ct; size_t n_ballots; int n_fields; uint8_t* out; uint8_t* status; } tally_args;
static void tally_range(void* vp, size_t lo, size_t hi) {
  tally_args* a = (tally_args*)vp;
  for (size_t f = lo; f < hi; f++) {
    prj acc[2] = {{ZERO, ONE, ONE}, {ZERO, ONE, ONE}};
    uint8_t st = 0;
    for (size_t b = 0; b < a->n_ballots && !st; b++)
      for (int h = 0; h < 2; h++) {
        aff p;
        if (!pt_from_bytes(&p, a->ct + ((b * a->n_fields + f) * 128) + 64 * h)) { st = 1; break; }
        prj q = {p.x, p.y, ONE};
        prj_add(&acc[h], &acc[h], &q);
      }
    memset(a->out + f * 128, 0, 128);
    for (int h = 0; h < 2 && !st; h++) {
      if (fe_is_zero(&acc[h].z)) { st = 5; break; }
      fe zi; aff r;
      fe_inv(&zi, &acc[h].z); fe_mul(&r.x, &acc[h].x, &zi); fe_mul(&r.y, &acc[h].y, &zi);
      pt_to_bytes(a->out + f * 128 + 64 * h, &r);
    }
    if (a->status) a->status[f] = st;
  }
}
int oracle_elgamal_tally(const uint8_t* ct, size_t n_ballots, int n_fields, uint8_t* out, uint8_t* status, int threads) {
  ed_init();
  tally_args a = {ct, n_ballots, n_fields, out, status};
  parallel_for(tally_range, &a, (size_t)n_fields, threads);
  return 0;
}
int oracle_fixed_base_mul(const uint8_t* scalars, size_t n, uint8_t* out) {
  ed_init();
  for (size_t i = 0; i < n; i++) {
    aff r; fe t;
    memset(out + i * 64, 0, 64);
    if (fe_from_bytes(&t, scalars + i * 32) && fixed_base_mul(&r, scalars + i * 32)) pt_to_bytes(out + i * 64, &r);
  }
  return 0;
}

/* ---- legacy Keccak-256 / address (address.go:14-40) ------------------------------------------------ */
static const u64 KRC[24] = {
    0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808AULL, 0x8000000080008000ULL, 0x000000000000808BULL,
    0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL, 0x000000000000008AULL, 0x0000000000000088ULL,
    0x0000000080008009ULL, 0x000000008000000AULL, 0x000000008000808BULL, 0x800000000000008BULL, 0x8000000000008089ULL,
    0x8000000000008003ULL, 0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800AULL, 0x800000008000000AULL,
    0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
static const int KROT[5][5] = {{0, 36, 3, 41, 18}, {1, 44, 10, 45, 2}, {62, 6, 43, 15, 61}, {28, 55, 25, 21, 56}, {27, 20, 39, 8, 14}};
static inline u64 rol(u64 x, int n) { return n ? (x << n) | (x >> (64 - n)) : x; }
static void keccak_f(u64 a[25]) {
  for (int rnd = 0; rnd < 24; rnd++) {
    u64 c[5], d[5], b[25];
    for (int x = 0; x < 5; x++) c[x] = a[x] ^ a[x + 5] ^ a[x + 10] ^ a[x + 15] ^ a[x + 20];
    for (int x = 0; x < 5; x++) d[x] = c[(x + 4) % 5] ^ rol(c[(x + 1) % 5], 1);
    for (int i = 0; i < 25; i++) a[i] ^= d[i % 5];
    for (int x = 0; x < 5; x++) for (int y = 0; y < 5; y++) b[y + 5 * ((2 * x + 3 * y) % 5)] = rol(a[x + 5 * y], KROT[x][y]);
    for (int y = 0; y < 5; y++) for (int x = 0; x < 5; x++) a[x + 5 * y] = b[x + 5 * y] ^ (~b[(x + 1) % 5 + 5 * y] & b[(x + 2) % 5 + 5 * y]);
    a[0] ^= KRC[rnd];
  }
}
typedef struct { const uint8_t* in; uint8_t* out; } addr_args;
static void addr_range(void* vp, size_t lo, size_t hi) {
  addr_args* a = (addr_args*)vp;
  for (size_t i = lo; i < hi; i++) {
    u64 st[25]; uint8_t blk[136];
    memset(st, 0, sizeof st); memset(blk, 0, sizeof blk);
    memcpy(blk, a->in + i * 64, 64); blk[64] = 0x01; blk[135] |= 0x80;
    for (int w = 0; w < 17; w++) { u64 v; memcpy(&v, blk + 8 * w, 8); st[w] ^= v; }
    keccak_f(st);
    uint8_t h[32]; memcpy(h, st, 32);
    memcpy(a->out + i * 20, h + 12, 20);
  }
}
int oracle_keccak_address(const uint8_t* pub_xy_be, size_t n, uint8_t* out_addr, int threads) {
  addr_args a = {pub_xy_be, out_addr};
  parallel_for(addr_range, &a, n, threads);
  return 0;
}
