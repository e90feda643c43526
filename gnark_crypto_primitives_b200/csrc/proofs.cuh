// "Next" rows of SURVEY.md 8(f): compositions of the hot-path kernels.
//   AssertDecrypt            /root/reference/elgamal/ciphertext.go:50-67
//   DecryptionProof.Verify   /root/reference/elgamal/ciphertext.go:124-168 (+ hashPointsToScalar :173-184), hFn = MultiHash
//   FromTEtoRTE / FromRTEtoTE /root/reference/ecc/format/twistededwards.go:29-48 (scalingFactor :17)
// The gadgets ASSERT; here every item gets a flag (1 = all assertions hold) and a status for malformed input.
#pragma once
#include "edwards.cuh"
#include "elgamal.cuh"
#include "poseidon.cuh"
#include "smt.cuh"

namespace gcp {

// [k]P for an on-curve P and an integer k < 2^254: signed 4-bit windows (edwards.cuh).
__device__ __forceinline__ void ext_scalar_mul(ExtPoint& out, const ExtPoint& base, const u32 (&k)[8]) {
  ext_scalar_mul_windowed(out, base, k);
}

// projective equality of two extended points (Z != 0 on both sides for curve points)
__device__ __noinline__ bool ext_equal(const ExtPoint& p, const ExtPoint& q) {
  u32 a[8], b[8];
  fr_mul_call(a, p.X, q.Z);
  fr_mul_call(b, q.X, p.Z);
  fr_canon(a);
  fr_canon(b);
  if (!eq256(a, b)) return false;
  fr_mul_call(a, p.Y, q.Z);
  fr_mul_call(b, q.Y, p.Z);
  fr_canon(a);
  fr_canon(b);
  return eq256(a, b);
}

__device__ __forceinline__ void ext_neg(ExtPoint& p) {
  u32 t[8];
  fr_neg(t, p.X);
  fr_copy(p.X, t);
  fr_neg(t, p.T);
  fr_copy(p.T, t);
}

// affine point from memory: canonical check, Montgomery conversion, on-curve check; result extended
// (out of line: five call sites in the decryption-proof kernel)
__device__ __noinline__ void load_curve_point(ExtPoint& p, u32 (&x)[8], u32 (&y)[8], bool& canonical, bool& on_curve,
                                                 const u32* src, int mont) {
  u32 xs[8], ys[8];
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
  on_curve = on_curve && ed_is_on_curve(x, y);
  ext_from_affine(p, x, y);
}

__device__ __noinline__ bool ed_is_on_curve_ool(const u32 (&x)[8], const u32 (&y)[8]) { return ed_is_on_curve(x, y); }

// ---- AssertDecrypt: C1, C2 on curve;  C2 - [priv]C1 == [m]G ------------------------------------------------------
__global__ void __launch_bounds__(128, 5) assert_decrypt_kernel(const u32* __restrict__ tabG, const u32* __restrict__ cts,
                                                             const u32* __restrict__ privs, const u32* __restrict__ msgs, size_t n,
                                                             u8* __restrict__ flags, u8* __restrict__ status, int mont) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  bool canon = true, on_curve = true;
  ExtPoint c1, c2;
  u32 x[8], y[8], priv[8], msg[8];
  load_curve_point(c1, x, y, canon, on_curve, cts + idx * 32, mont);
  load_curve_point(c2, x, y, canon, on_curve, cts + idx * 32 + 16, mont);
  load_scalar(priv, canon, privs + idx * 8, mont);
  load_scalar(msg, canon, msgs + idx * 8, mont);
  u8 st = !canon ? GCP_STATUS_NONCANONICAL : (!on_curve ? GCP_STATUS_OFF_CURVE : GCP_STATUS_OK);
  u8 flag = 0;
  if (st == GCP_STATUS_OK) {
    ExtPoint s, m;
    ext_scalar_mul(s, c1, priv);          // ciphertext.go:58
    ext_identity(m);
    fixed_base_accumulate_ool(m, msg, tabG);  // ciphertext.go:60
    ext_neg(s);
    ext_add_call(c2, s);                       // ciphertext.go:62
    flag = ext_equal(c2, m) ? 1 : 0;      // ciphertext.go:64-65
  }
  flags[idx] = flag;
  status[idx] = st;
}

// ---- DecryptionProof.Verify -----------------------------------------------------------------------------------------
// Inputs per item: pubkey (2), ciphertext (4), msg (1), A1 (2), A2 (2), Z (1) elements.
__global__ void __launch_bounds__(128, 5) decryption_proof_kernel(const u32* __restrict__ tabG, PoseidonTable tab13,
                                                              const u32* __restrict__ pks, const u32* __restrict__ cts,
                                                              const u32* __restrict__ msgs, const u32* __restrict__ a1s,
                                                              const u32* __restrict__ a2s, const u32* __restrict__ zs, size_t n,
                                                              u8* __restrict__ flags, u8* __restrict__ status, int mont) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  bool canon = true, on_curve = true;
  ExtPoint pk, c1, c2, a1, a2;
  u32 coords[13][8], scratch[13][8];  // Poseidon state: [0] capacity, then PK.x, PK.y, PK.x, PK.y, C1, D, A1, A2
  u32 msg[8], z[8];
  load_curve_point(pk, coords[1], coords[2], canon, on_curve, pks + idx * 16, mont);
  load_curve_point(c1, coords[5], coords[6], canon, on_curve, cts + idx * 32, mont);
  {
    u32 x[8], y[8];
    load_curve_point(c2, x, y, canon, on_curve, cts + idx * 32 + 16, mont);
  }
  load_curve_point(a1, coords[9], coords[10], canon, on_curve, a1s + idx * 16, mont);
  load_curve_point(a2, coords[11], coords[12], canon, on_curve, a2s + idx * 16, mont);
  load_scalar(msg, canon, msgs + idx * 8, mont);
  load_scalar(z, canon, zs + idx * 8, mont);
  u8 st = !canon ? GCP_STATUS_NONCANONICAL : (!on_curve ? GCP_STATUS_OFF_CURVE : GCP_STATUS_OK);
  u8 flag = 0;
  if (st == GCP_STATUS_OK) {
    // D = C2 - [msg]G   (ciphertext.go:137-139)
    ExtPoint m, d;
    ext_identity(m);
    fixed_base_accumulate_ool(m, msg, tabG);
    ext_neg(m);
    d = c2;
    ext_add_call(d, m);
    // affine D for the Fiat-Shamir hash
    u32 zi[8], zc[8];
    fr_copy(zc, d.Z);
    fr_canon(zc);
    if (is_zero256(zc)) {
      st = GCP_STATUS_ZERO_DENOM;
    } else {
      fr_inv(zi, d.Z);
      fr_mul_call(coords[7], d.X, zi);
      fr_mul_call(coords[8], d.Y, zi);
      fr_copy(coords[3], coords[1]);
      fr_copy(coords[4], coords[2]);
      fr_set_zero(coords[0]);
      // E = MultiHash(PK, PK, C1, D, A1, A2): 12 inputs -> one Hash with t = 13 (ciphertext.go:141, :173-184)
      u32 e_m[8], e[8];
      poseidon_permute_generic(coords, scratch, e_m, tab13);
      fr_from_mont(e, e_m);  // the challenge is used as an integer scalar
      // z*G == A1 + e*P   (ciphertext.go:143-151)
      ExtPoint zg, ep;
      ext_identity(zg);
      fixed_base_accumulate_ool(zg, z, tabG);
      ext_scalar_mul(ep, pk, e);
      ext_add_call(ep, a1);
      bool ok = ext_equal(ep, zg);
      // z*C1 == A2 + e*D   (ciphertext.go:153-166)
      ExtPoint zc1, ed;
      ext_scalar_mul(zc1, c1, z);
      ext_scalar_mul(ed, d, e);
      ext_add_call(ed, a2);
      ok = ok && ext_equal(ed, zc1);
      flag = ok ? 1 : 0;
    }
  }
  flags[idx] = flag;
  status[idx] = st;
}

// ---- EdDSA-Poseidon IsValid (/root/reference/ecc/bn254/eddsa/verifier.go:55-88) ---------------------------------------
// A, R in TE (circom/iden3) coordinates; h = Poseidon(R.x, R.y, A.x, A.y, msg) on those coordinates (t = 6);
// A' = RTE(A), R' = RTE(R) asserted on the a = -1 curve; flag = ([S]G == 8*[h]A' + R')  (rteB8 == G, constants.go:11-18).
__global__ void __launch_bounds__(128, 5) eddsa_verify_kernel(const u32* __restrict__ tabG, PoseidonTable tab6,
                                                          const u32* __restrict__ pub_a, const u32* __restrict__ sig_r,
                                                          const u32* __restrict__ sig_s, const u32* __restrict__ msgs, size_t n,
                                                          u8* __restrict__ flags, u8* __restrict__ status, int mont) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  const u32 negf[8] = {0xc9603c7bu, 0x5c62c8e0u, 0x8fabc7f1u, 0xf8382911u, 0x6aa07f4du, 0x7d53da81u, 0x6ba06ab6u, 0x1da7c5b3u};
  bool canon = true;
  u32 st6[POSEIDON_MAX_T][8], scratch[POSEIDON_MAX_T][8];
  u32 s_int[8];
  // state: [0], R.x, R.y, A.x, A.y, msg (lazy Montgomery)
  load_elem(st6[1], canon, sig_r + idx * 16, mont);
  load_elem(st6[2], canon, sig_r + idx * 16 + 8, mont);
  load_elem(st6[3], canon, pub_a + idx * 16, mont);
  load_elem(st6[4], canon, pub_a + idx * 16 + 8, mont);
  load_elem(st6[5], canon, msgs + idx * 8, mont);
  load_scalar(s_int, canon, sig_s + idx * 8, mont);
  fr_set_zero(st6[0]);
  // RTE conversion: x * (-f); y unchanged
  u32 ax[8], ay[8], rx[8], ry[8];
  fr_mul_call(rx, st6[1], negf);
  fr_copy(ry, st6[2]);
  fr_mul_call(ax, st6[3], negf);
  fr_copy(ay, st6[4]);
  bool on_curve = ed_is_on_curve_ool(ax, ay) && ed_is_on_curve_ool(rx, ry);  // PointToRTE, verifier.go:46
  u8 st = !canon ? GCP_STATUS_NONCANONICAL : (!on_curve ? GCP_STATUS_OFF_CURVE : GCP_STATUS_OK);
  u8 flag = 0;
  if (st == GCP_STATUS_OK) {
    u32 h_m[8], h[8];
    poseidon_permute_generic(st6, scratch, h_m, tab6);
    fr_from_mont(h, h_m);
    ExtPoint left, a, r, r1;
    ext_identity(left);
    fixed_base_accumulate_ool(left, s_int, tabG);  // [S] rteB8
    ext_from_affine(a, ax, ay);
    ext_from_affine(r, rx, ry);
    ext_scalar_mul(r1, a, h);
    ext_double_call(r1, false);  // verifier.go:72-74: three doublings; T is only read by the addition after the last
    ext_double_call(r1, false);
    ext_double_call(r1, true);
    ext_add_call(r1, r);
    flag = ext_equal(left, r1) ? 1 : 0;
  }
  flags[idx] = flag;
  status[idx] = st;
}

// ---- TE <-> RTE: x' = x * (-f) or x / (-f), y unchanged (ecc/format/twistededwards.go:29-48) -------------------------
__global__ void te_rte_kernel(const u32* __restrict__ in, size_t n_points, u32* __restrict__ out, u8* __restrict__ status,
                              int to_rte) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_points) return;
  const u32 negf[8] = {0xc9603c7bu, 0x5c62c8e0u, 0x8fabc7f1u, 0xf8382911u, 0x6aa07f4du, 0x7d53da81u, 0x6ba06ab6u, 0x1da7c5b3u};
  const u32 negf_inv[8] = {0xb1b017d8u, 0x61d380bfu, 0x8415d72eu, 0x7f5d8063u, 0x294f7a18u, 0x77e18e30u, 0x305733c2u, 0x10d2ede5u};
  u32 x[8], y[8], r[8];
  load_fr(x, in + idx * 16);
  load_fr(y, in + idx * 16 + 8);
  bool canon = fr_is_canonical(x) && fr_is_canonical(y);
  // mont_mul(x, c*R) = x*c in whichever representation x is in (standard or Montgomery)
  if (to_rte)
    fr_mul(r, x, negf);
  else
    fr_mul(r, x, negf_inv);
  fr_canon(r);
  if (!canon) {
    fr_set_zero(r);
    fr_set_zero(y);
  }
  store_fr(out + idx * 16, r);
  store_fr(out + idx * 16 + 8, y);
  status[idx] = canon ? GCP_STATUS_OK : GCP_STATUS_NONCANONICAL;
}

}  // namespace gcp
