// BN254 Fr on the FP64 pipe of sm_100a: 5 x 52-bit limbs held in doubles, Montgomery form with R = 2^260.
//
// Why: on B200 DFMA issues at 64 lanes/clk/SM and IMAD.WIDE.U32 at 32 (bench_micro/imad_peak.cu), and the two
// share one pipe (bench_micro/dfma_mix.cu: a warp mix of the two takes the SUM of their stand-alone times).  A
// 52x52-bit limb product costs 2 DFMA + 1 DADD = 1.5 pipe slots and yields 2704 bit-products; a 32x32 one costs 1
// slot and yields 1024.  A Montgomery multiplication is 128 wide multiplies in fr.cuh and 180 FP64 issues (= 90
// slots) here.
//
// The limb product (Emmart, "Faster modular exponentiation using double precision floating point arithmetic on
// the GPU"): for integers 0 <= a, b < 2^52 held exactly in doubles,
//     hi = fma_rz(a, b, 2^104)            = 2^104 + floor(ab / 2^52) * 2^52   (ulp of [2^104, 2^105) is 2^52)
//     lo = fma_rz(a, b, (2^104 + 2^52) - hi) = 2^52 + (ab mod 2^52)           (exact)
// so the IEEE bit patterns are  C1B + H  and  C3B + L  with H, L the two 52-bit halves as plain integers, and the
// column sums of a schoolbook product are 64-bit INTEGER additions of bit patterns (ptxas folds two of them into
// one IADD3 / IADD3.X pair).  The constants that pile up (a known multiple of C1B / C3B per column) are folded into
// the accumulator's initial value, so after the last addition a column holds its true value.
//
// Values: every element is 5 limbs < 2^52 (value < 2^260 ~ 84.6 r); mont(a, b) = a b 2^-260 mod r satisfies
// mont(a,b) < (a/r)(b/r) r / 77.3 + r, so inputs up to several r never need a conditional subtraction.
//
// STATUS: experiment, NOT used by the product.  Exact (20 000 products incl. extreme operands checked against Python
// integers through the host build below; the device result is checked against the host build by fr52_mul.cu), but
// measured at 61.6 G mul/s against 63.9 G mul/s for fr.cuh's fr_mul on the same B200 (profiles/r01_dfma_*.jsonl):
// the FP64 pipe is only 59.5 % busy because every FP64 issue holds the dispatch port for two cycles and the ~170
// integer instructions per product (bit-pattern additions, masks, carries) do not hide behind them
// (ncu: math_pipe_throttle 3.0, dispatch_stall 1.6 per issue, 1.96 eligible warps per cycle, issue 0.57/clk).
//
// Host build: the same code compiles with g++ (-mfma -frounding-math) when the caller sets FE_TOWARDZERO.
#pragma once
#include <cstdint>
#include <cmath>
#include <cstring>

namespace gcp {
namespace f52 {

typedef unsigned long long u64;
typedef unsigned int u32;

#if defined(__CUDACC__)
#define F52_FN __host__ __device__ __forceinline__
#define F52_CX __host__ __device__ constexpr
#else
#define F52_FN static inline
#define F52_CX constexpr
#endif
// round-toward-zero fma: DFMA.RZ on the device; on the host the caller sets fesetround(FE_TOWARDZERO)
F52_FN double fma_rz(double a, double b, double c) {
#if defined(__CUDA_ARCH__)
  return __fma_rz(a, b, c);
#else
  return std::fma(a, b, c);
#endif
}
F52_FN u64 d2b(double x) {
#if defined(__CUDA_ARCH__)
  return (u64)__double_as_longlong(x);
#else
  u64 r; std::memcpy(&r, &x, 8); return r;
#endif
}
F52_FN double b2d(u64 x) {
#if defined(__CUDA_ARCH__)
  return __longlong_as_double((long long)x);
#else
  double r; std::memcpy(&r, &x, 8); return r;
#endif
}

constexpr u64 C1B = 0x4670000000000000ull;  // bits of 2^104
constexpr u64 C3B = 0x4330000000000000ull;  // bits of 2^52
constexpr u64 M52 = 0x000fffffffffffffull;
#define F52_C1 0x1.0p104
#define F52_C2 0x1.0000000000001p104
#define F52_C3 0x1.0p52

// r and -r^-1 mod 2^52 as 52-bit limbs (exact doubles)
#define F52_R0 0x1f593f0000001p0
#define F52_R1 0x4879b9709143ep0
#define F52_R2 0x181585d2833e8p0
#define F52_R3 0xa029b85045b68p0
#define F52_R4 0x30644e72e131p0
#define F52_NP 0x1f593efffffffp0

struct El {
  double l[5];
};

// number of (i, j) in 5 x 5 with i + j = k
F52_CX int nlo(int k) { return (k < 0 || k > 8) ? 0 : (k <= 4 ? k + 1 : 9 - k); }

// Initial value of column k for an accumulator that will take NPROD 5x5 products and NADD biased additions
// (add_biased) before ONE reduction.  Lower columns are read by reduction row k when rows 0..k-1 have run; upper
// columns are read after everything.  0x433 = C3B >> 52 is the bias that rides on every inter-row carry.
F52_CX u64 col_init(int k, int nprod, int nadd) {
  u64 s = (u64)nprod * ((u64)nlo(k) * C3B + (u64)nlo(k - 1) * C1B);
  if (k <= 4) {
    s += (u64)k * C3B + (u64)k * C1B + (k > 0 ? 0x433ull : 0ull);
  } else {
    s += (u64)nlo(k) * C3B + (u64)nlo(k - 1) * C1B + (u64)nadd * C3B + (k == 5 ? 0x433ull : 0ull);
  }
  return 0ull - s;
}

struct Acc {
  u64 c[10];
};

template <int NPROD, int NADD>
F52_FN void acc_init(Acc& w) {
#pragma unroll
  for (int k = 0; k < 10; k++) w.c[k] = col_init(k, NPROD, NADD);
}

// one limb product into columns k (low half) and k + 1 (high half)
F52_FN void dprod(u64& clo, u64& chi, double a, double b) {
  double hi = fma_rz(a, b, F52_C1);
#if defined(F52_EXPERIMENT) && F52_EXPERIMENT == 3  // timing experiments only (results are wrong): see fr52_mul.cu
  double lo = fma_rz(a, b, F52_C2);
#else
  double lo = fma_rz(a, b, F52_C2 - hi);
#endif
#if defined(F52_EXPERIMENT) && F52_EXPERIMENT == 1
  clo ^= d2b(lo);
  chi ^= d2b(hi);
#elif defined(F52_EXPERIMENT) && F52_EXPERIMENT == 2
  clo = (u64)((u32)clo + (u32)d2b(lo)) | ((u64)((u32)(clo >> 32) + (u32)(d2b(lo) >> 32)) << 32);
  chi = (u64)((u32)chi + (u32)d2b(hi)) | ((u64)((u32)(chi >> 32) + (u32)(d2b(hi) >> 32)) << 32);
#else
  clo += d2b(lo);
  chi += d2b(hi);
#endif
}

// w += a * b  (25 limb products)
F52_FN void mac(Acc& w, const double (&a)[5], const double (&b)[5]) {
#pragma unroll
  for (int i = 0; i < 5; i++)
#pragma unroll
    for (int j = 0; j < 5; j++) dprod(w.c[i + j], w.c[i + j + 1], a[i], b[j]);
}

// w += x * 2^260 for x given in biased form (limb + 2^52 as a double); counted by NADD
F52_FN void add_biased(Acc& w, const double (&xb)[5]) {
#pragma unroll
  for (int i = 0; i < 5; i++) w.c[5 + i] += d2b(xb[i]);
}
// same for a plain element (one DADD per limb to bias it)
F52_FN void add_plain(Acc& w, const double (&x)[5]) {
#pragma unroll
  for (int i = 0; i < 5; i++) w.c[5 + i] += d2b(x[i] + F52_C3);
}

// Montgomery reduction: out = w / 2^260 mod r (not canonical: out < w / 2^260 + r), limbs normalised to < 2^52.
F52_FN void redc(Acc& w, double (&out)[5]) {
  const double rr[5] = {F52_R0, F52_R1, F52_R2, F52_R3, F52_R4};
#pragma unroll
  for (int i = 0; i < 5; i++) {
    if (i > 0) w.c[i] += w.c[i - 1] >> 52;
    // q = (column mod 2^52) * (-r^-1) mod 2^52
    double t = b2d((w.c[i] & M52) | C3B) - F52_C3;
    double qh = fma_rz(t, F52_NP, F52_C1);
    double q = fma_rz(t, F52_NP, F52_C2 - qh) - F52_C3;
#pragma unroll
    for (int j = 0; j < 5; j++) dprod(w.c[i + j], w.c[i + j + 1], q, rr[j]);
  }
  u64 carry = w.c[4] >> 52;
#pragma unroll
  for (int i = 0; i < 5; i++) {
    u64 v = w.c[5 + i] + carry;
    carry = v >> 52;
    out[i] = b2d((v & M52) | C3B) - F52_C3;
  }
}

// out = a * b * 2^-260 mod r
F52_FN void mul(double (&out)[5], const double (&a)[5], const double (&b)[5]) {
  Acc w;
  acc_init<1, 0>(w);
  mac(w, a, b);
  redc(w, out);
}

}  // namespace f52
}  // namespace gcp
