// Ethereum address derivation: last 20 bytes of legacy Keccak-256 over X_be32 || Y_be32.
// Values of DeriveAddress, /root/reference/ecc/secp256k1/ecdsa/address.go:14-40 (no EC math: the input already is
// a public key); Keccak-f[1600] itself is gnark's std/hash/sha3 (un-vendored), restated from the Keccak
// specification: rate 136, one block for a 64-byte message, legacy padding 0x01 .. 0x80.
// One thread per key; the 25 lanes live in registers; bound by the ALU pipe (LOP3 / SHF), not by HBM
// (84 bytes per item).
#pragma once
#include <cstdint>
#include "kernels.h"

namespace gcp {

__device__ __constant__ const u64 KECCAK_RC[24] = {
    0x0000000000000001ull, 0x0000000000008082ull, 0x800000000000808Aull, 0x8000000080008000ull, 0x000000000000808Bull,
    0x0000000080000001ull, 0x8000000080008081ull, 0x8000000000008009ull, 0x000000000000008Aull, 0x0000000000000088ull,
    0x0000000080008009ull, 0x000000008000000Aull, 0x000000008000808Bull, 0x800000000000008Bull, 0x8000000000008089ull,
    0x8000000000008003ull, 0x8000000000008002ull, 0x8000000000000080ull, 0x000000000000800Aull, 0x800000008000000Aull,
    0x8000000080008081ull, 0x8000000000008080ull, 0x0000000080000001ull, 0x8000000080008008ull};

// The permutation on 32-bit halves with explicit three-input logic ops.  ncu of the first version (u64 expressions left
// to the compiler) showed the ALU pipe 97 % busy with 5 184 LOP3 / SHF / IADD3 per address: the five-way column parities
// were built from two-input XORs and the theta offsets d[x] materialised.  Here per round: column parity = 2 LOP3 per
// half (a ^ b ^ c, LUT 0x96), theta folded into the rho input as a ^ c[x-1] ^ rot(c[x+1], 1) (one LOP3 per half lane, no
// d[x]), rho = two funnel shifts per lane, chi = one LOP3 per half lane (a ^ (~b & c), LUT 0xD2): 122 LOP3 + 58 SHF.
__device__ __forceinline__ u32 xor3(u32 a, u32 b, u32 c) {
  u32 r;
  asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
  return r;
}
__device__ __forceinline__ u32 chi32(u32 a, u32 b, u32 c) {  // a ^ (~b & c)
  u32 r;
  asm("lop3.b32 %0, %1, %2, %3, 0xD2;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
  return r;
}
// (lo, hi) <- rotl64((lo, hi), N), N a compile-time constant in [0, 63]
template <int N>
__device__ __forceinline__ void rotl64_halves(u32& lo, u32& hi, u32 xl, u32 xh) {
  if (N == 0) {
    lo = xl;
    hi = xh;
  } else if (N < 32) {
    hi = __funnelshift_l(xl, xh, N);
    lo = __funnelshift_l(xh, xl, N);
  } else if (N == 32) {
    lo = xh;
    hi = xl;
  } else {
    hi = __funnelshift_l(xh, xl, N >= 32 ? N - 32 : 0);
    lo = __funnelshift_l(xl, xh, N >= 32 ? N - 32 : 0);
  }
}

// b[DST] = rotl(a[SRC] ^ d[x], R) with d[x] = c[x-1] ^ rot(c[x+1], 1) folded in: (pl, ph) = c[x-1], (rl, rh) = rot(c[x+1], 1)
#define KECCAK_RHO_PI(DST, SRC, R, X)                                                                       \
  rotl64_halves<R>(bl[DST], bh[DST], xor3(al[SRC], cl[(X + 4) % 5], rl[(X + 1) % 5]),                         \
                   xor3(ah[SRC], ch[(X + 4) % 5], rh[(X + 1) % 5]))

__device__ __forceinline__ void keccak_f1600(u32 (&al)[25], u32 (&ah)[25]) {
#pragma unroll 1
  for (int rnd = 0; rnd < 24; rnd++) {
    u32 cl[5], ch[5], rl[5], rh[5];
#pragma unroll
    for (int x = 0; x < 5; x++) {
      cl[x] = xor3(xor3(al[x], al[x + 5], al[x + 10]), al[x + 15], al[x + 20]);
      ch[x] = xor3(xor3(ah[x], ah[x + 5], ah[x + 10]), ah[x + 15], ah[x + 20]);
    }
#pragma unroll
    for (int x = 0; x < 5; x++) rotl64_halves<1>(rl[x], rh[x], cl[x], ch[x]);
    // theta + rho + pi: b[y + 5*((2x+3y) % 5)] = rotl(a[x + 5y] ^ d[x], r[x][y])
    u32 bl[25], bh[25];
    KECCAK_RHO_PI(0, 0, 0, 0);
    KECCAK_RHO_PI(10, 1, 1, 1);
    KECCAK_RHO_PI(20, 2, 62, 2);
    KECCAK_RHO_PI(5, 3, 28, 3);
    KECCAK_RHO_PI(15, 4, 27, 4);
    KECCAK_RHO_PI(16, 5, 36, 0);
    KECCAK_RHO_PI(1, 6, 44, 1);
    KECCAK_RHO_PI(11, 7, 6, 2);
    KECCAK_RHO_PI(21, 8, 55, 3);
    KECCAK_RHO_PI(6, 9, 20, 4);
    KECCAK_RHO_PI(7, 10, 3, 0);
    KECCAK_RHO_PI(17, 11, 10, 1);
    KECCAK_RHO_PI(2, 12, 43, 2);
    KECCAK_RHO_PI(12, 13, 25, 3);
    KECCAK_RHO_PI(22, 14, 39, 4);
    KECCAK_RHO_PI(23, 15, 41, 0);
    KECCAK_RHO_PI(8, 16, 45, 1);
    KECCAK_RHO_PI(18, 17, 15, 2);
    KECCAK_RHO_PI(3, 18, 21, 3);
    KECCAK_RHO_PI(13, 19, 8, 4);
    KECCAK_RHO_PI(14, 20, 18, 0);
    KECCAK_RHO_PI(24, 21, 2, 1);
    KECCAK_RHO_PI(9, 22, 61, 2);
    KECCAK_RHO_PI(19, 23, 56, 3);
    KECCAK_RHO_PI(4, 24, 14, 4);
    // chi
#pragma unroll
    for (int y = 0; y < 25; y += 5) {
#pragma unroll
      for (int x = 0; x < 5; x++) {
        al[y + x] = chi32(bl[y + x], bl[y + (x + 1) % 5], bl[y + (x + 2) % 5]);
        ah[y + x] = chi32(bh[y + x], bh[y + (x + 1) % 5], bh[y + (x + 2) % 5]);
      }
    }
    const u64 rc = KECCAK_RC[rnd];  // iota
    al[0] ^= (u32)rc;
    ah[0] ^= (u32)(rc >> 32);
  }
}
#undef KECCAK_RHO_PI

// in: n x 64 bytes (X_be || Y_be), out: n x 20 bytes
__global__ void __launch_bounds__(256) keccak_address_kernel(const u8* __restrict__ in, size_t n, u8* __restrict__ out) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  u32 al[25], ah[25];
  const uint4* src = reinterpret_cast<const uint4*>(in + idx * 64);
#pragma unroll
  for (int i = 0; i < 4; i++) {
    uint4 v = __ldg(src + i);
    al[2 * i] = v.x;  // lanes are little-endian words of the byte stream
    ah[2 * i] = v.y;
    al[2 * i + 1] = v.z;
    ah[2 * i + 1] = v.w;
  }
#pragma unroll
  for (int i = 8; i < 25; i++) al[i] = ah[i] = 0;
  al[8] = 0x01u;                     // legacy Keccak padding byte right after the 64-byte message
  ah[16] = 0x80000000u;              // last byte of the 136-byte rate block
  keccak_f1600(al, ah);
  // digest bytes 12..31 = high half of lane 1, lanes 2 and 3
  u32* dst = reinterpret_cast<u32*>(out + idx * 20);
  dst[0] = ah[1];
  dst[1] = al[2];
  dst[2] = ah[2];
  dst[3] = al[3];
  dst[4] = ah[3];
}

}  // namespace gcp
