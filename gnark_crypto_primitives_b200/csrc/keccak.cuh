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

__device__ __forceinline__ u64 rotl64(u64 x, int n) {  // n is a compile-time constant at every call site
  return n == 0 ? x : (x << n) | (x >> (64 - n));
}

__device__ __forceinline__ void keccak_f1600(u64 (&a)[25]) {
#pragma unroll 1
  for (int rnd = 0; rnd < 24; rnd++) {
    u64 c0 = a[0] ^ a[5] ^ a[10] ^ a[15] ^ a[20];
    u64 c1 = a[1] ^ a[6] ^ a[11] ^ a[16] ^ a[21];
    u64 c2 = a[2] ^ a[7] ^ a[12] ^ a[17] ^ a[22];
    u64 c3 = a[3] ^ a[8] ^ a[13] ^ a[18] ^ a[23];
    u64 c4 = a[4] ^ a[9] ^ a[14] ^ a[19] ^ a[24];
    u64 d0 = c4 ^ rotl64(c1, 1), d1 = c0 ^ rotl64(c2, 1), d2 = c1 ^ rotl64(c3, 1), d3 = c2 ^ rotl64(c4, 1),
        d4 = c3 ^ rotl64(c0, 1);
    // theta + rho + pi: b[y + 5*((2x+3y) % 5)] = rotl(a[x + 5y] ^ d[x], r[x][y])
    u64 b[25];
    b[0] = a[0] ^ d0;
    b[10] = rotl64(a[1] ^ d1, 1);
    b[20] = rotl64(a[2] ^ d2, 62);
    b[5] = rotl64(a[3] ^ d3, 28);
    b[15] = rotl64(a[4] ^ d4, 27);
    b[16] = rotl64(a[5] ^ d0, 36);
    b[1] = rotl64(a[6] ^ d1, 44);
    b[11] = rotl64(a[7] ^ d2, 6);
    b[21] = rotl64(a[8] ^ d3, 55);
    b[6] = rotl64(a[9] ^ d4, 20);
    b[7] = rotl64(a[10] ^ d0, 3);
    b[17] = rotl64(a[11] ^ d1, 10);
    b[2] = rotl64(a[12] ^ d2, 43);
    b[12] = rotl64(a[13] ^ d3, 25);
    b[22] = rotl64(a[14] ^ d4, 39);
    b[23] = rotl64(a[15] ^ d0, 41);
    b[8] = rotl64(a[16] ^ d1, 45);
    b[18] = rotl64(a[17] ^ d2, 15);
    b[3] = rotl64(a[18] ^ d3, 21);
    b[13] = rotl64(a[19] ^ d4, 8);
    b[14] = rotl64(a[20] ^ d0, 18);
    b[24] = rotl64(a[21] ^ d1, 2);
    b[9] = rotl64(a[22] ^ d2, 61);
    b[19] = rotl64(a[23] ^ d3, 56);
    b[4] = rotl64(a[24] ^ d4, 14);
    // chi
#pragma unroll
    for (int y = 0; y < 25; y += 5) {
#pragma unroll
      for (int x = 0; x < 5; x++) a[y + x] = b[y + x] ^ (~b[y + (x + 1) % 5] & b[y + (x + 2) % 5]);
    }
    a[0] ^= KECCAK_RC[rnd];  // iota
  }
}

// in: n x 64 bytes (X_be || Y_be), out: n x 20 bytes
__global__ void __launch_bounds__(256) keccak_address_kernel(const u8* __restrict__ in, size_t n, u8* __restrict__ out) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  u64 a[25];
  const uint4* src = reinterpret_cast<const uint4*>(in + idx * 64);
#pragma unroll
  for (int i = 0; i < 4; i++) {
    uint4 v = __ldg(src + i);
    a[2 * i] = ((u64)v.y << 32) | v.x;  // lanes are little-endian words of the byte stream
    a[2 * i + 1] = ((u64)v.w << 32) | v.z;
  }
  a[8] = 0x01ull;                    // legacy Keccak padding byte right after the 64-byte message
#pragma unroll
  for (int i = 9; i < 25; i++) a[i] = 0;
  a[16] = 0x8000000000000000ull;     // last byte of the 136-byte rate block
  keccak_f1600(a);
  // digest bytes 12..31 = high half of lane 1, lanes 2 and 3
  u32* dst = reinterpret_cast<u32*>(out + idx * 20);
  dst[0] = (u32)(a[1] >> 32);
  dst[1] = (u32)a[2];
  dst[2] = (u32)(a[2] >> 32);
  dst[3] = (u32)a[3];
  dst[4] = (u32)(a[3] >> 32);
}

}  // namespace gcp
