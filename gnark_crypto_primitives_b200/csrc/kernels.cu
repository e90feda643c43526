// All device code of the engine lives in this one translation unit (the constant-memory tables are shared by
// the Poseidon, SMT and ElGamal kernels without relocatable device code).  Host-side launch wrappers only;
// the C ABI, chunking and memory pools are in capi.cu.
#include "fr.cuh"
#include "poseidon.cuh"
#include "poseidon_lanes.cuh"
#include "smt.cuh"
#include "elgamal.cuh"
#include "keccak.cuh"
#include "proofs.cuh"
#include "varbase.cuh"
#include "varbase_reg.cuh"
#include "varbase_flat.cuh"
#include "mimc7.cuh"
#include "poseidon2.cuh"
#include "kernels.h"
#include "chunkplan.h"

#include <algorithm>
#include <cstdlib>

namespace gcp {

// ---------------------------------------------------------------------------------------------------
// setup kernels
// ---------------------------------------------------------------------------------------------------
__global__ void to_mont_kernel(u32* elems, size_t n) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  u32 x[8], m[8];
  load_fr(x, elems + idx * 8);
  fr_to_mont(m, x);
  fr_canon(m);
  store_fr(elems + idx * 8, m);
}

cudaError_t launch_to_mont(u32* d_elems, size_t n, cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  to_mont_kernel<<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(d_elems, n);
  return cudaGetLastError();
}

__global__ void poseidon_pair_constants_kernel(PoseidonTable tab, u32* __restrict__ out) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= tab.RP / 2) return;
  const int t = tab.t;
  const u32* srow_a = tab.S + (size_t)(2 * t - 1) * ((tab.RP & 1) + 2 * p) * 8;
  const u32* srow_b = srow_a + (2 * t - 1) * 8;
  u32 acc[8];
  fr_set_zero(acc);
#pragma unroll 1
  for (int k = 1; k < t; k++) {
    u32 a[8], c[8], m[8];
    load_fr(a, srow_b + k * 8);
    load_fr(c, srow_a + (t + k - 1) * 8);
    fr_mul(m, a, c);
    fr_add(acc, acc, m);
  }
  fr_canon(acc);
  store_fr(out + p * 8, acc);
}

cudaError_t launch_poseidon_pair_constants(const PoseidonTable& tab, u32* d_out, cudaStream_t stream) {
  poseidon_pair_constants_kernel<<<1, 64, 0, stream>>>(tab, d_out);
  return cudaGetLastError();
}

// d_q = S_q[1] S_{q-1}[3] + S_q[2] S_{q-1}[4] for q = 2, 4, ..., 56 (poseidon.cuh, t = 3 pair schedule); t3: the t = 3
// table in Montgomery form, C | S | M | P
__global__ void pos3_pair_kernel(const u32* __restrict__ t3) {
  const int p = threadIdx.x;
  if (p >= POS3_PAIR_COUNT) return;
  const u32* srow = t3 + (size_t)((8 * 3 + 57) + 5 * (2 * p + 2)) * 8;
  const u32* prow = srow - 5 * 8;
  u32 a[8], b[8], c1[8], c2[8], t1[8], t2[8];
  load_fr(a, srow + 8);
  load_fr(b, srow + 16);
  load_fr(c1, prow + 24);
  load_fr(c2, prow + 32);
  fr_mul(t1, a, c1);
  fr_mul(t2, b, c2);
  fr_add(t1, t1, t2);
  fr_canon(t1);
  store_fr(g_pos3_pair + p * 8, t1);
}

cudaError_t upload_const_tables(const u32* d_t3, const u32* d_t4, cudaStream_t stream) {
  cudaError_t e = cudaMemcpyToSymbolAsync(c_pos3, d_t3, sizeof(u32) * POS3_ELEMS * 8, 0, cudaMemcpyDeviceToDevice, stream);
  if (e != cudaSuccess) return e;
  pos3_pair_kernel<<<1, 32, 0, stream>>>(d_t3);
  if ((e = cudaGetLastError()) != cudaSuccess) return e;
  void* d_pair = nullptr;
  if ((e = cudaGetSymbolAddress(&d_pair, g_pos3_pair)) != cudaSuccess) return e;
  e = cudaMemcpyToSymbolAsync(c_pos3_pair, d_pair, sizeof(u32) * POS3_PAIR_COUNT * 8, 0, cudaMemcpyDeviceToDevice, stream);
  if (e != cudaSuccess) return e;
  return cudaMemcpyToSymbolAsync(c_pos4, d_t4, sizeof(u32) * POS4_ELEMS * 8, 0, cudaMemcpyDeviceToDevice, stream);
}


// ---------------------------------------------------------------------------------------------------
// Integer-pipe probe: carry-chained IMAD.WIDE.U32 / IMAD.WIDE.U32.X pairs on 8 independent accumulator sets per
// thread, no memory traffic.  Its rate is the roofline denominator bench.py reports against (same measurement as
// bench_micro/imad_peak.cu, "wide_cc_pair_chain": 32 lanes/clk/SM, about 9.0 T/s on a B200 at 1.9 GHz).
// ---------------------------------------------------------------------------------------------------
constexpr int PROBE_ITERS = 16384;
constexpr int PROBE_CHAINS = 8;
__global__ void __launch_bounds__(256) imad_probe_kernel(u32* out, u32 seed) {
  // PROBE_CHAINS independent pairs of 64-bit accumulators; each step is IMAD.WIDE.U32 (carry out) followed by
  // IMAD.WIDE.U32.X (carry in): exactly the two instruction forms the multiplier rows are made of.
  u64 a0[PROBE_CHAINS], a1[PROBE_CHAINS];
  u32 mult[PROBE_CHAINS];
  u32 x = seed * 2654435761u + threadIdx.x, y = x ^ 0x9e3779b9u;
#pragma unroll
  for (int c = 0; c < PROBE_CHAINS; c++) {
    mult[c] = x + c * 0x9e3779b9u;
    asm volatile("xor.b32 %0, %0, %1;" : "+r"(mult[c]) : "r"(seed));
    a0[c] = ((u64)(blockIdx.x + c) << 32) | (threadIdx.x * 4 + c);
    a1[c] = a0[c] * 0x9e3779b97f4a7c15ull;
  }
#pragma unroll 1
  for (int it = 0; it < PROBE_ITERS / 4; it++) {
#pragma unroll
    for (int rep = 0; rep < 4; rep++) {
#pragma unroll
      for (int c = 0; c < PROBE_CHAINS; c++) {
        asm volatile(
            "{\n\t"
            ".reg .u32 l0, h0, l1, h1;\n\t"
            "mov.b64 {l0, h0}, %0;\n\t"
            "mov.b64 {l1, h1}, %1;\n\t"
            "mad.lo.cc.u32 l0, %4, %2, l0;\n\t"
            "madc.hi.cc.u32 h0, %4, %2, h0;\n\t"
            "madc.lo.cc.u32 l1, %4, %3, l1;\n\t"
            "madc.hi.u32 h1, %4, %3, h1;\n\t"
            "mov.b64 %0, {l0, h0};\n\t"
            "mov.b64 %1, {l1, h1};\n\t"
            "}"
            : "+l"(a0[c]), "+l"(a1[c])
            : "r"(x), "r"(y), "r"(mult[c]));
      }
    }
  }
  u32 r = 0;
#pragma unroll
  for (int c = 0; c < PROBE_CHAINS; c++) r ^= lo32(a0[c]) ^ hi32(a0[c]) ^ lo32(a1[c]) ^ hi32(a1[c]) ^ mult[c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

// returns the number of wide multiplies one launch executes
double launch_imad_probe(u32* d_out, int blocks, u32 seed, cudaStream_t stream) {
  imad_probe_kernel<<<blocks, 256, 0, stream>>>(d_out, seed);
  return (double)blocks * 256.0 * PROBE_ITERS * PROBE_CHAINS * 2;
}

// ---------------------------------------------------------------------------------------------------
// Poseidon batch kernels.  One thread per hash.
// Addressing (in elements): input j of hash (item, chunk) at in[item*in_item_stride + chunk*in_chunk_stride + j],
// output at out[item*out_item_stride + chunk]; a plain batch is chunks_per_item = 1.
// ---------------------------------------------------------------------------------------------------
struct HashGeom {
  size_t total;            // n_items * chunks_per_item
  int chunks_per_item;
  size_t in_item_stride, in_chunk_stride, out_item_stride;
  int in_mont, out_mont;   // element format on each side (intermediate MultiHash levels stay Montgomery)
  int final_level;         // zero the output of items whose status is set
};

template <int T>
__global__ void __launch_bounds__(128) poseidon_fixed_kernel(const u32* __restrict__ in, u32* __restrict__ out,
                                                              u8* __restrict__ status, HashGeom g) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= g.total) return;
  size_t item = idx / g.chunks_per_item;
  int chunk = (int)(idx - item * g.chunks_per_item);
  const u32* src = in + (item * g.in_item_stride + (size_t)chunk * g.in_chunk_stride) * 8;
  u32 s[T][8];
  bool ok = true;
#pragma unroll
  for (int l = 0; l < 8; l++) s[0][l] = 0;
#pragma unroll
  for (int j = 1; j < T; j++) {
    u32 x[8];
    load_fr(x, src + (j - 1) * 8);
    ok = ok && fr_is_canonical(x);
    if (g.in_mont) {
#pragma unroll
      for (int l = 0; l < 8; l++) s[j][l] = x[l];
    } else {
      fr_to_mont(s[j], x);
    }
  }
  u32 h[8];
  poseidon_permute_const<T, true>(s, h);  // t = 3: partial rounds in pairs (poseidon.cuh)
  if (g.out_mont) {
    fr_canon(h);
  } else {
    u32 t[8];
    fr_from_mont(t, h);
#pragma unroll
    for (int l = 0; l < 8; l++) h[l] = t[l];
  }
  if (status) {
    if (!ok) status[item] = GCP_STATUS_NONCANONICAL;
    if (g.final_level && (!ok || status[item] != GCP_STATUS_OK)) {
#pragma unroll
      for (int l = 0; l < 8; l++) h[l] = 0;
    }
  }
  store_fr(out + (item * g.out_item_stride + chunk) * 8, h);
}

__global__ void __launch_bounds__(128) poseidon_generic_kernel(PoseidonTable tab, const u32* __restrict__ in,
                                                                u32* __restrict__ out, u8* __restrict__ status,
                                                                HashGeom g) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= g.total) return;
  size_t item = idx / g.chunks_per_item;
  int chunk = (int)(idx - item * g.chunks_per_item);
  const u32* src = in + (item * g.in_item_stride + (size_t)chunk * g.in_chunk_stride) * 8;
  u32 s[POSEIDON_MAX_T][8], n[POSEIDON_MAX_T][8];
  bool ok = true;
#pragma unroll
  for (int l = 0; l < 8; l++) s[0][l] = 0;
#pragma unroll 1
  for (int j = 1; j < tab.t; j++) {
    u32 x[8], m[8];
    load_fr(x, src + (j - 1) * 8);
    ok = ok && fr_is_canonical(x);
    if (g.in_mont) {
#pragma unroll
      for (int l = 0; l < 8; l++) m[l] = x[l];
    } else {
      fr_to_mont(m, x);
    }
#pragma unroll
    for (int l = 0; l < 8; l++) s[j][l] = m[l];
  }
  u32 h[8];
  poseidon_permute_generic(s, n, h, tab);
  if (g.out_mont) {
    fr_canon(h);
  } else {
    u32 t[8];
    fr_from_mont(t, h);
#pragma unroll
    for (int l = 0; l < 8; l++) h[l] = t[l];
  }
  if (status) {
    if (!ok) status[item] = GCP_STATUS_NONCANONICAL;
    if (g.final_level && (!ok || status[item] != GCP_STATUS_OK)) {
#pragma unroll
      for (int l = 0; l < 8; l++) h[l] = 0;
    }
  }
  store_fr(out + (item * g.out_item_stride + chunk) * 8, h);
}

// batches up to this many two-input hashes take the three-lanes-per-hash latency layout (one warp per scheduler partition up
// to 592 x 10 hashes; beyond ~4 k the throughput layout wins); GCP_B200_NO_LANES=1 keeps the one-thread layout (measurements)
constexpr size_t POSEIDON_LANES_MAX = 4096;
static bool poseidon_lanes_disabled() {
  static const bool off = getenv("GCP_B200_NO_LANES") != nullptr;
  return off;
}

cudaError_t launch_poseidon(const PoseidonTable& tab, const u32* in, u32* out, u8* status, size_t n_items,
                            int chunks_per_item, size_t in_item_stride, size_t in_chunk_stride,
                            size_t out_item_stride, int in_mont, int out_mont, int final_level,
                            cudaStream_t stream) {
  HashGeom g;
  g.total = n_items * (size_t)chunks_per_item;
  g.chunks_per_item = chunks_per_item;
  g.in_item_stride = in_item_stride;
  g.in_chunk_stride = in_chunk_stride;
  g.out_item_stride = out_item_stride;
  g.in_mont = in_mont;
  g.out_mont = out_mont;
  g.final_level = final_level;
  if (g.total == 0) return cudaSuccess;
  unsigned blocks = (unsigned)((g.total + 127) / 128);
  if (tab.t == 3 && g.total <= POSEIDON_LANES_MAX && !poseidon_lanes_disabled()) {
    // small batch: the latency layout, three lanes per hash (poseidon_lanes.cuh); 10 hashes per one-warp block
    HashGeomLanes gl{g.total, g.chunks_per_item, g.in_item_stride, g.in_chunk_stride, g.out_item_stride, g.in_mont, g.out_mont,
                     g.final_level};
    poseidon_hash2_lanes_kernel<<<(unsigned)((g.total + PL_GROUPS_PER_WARP - 1) / PL_GROUPS_PER_WARP), 32, 0, stream>>>(
        tab.C, in, out, status, gl);
    return cudaGetLastError();
  }
  if (tab.t == 3)
    poseidon_fixed_kernel<3><<<blocks, 128, 0, stream>>>(in, out, status, g);
  else if (tab.t == 4)
    poseidon_fixed_kernel<4><<<blocks, 128, 0, stream>>>(in, out, status, g);
  else
    poseidon_generic_kernel<<<blocks, 128, 0, stream>>>(tab, in, out, status, g);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------
// SMT
// ---------------------------------------------------------------------------------------------------
cudaError_t launch_smt_scan(const u32* siblings, size_t n, int n_levels, u16* lidx, u8* info, u32* hist, int sm_count,
                            cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  size_t warps_needed = n;
  unsigned blocks = (unsigned)std::min<size_t>((warps_needed + 7) / 8, (size_t)sm_count * 8);
  smt_scan_kernel<<<blocks, 256, 0, stream>>>(siblings, n, n_levels, lidx, info, hist);
  return cudaGetLastError();
}

// scan (HBM-bound) -> counting sort by path length -> leaf hashes -> path fold
cudaError_t launch_smt_verify(const SmtArgs& a, const SmtScratch& sc, int sm_count, cudaStream_t stream) {
  if (a.n == 0) return cudaSuccess;
  cudaError_t e = cudaMemsetAsync(sc.hist, 0, 256 * sizeof(u32), stream);
  if (e != cudaSuccess) return e;
  e = launch_smt_scan(a.siblings, a.n, a.n_levels, sc.lidx, sc.info, sc.hist, sm_count, stream);
  if (e != cudaSuccess) return e;
  smt_sort_prefix_kernel<<<1, 256, 0, stream>>>(sc.hist, sc.cursor, (u32)a.n);
  unsigned blocks256 = (unsigned)((a.n + 255) / 256);
  smt_sort_scatter_kernel<<<blocks256, 256, 0, stream>>>(sc.lidx, a.n, sc.cursor, sc.perm);
  unsigned blocks = (unsigned)((a.n + 127) / 128);
  const unsigned path_blocks = (unsigned)((a.n + SMT_WARPS * 32 - 1) / (SMT_WARPS * 32));
  // 4-5 resident blocks x 36 KB of staging tiles per SM: ask for the large shared-memory carve-out
  // (set on every launch: the attribute is per device, and a process may drive several GPUs)
  if (a.hasher == 1) {
    smt_leaf_kernel<1><<<blocks, 128, 0, stream>>>(a);
    cudaFuncSetAttribute(smt_path_kernel<1>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    smt_path_kernel<1><<<path_blocks, SMT_WARPS * 32, 0, stream>>>(a, sc.perm, sc.lidx, sc.info);
  } else {
    smt_leaf_kernel<0><<<blocks, 128, 0, stream>>>(a);
    cudaFuncSetAttribute(smt_path_kernel<0>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    smt_path_kernel<0><<<path_blocks, SMT_WARPS * 32, 0, stream>>>(a, sc.perm, sc.lidx, sc.info);
  }
  return cudaGetLastError();
}

// items one resident wave of a kernel covers on the current device (blocks per SM from the occupancy calculator, not a
// constant: round 1 sized chunks for 5 blocks per SM while the path kernel was resident at 4)
template <typename K>
static size_t wave_items(K kernel, int threads, size_t smem, int sm_count, int fallback_blocks) {
  int blocks = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks, kernel, threads, smem) != cudaSuccess || blocks < 1) {
    cudaGetLastError();
    blocks = fallback_blocks;
  }
  return (size_t)sm_count * blocks * threads;
}

size_t smt_path_wave_items(int sm_count) {
  cudaFuncSetAttribute(smt_path_kernel<0>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  return wave_items(smt_path_kernel<0>, SMT_WARPS * 32, 0, sm_count, 4);
}

// scratch: the verifier's (perm, lidx, info, hist, cursor) plus two accumulators per item; nullptr: thread-per-proof form
cudaError_t launch_smt_process(const SmtProcessArgs& a, const SmtScratch* sc, u32* acc_old, u32* acc_new, int sm_count,
                               cudaStream_t stream) {
  if (a.n == 0) return cudaSuccess;
  if (!sc) {
    if (a.hasher == 1)
      smt_process_kernel<1><<<(unsigned)((a.n + 127) / 128), 128, 0, stream>>>(a);
    else
      smt_process_kernel<0><<<(unsigned)((a.n + 127) / 128), 128, 0, stream>>>(a);
    return cudaGetLastError();
  }
  cudaError_t e = cudaMemsetAsync(sc->hist, 0, 256 * sizeof(u32), stream);
  if (e != cudaSuccess) return e;
  e = launch_smt_scan(a.siblings, a.n, a.n_levels, sc->lidx, sc->info, sc->hist, sm_count, stream);
  if (e != cudaSuccess) return e;
  smt_sort_prefix_kernel<<<1, 256, 0, stream>>>(sc->hist, sc->cursor, (u32)a.n);
  smt_sort_scatter_kernel<<<(unsigned)((a.n + 255) / 256), 256, 0, stream>>>(sc->lidx, a.n, sc->cursor, sc->perm);
  if (a.hasher == 1)
    smt_process_prep_kernel<1><<<(unsigned)((a.n + 127) / 128), 128, 0, stream>>>(a, sc->lidx, sc->info, acc_old, acc_new);
  else
    smt_process_prep_kernel<0><<<(unsigned)((a.n + 127) / 128), 128, 0, stream>>>(a, sc->lidx, sc->info, acc_old, acc_new);
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  // 4 resident blocks per SM (128 registers, ~80 B of spills) or 3 (148 registers, none): GCP_B200_PROC_BLOCKS, for measurements
  static const int min_blocks = [] {
    const char* env = getenv("GCP_B200_PROC_BLOCKS");
    return (env && atoi(env) == 3) ? 3 : 4;
  }();
  const unsigned grid = (unsigned)((a.n + SMT_WARPS * 32 - 1) / (SMT_WARPS * 32));
  if (a.hasher == 1) {
    cudaFuncSetAttribute(smt_process_path_kernel<4, 1>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    smt_process_path_kernel<4, 1><<<grid, SMT_WARPS * 32, 0, stream>>>(a, sc->perm, sc->lidx, acc_old, acc_new);
  } else if (min_blocks == 3) {
    cudaFuncSetAttribute(smt_process_path_kernel<3, 0>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    smt_process_path_kernel<3, 0><<<grid, SMT_WARPS * 32, 0, stream>>>(a, sc->perm, sc->lidx, acc_old, acc_new);
  } else {
    cudaFuncSetAttribute(smt_process_path_kernel<4, 0>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    smt_process_path_kernel<4, 0><<<grid, SMT_WARPS * 32, 0, stream>>>(a, sc->perm, sc->lidx, acc_old, acc_new);
  }
  return cudaGetLastError();
}

cudaError_t launch_smt_leaf_rows(const u32* keys, const u32* values, int n_values, size_t n, u32* rows, int mont,
                                 cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  const size_t total = n * (size_t)(n_values + 2);
  smt_leaf_rows_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(keys, values, n_values, n, rows, mont);
  return cudaGetLastError();
}

cudaError_t launch_smt_unpack(const u8* packed, const u64* offsets, u64 base, u64 packed_bytes, size_t n, int n_levels,
                              u32* siblings, u8* bad, int mont, cudaStream_t stream, const u8* drop_is_old0,
                              const u8* drop_fnc1) {
  if (n == 0) return cudaSuccess;
  SmtUnpackArgs a;
  a.packed = packed;
  a.offsets = offsets;
  a.base = base;
  a.packed_bytes = packed_bytes;
  a.n = n;
  a.n_levels = n_levels;
  a.siblings = siblings;
  a.bad = bad;
  a.mont = mont;
  a.drop_is_old0 = drop_fnc1 ? drop_is_old0 : nullptr;
  a.drop_fnc1 = drop_is_old0 ? drop_fnc1 : nullptr;
  smt_unpack_kernel<<<(unsigned)((n + 3) / 4), 128, 0, stream>>>(a);  // one warp per proof
  return cudaGetLastError();
}

cudaError_t launch_smt_apply_bad(const u8* bad, size_t n, u8* flags, u8* status, u32* out_roots, cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  smt_apply_bad_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(bad, n, flags, status, out_roots);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------
// ElGamal
// ---------------------------------------------------------------------------------------------------
size_t fb_table_bytes(int wbits) {  // header + entries
  return ((size_t)FB_HEADER_WORDS + ((size_t)fb_windows(wbits) << (wbits - 1)) * 24) * sizeof(u32);
}
size_t fb_small_scratch_bytes() {  // table-construction scratch, sized for the widest window
  size_t m = 0;
  for (int w = FB_MIN_WBITS; w <= FB_MAX_WBITS; w++) {
    size_t b = (size_t)fb_windows(w) * (1 + (size_t)fb_small_per_window(w)) * 32 * sizeof(u32);
    m = b > m ? b : m;
  }
  return m;
}

static inline unsigned blocks_for(size_t n, int threads) { return (unsigned)((n + threads - 1) / threads); }

// d_table: fb_table_bytes(wbits) bytes; kernels take d_table + FB_HEADER_WORDS (fb_table_entries)
cudaError_t launch_fb_table_build(const u32* d_base_xy, int base_mont, u32* d_small, u32* d_table, u32* d_flag,
                                  cudaStream_t stream, int te, int wbits) {
  if (wbits < FB_MIN_WBITS || wbits > FB_MAX_WBITS) return cudaErrorInvalidValue;
  u32* entries = d_table + FB_HEADER_WORDS;
  const size_t total = (size_t)fb_windows(wbits) << (wbits - 1);
  fb_table_bases_kernel<<<1, 32, 0, stream>>>(d_base_xy, base_mont, te, d_small, d_table, d_flag, wbits);
  fb_table_small_kernel<<<blocks_for((size_t)fb_windows(wbits) * fb_small_per_window(wbits), 64), 64, 0, stream>>>(d_small, wbits);
  fb_table_sum_kernel<<<blocks_for(total, 128), 128, 0, stream>>>(d_small, entries, wbits);
  fb_table_niels_kernel<<<blocks_for((total + BATCH_INV - 1) / BATCH_INV, 128), 128, 0, stream>>>(entries, total);
  return cudaGetLastError();
}

cudaError_t upload_generator(u32* d_xy, cudaStream_t stream) {
  static const u32 g[16] = {0xfe553f9fu, 0xf1f9195au, 0xe6f2a277u, 0x377c749au, 0xc199e94cu, 0x8a4eb7a4u, 0x6ce19d35u, 0x1561ff83u,
                            0x872d7d8bu, 0x4b3c257au, 0xb9e13377u, 0xfce0051fu, 0xd16bf9edu, 0x25572e1cu, 0xf7a0b249u, 0x25797203u};
  return cudaMemcpyAsync(d_xy, g, sizeof(g), cudaMemcpyHostToDevice, stream);
}

cudaError_t launch_fixed_base_mul(const u32* tabG, const u32* scalars, size_t n, u32* out_xyz, u8* status, int mont,
                                  cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  fixed_base_mul_kernel<<<blocks_for(n, 128), 128, 0, stream>>>(tabG, scalars, n, out_xyz, status, mont);
  return cudaGetLastError();
}

cudaError_t launch_encrypt_shared(const u32* tabG, const u32* tabPK, const u32* pk_flag, const u32* ks, const u32* ms,
                                  size_t n, u32* out_xyz, u8* status, int mont, cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  encrypt_shared_kernel<<<blocks_for(2 * n, 128), 128, 0, stream>>>(tabG, tabPK, pk_flag, ks, ms, n, out_xyz, status, mont);
  return cudaGetLastError();
}

// points per thread of normalize_kernel: chunkplan.h (host logic, unit-tested without a GPU)
static size_t points_per_thread(size_t n_points) {
  static const int sms = []() {
    int dev = 0, n = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    return n;
  }();
  return normalize_points_per_thread(n_points, sms, BATCH_INV, BATCH_INV_MAX);
}

cudaError_t launch_normalize(const u32* xyz, size_t n_points, u32* out, u8* status, int pts_per_item, int mont,
                             cudaStream_t stream, int xyz_words, int te) {
  if (n_points == 0) return cudaSuccess;
  const size_t per = points_per_thread(n_points);
  size_t threads = (n_points + per - 1) / per;
  normalize_kernel<<<blocks_for(threads, 128), 128, 0, stream>>>(xyz, n_points, out, status, pts_per_item, mont, xyz_words, te,
                                                                 (int)per);
  return cudaGetLastError();
}

cudaError_t launch_ct_add(const u32* a, const u32* b, size_t n, u32* out_xyz, u8* status, int mont, cudaStream_t stream,
                          int te) {
  if (n == 0) return cudaSuccess;
  ct_add_kernel<<<blocks_for(2 * n, 128), 128, 0, stream>>>(a, b, n, out_xyz, status, mont, te);
  return cudaGetLastError();
}

cudaError_t launch_ct_neg(const u32* a, size_t n, u32* out, u8* status, cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  ct_neg_kernel<<<blocks_for(2 * n, 128), 128, 0, stream>>>(a, 2 * n, out, status);
  return cudaGetLastError();
}

cudaError_t launch_ct_is_equal(const u32* a, const u32* b, size_t n, u8* flags, u8* status, cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  ct_is_equal_kernel<<<blocks_for(n, 128), 128, 0, stream>>>(a, b, n, flags, status);
  return cudaGetLastError();
}

cudaError_t launch_ct_select(const u8* sel, const u32* i1, const u32* i2, size_t n, u32* out, u8* status, cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  ct_select_kernel<<<blocks_for(n * 8, 256), 256, 0, stream>>>(sel, i1, i2, n, out, status);
  return cudaGetLastError();
}

int tally_max_blocks(size_t n_ballots, int n_fields, int sm_count) {
  int rows = (TALLY_THREADS / 2) / n_fields;  // same for both layouts: 128 / (2 n_fields) == 64 / n_fields
  size_t need = (n_ballots + rows - 1) / rows;
  size_t cap = (size_t)sm_count * 8;
  return (int)std::max<size_t>(1, std::min(need, cap));
}

// partials: n_blocks x (n_fields*2) x 32 words; bad_count: n_fields words (zeroed here); out_xyz: n_fields*2 x 24 words
cudaError_t launch_tally(const u32* ct, size_t n_ballots, int n_fields, int n_blocks, u32* partials, u32* bad_count,
                         u32* out_xyz, u8* status, int mont, cudaStream_t stream, int te) {
  cudaError_t e = cudaMemsetAsync(bad_count, 0, sizeof(u32) * n_fields, stream);
  if (e != cudaSuccess) return e;
  tally_partial_kernel<<<n_blocks, TALLY_THREADS, TALLY_THREADS * 32 * sizeof(u32), stream>>>(ct, n_ballots, n_fields, partials,
                                                                                              bad_count, mont, te);
  const int cols = n_fields * 2;
  tally_final_kernel<<<cols, TALLY_THREADS, 0, stream>>>(partials, n_blocks, cols, out_xyz, bad_count, status);
  return cudaGetLastError();
}

cudaError_t launch_encrypt_tally(const u32* tabG, const u32* tabPK, const u32* ks, const u32* ms, const u8* mask,
                                 size_t n_ballots, int n_fields, int n_blocks, u32* partials, u32* bad_count, u32* out_xyz, u8* status, int mont,
                                 cudaStream_t stream, int m_words) {
  cudaError_t e = cudaMemsetAsync(bad_count, 0, sizeof(u32) * n_fields, stream);
  if (e != cudaSuccess) return e;
  if (FB_BUFS > 2) {  // static staging buffers + the dynamic reduction tile exceed 48 KB: opt in (once per process)
    static const bool opted = []() {
      const int dyn = TALLY_THREADS * 32 * sizeof(u32);
      cudaFuncSetAttribute(encrypt_tally_partial_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn);
      cudaFuncSetAttribute(encrypt_tally_partial_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn);
      return true;
    }();
    (void)opted;
  }
  if (m_words == 2)
    encrypt_tally_partial_kernel<2><<<n_blocks, TALLY_THREADS, TALLY_THREADS * 32 * sizeof(u32), stream>>>(
        tabG, tabPK, ks, ms, mask, n_ballots, n_fields, partials, bad_count, mont);
  else
    encrypt_tally_partial_kernel<8><<<n_blocks, TALLY_THREADS, TALLY_THREADS * 32 * sizeof(u32), stream>>>(
        tabG, tabPK, ks, ms, mask, n_ballots, n_fields, partials, bad_count, mont);
  const int cols = n_fields * 2;
  tally_final_kernel<<<cols, TALLY_THREADS, 0, stream>>>(partials, n_blocks, cols, out_xyz, bad_count, status);
  return cudaGetLastError();
}

// Final status of a tally, on the device: status[f] = the fold's own status (have_final) else 0, then the first non-zero
// per-chunk status of the field, then GCP_STATUS_OFF_CURVE for every field when the cached public key failed
// AssertIsOnCurve (elgamal/encrypt.go:49; pk_flag = nullptr for a plain tally).  A field with a status carries no result:
// its ciphertext is zeroed.
__global__ void tally_status_merge_kernel(const u8* __restrict__ part_status, int n_chunks, int n_fields, int have_final,
                                          const u32* __restrict__ pk_flag, u32* __restrict__ ct, u8* __restrict__ status) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= n_fields) return;
  u8 st = have_final ? status[f] : (u8)GCP_STATUS_OK;
  for (int c = 0; c < n_chunks && st == GCP_STATUS_OK; c++) st = part_status[(size_t)c * n_fields + f];
  if (pk_flag && *pk_flag == 0) st = GCP_STATUS_OFF_CURVE;
  status[f] = st;
  if (st != GCP_STATUS_OK) {
#pragma unroll
    for (int l = 0; l < 32; l++) ct[f * 32 + l] = 0;
  }
}

cudaError_t launch_tally_status_merge(const u8* part_status, int n_chunks, int n_fields, int have_final, const u32* pk_flag,
                                      u32* ct, u8* status, cudaStream_t stream) {
  tally_status_merge_kernel<<<1, 64, 0, stream>>>(part_status, n_chunks, n_fields, have_final, pk_flag, ct, status);
  return cudaGetLastError();
}

cudaError_t launch_keccak_address(const u8* in, size_t n, u8* out, cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  keccak_address_kernel<<<blocks_for(n, 256), 256, 0, stream>>>(in, n, out);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------
// Variable-base family (varbase.cuh): pre pass -> [Poseidon batch] -> window kernel(s) -> post pass.
// `scratch` holds the per-item intermediates (varbase_scratch_bytes), arrays laid out one after the other.
// ---------------------------------------------------------------------------------------------------
// resident blocks per SM the window kernel is compiled for: 4 (128 registers) or 5 (96 registers, 28 B of spills);
// GCP_B200_VB_BLOCKS selects, for measurements
static int varbase_min_blocks() {
  static const int v = [] {
    const char* e = getenv("GCP_B200_VB_BLOCKS");
    return (e && atoi(e) == 4) ? 4 : ((e && atoi(e) == 5) ? 5 : 4);
  }();
  return v;
}

// The window kernel: the flat interpreter (varbase_flat.cuh) unless GCP_B200_VB_IMPL selects, for measurements, `interp`
// (the step-structured interpreter of varbase.cuh) or `reg` (the register-resident kernel with out-of-line multipliers,
// varbase_reg.cuh)
static int varbase_impl() {  // 0 flat, 1 interp, 2 reg
  static const int v = [] {
    const char* e = getenv("GCP_B200_VB_IMPL");
    return (e && e[0] == 'r') ? 2 : ((e && e[0] == 'i') ? 1 : 0);
  }();
  return v;
}
static bool varbase_reg_form() { return varbase_impl() == 2; }

static cudaError_t launch_varbase_window(const u32* bases, const u32* s0, const u32* s1, int n_bases, size_t n,
                                         const u8* status, u32* table, u32* out, cudaStream_t stream) {
  VarbaseArgs a;
  a.bases = bases;
  a.scalars[0] = s0;
  a.scalars[1] = s1 ? s1 : s0;
  a.n_bases = n_bases;
  a.n = n;
  a.status = status;
  a.table = table;
  a.out = out;
  const size_t smem = (size_t)VB_SLOTS * 2 * VB_THREADS * sizeof(uint4);
  if (varbase_reg_form()) {
    varbase_window_reg_kernel<4><<<blocks_for(n, VB_THREADS), VB_THREADS, 0, stream>>>(a);
  } else if (varbase_impl() == 0) {
    cudaFuncSetAttribute(varbase_flat_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)VF_SMEM_BYTES);
    cudaFuncSetAttribute(varbase_flat_kernel<4>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    varbase_flat_kernel<4><<<blocks_for(n, VB_THREADS), VB_THREADS, VF_SMEM_BYTES, stream>>>(a);
  } else if (varbase_min_blocks() == 5) {
    cudaFuncSetAttribute(varbase_window_kernel<5>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    varbase_window_kernel<5><<<blocks_for(n, VB_THREADS), VB_THREADS, smem, stream>>>(a);
  } else {
    cudaFuncSetAttribute(varbase_window_kernel<4>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    varbase_window_kernel<4><<<blocks_for(n, VB_THREADS), VB_THREADS, smem, stream>>>(a);
  }
  return cudaGetLastError();
}

size_t varbase_wave_items(int sm_count) {
  const size_t smem = (size_t)VB_SLOTS * 2 * VB_THREADS * sizeof(uint4);
  if (varbase_reg_form()) return wave_items(varbase_window_reg_kernel<4>, VB_THREADS, 0, sm_count, 4);
  if (varbase_impl() == 0) {
    cudaFuncSetAttribute(varbase_flat_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)VF_SMEM_BYTES);
    cudaFuncSetAttribute(varbase_flat_kernel<4>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    return wave_items(varbase_flat_kernel<4>, VB_THREADS, VF_SMEM_BYTES, sm_count, 4);
  }
  if (varbase_min_blocks() == 5) {
    cudaFuncSetAttribute(varbase_window_kernel<5>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    return wave_items(varbase_window_kernel<5>, VB_THREADS, smem, sm_count, 5);
  }
  cudaFuncSetAttribute(varbase_window_kernel<4>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  return wave_items(varbase_window_kernel<4>, VB_THREADS, smem, sm_count, 4);
}

// words of scratch per item: kind 0 Encrypt with per-item keys, 1 AssertDecrypt, 2 DecryptionProof.Verify, 3 EdDSA
static constexpr size_t VB_SCRATCH_WORDS[4] = {32 + 8 + 256 + 32, 32 + 8 + 32 + 256 + 32,
                                               96 + 32 + 32 + 64 + 8 + 8 + 512 + 32 + 32, 40 + 32 + 32 + 8 + 256 + 32};
size_t varbase_scratch_bytes(int kind, size_t n) { return VB_SCRATCH_WORDS[kind] * sizeof(u32) * n; }

// curve.ScalarMul over a batch: out (n x 32 words, extended) = [s]P (+ [s2]P2).  scratch: (32 + 8 + 256) * n_bases words per item
size_t scalar_mul_scratch_bytes(size_t n, int n_bases) { return (size_t)(296 * n_bases) * sizeof(u32) * n; }
cudaError_t launch_scalar_mul(const u32* points, const u32* scalars, const u32* points2, const u32* scalars2, size_t n,
                              u32* out, u8* status, int mont, u32* scratch, int* n_launches, cudaStream_t stream, int te) {
  if (n == 0) return cudaSuccess;
  const int nb = points2 ? 2 : 1;
  u32 *bases = scratch, *k0 = bases + n * 32 * nb, *k1 = k0 + n * 8, *table = k0 + n * 8 * nb;
  scalar_mul_pre_kernel<<<blocks_for(n, 128), 128, 0, stream>>>(points, scalars, points2, scalars2, n, mont, te, status, bases, k0, k1);
  cudaError_t e = launch_varbase_window(bases, k0, nb == 2 ? k1 : nullptr, nb, n, status, table, out, stream);
  if (e != cudaSuccess) return e;
  *n_launches += 2;
  return cudaGetLastError();
}

cudaError_t launch_encrypt_per_key(const u32* tabG, const u32* pks, const u32* ks, const u32* ms, size_t n, u32* out_xyz,
                                   u8* status, int mont, u32* scratch, int* n_launches, cudaStream_t stream, int te) {
  if (n == 0) return cudaSuccess;
  u32 *bases = scratch, *kint = bases + n * 32, *table = kint + n * 8, *res = table + n * 256;
  encrypt_per_key_pre_kernel<<<blocks_for(n, 128), 128, 0, stream>>>(pks, ks, ms, n, mont, te, status, bases, kint);
  cudaError_t e = launch_varbase_window(bases, kint, nullptr, 1, n, status, table, res, stream);
  if (e != cudaSuccess) return e;
  encrypt_per_key_finish_kernel<<<blocks_for(2 * n, 128), 128, 0, stream>>>(tabG, ks, ms, res, status, n, mont, out_xyz);
  *n_launches += 3;
  return cudaGetLastError();
}

cudaError_t launch_assert_decrypt(const u32* tabG, const u32* cts, const u32* privs, const u32* msgs, size_t n, u8* flags,
                                  u8* status, int mont, u32* scratch, int* n_launches, cudaStream_t stream, int te) {
  if (n == 0) return cudaSuccess;
  u32 *bases = scratch, *kint = bases + n * 32, *rhs = kint + n * 8, *table = rhs + n * 32, *res = table + n * 256;
  assert_decrypt_pre_kernel<<<blocks_for(n, 128), 128, 0, stream>>>(tabG, cts, privs, msgs, n, mont, te, status, bases, kint, rhs);
  cudaError_t e = launch_varbase_window(bases, kint, nullptr, 1, n, status, table, res, stream);
  if (e != cudaSuccess) return e;
  ext_compare_kernel<<<blocks_for(n, 128), 128, 0, stream>>>(res, rhs, status, n, flags);
  *n_launches += 3;
  return cudaGetLastError();
}

cudaError_t launch_decryption_proof(const u32* tabG, const PoseidonTable& tab13, const u32* pks, const u32* cts,
                                    const u32* msgs, const u32* a1s, const u32* a2s, const u32* zs, size_t n, u8* flags,
                                    u8* status, int mont, u32* scratch, int* n_launches, cudaStream_t stream, int te) {
  if (n == 0) return cudaSuccess;
  u32 *hin = scratch, *zg = hin + n * 96, *base_pk = zg + n * 32, *base2 = base_pk + n * 32, *zint = base2 + n * 64;
  u32 *e_int = zint + n * 8, *table = e_int + n * 8, *res1 = table + n * 512, *res2 = res1 + n * 32;
  decryption_proof_pre_kernel<<<blocks_for(n, 128), 128, 0, stream>>>(tabG, pks, cts, msgs, a1s, a2s, zs, n, mont, te, status,
                                                                      hin, zg, base_pk, base2, zint);
  // e = MultiHash(PK, PK, C1, D, A1, A2): 12 inputs = one Hash with t = 13; Montgomery in, integer out
  cudaError_t e = launch_poseidon(tab13, hin, e_int, nullptr, n, 1, 12, 0, 1, 1, 0, 0, stream);
  if (e != cudaSuccess) return e;
  e = launch_varbase_window(base_pk, e_int, nullptr, 1, n, status, table, res1, stream);
  if (e != cudaSuccess) return e;
  e = launch_varbase_window(base2, zint, e_int, 2, n, status, table, res2, stream);
  if (e != cudaSuccess) return e;
  decryption_proof_post_kernel<<<blocks_for(n, 128), 128, 0, stream>>>(a1s, a2s, zg, res1, res2, status, n, mont, te, flags);
  *n_launches += 5;
  return cudaGetLastError();
}

cudaError_t launch_eddsa_verify(const u32* tabG, const PoseidonTable& tab6, const u32* pub_a, const u32* sig_r,
                                const u32* sig_s, const u32* msgs, size_t n, u8* flags, u8* status, int mont, u32* scratch,
                                int* n_launches, cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  u32 *hin = scratch, *left = hin + n * 40, *base = left + n * 32, *h_int = base + n * 32, *table = h_int + n * 8;
  u32* res = table + n * 256;
  eddsa_pre_kernel<<<blocks_for(n, 128), 128, 0, stream>>>(tabG, pub_a, sig_r, sig_s, msgs, n, mont, status, hin, left, base);
  cudaError_t e = launch_poseidon(tab6, hin, h_int, nullptr, n, 1, 5, 0, 1, 1, 0, 0, stream);
  if (e != cudaSuccess) return e;
  e = launch_varbase_window(base, h_int, nullptr, 1, n, status, table, res, stream);
  if (e != cudaSuccess) return e;
  eddsa_post_kernel<<<blocks_for(n, 128), 128, 0, stream>>>(sig_r, left, res, status, n, mont, flags);
  *n_launches += 4;
  return cudaGetLastError();
}

cudaError_t launch_te_rte(const u32* in, size_t n_points, u32* out, u8* status, int to_rte, cudaStream_t stream) {
  if (n_points == 0) return cudaSuccess;
  te_rte_kernel<<<blocks_for(n_points, 256), 256, 0, stream>>>(in, n_points, out, status, to_rte);
  return cudaGetLastError();
}

cudaError_t upload_mimc7_constants(const u32* d_mont, cudaStream_t stream) {
  return cudaMemcpyToSymbolAsync(c_mimc7, d_mont, sizeof(u32) * MIMC7_ROUNDS * 8, 0, cudaMemcpyDeviceToDevice, stream);
}

cudaError_t launch_mimc7(const u32* in, int len, size_t n, u32* out, u8* status, int mont, cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  mimc7_kernel<<<blocks_for(n, 128), 128, 0, stream>>>(in, len, n, out, status, mont);
  return cudaGetLastError();
}

cudaError_t launch_poseidon2_hash(const u32* keys, const u32* in, int len, size_t n, u32* out, u8* status, int mont,
                                  cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  poseidon2_hash_kernel<<<blocks_for(n, 128), 128, 0, stream>>>(keys, in, len, n, out, status, mont);
  return cudaGetLastError();
}

cudaError_t launch_poseidon2_permutation(const u32* keys, const u32* in, size_t n, u32* out, u8* status, int mont,
                                         cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  poseidon2_permutation_kernel<<<blocks_for(n, 128), 128, 0, stream>>>(keys, in, n, out, status, mont);
  return cudaGetLastError();
}

}  // namespace gcp
