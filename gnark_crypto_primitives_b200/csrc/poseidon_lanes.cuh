// Small-batch Hash2: three lanes per hash, one per state element.
//
// BASELINE config 1 is a batch of 1024 two-input hashes.  With one thread per hash that is 32 warps on 32 scheduler
// partitions, each issuing one dependent stream of 61 k wide multiplies: the call is bound by the latency of ONE hash
// (~270 us of a 310 us call), not by the machine.  Here lane r of a 3-lane group holds state element r of the group's
// hash.  Every round has the same shape on every lane, so the group never diverges:
//   y_r = s_r^5 + c_r           (all lanes in a full round; lane 0 only in a partial round: the others keep y_r = s_r, c_r = 0)
//   gather y_0, y_1, y_2        (warp shuffles inside the group, 16 words per lane)
//   s_r = <W_r, y>              (one lazy three-term dot row)
// with W = M or P in a full round (poseidon.go:213-224) and, in a partial round (poseidon.go:152-166),
//   W_0 = (S_0, S_1, S_2),  W_k = (S_{t+k-1}, e_k)   i.e. s_k + y_0 * S_{t+k-1} written as a dot row with the constants 1 and 0.
// Critical path per round: one S-box (328 wide multiplies) + one dot row (256) = 584 against 840 per partial round and
// 1 752 per full round of the one-thread schedule: 38 k instead of 61 k dependent wide multiplies per hash, at three
// times the lanes - a latency layout, used only below 4 096 hashes per call (launch_poseidon); the throughput layout
// (poseidon_fixed_kernel) stays the path for everything else.  Tables: the t = 3 block of the constant table, staged in
// shared memory because the lanes of a warp read three different rows at once (constant memory serialises that).
#pragma once
#include "poseidon.cuh"

namespace gcp {

constexpr int PL_GROUPS_PER_WARP = 10;  // 30 lanes busy, lanes 30 and 31 shadow group 9 (their results are dropped)

struct HashGeomLanes {
  size_t total;
  int chunks_per_item;
  size_t in_item_stride, in_chunk_stride, out_item_stride;
  int in_mont, out_mont, final_level;
};

__global__ void __launch_bounds__(32) poseidon_hash2_lanes_kernel(const u32* __restrict__ tab3, const u32* __restrict__ in,
                                                                 u32* __restrict__ out, u8* __restrict__ status,
                                                                 HashGeomLanes g) {
  constexpr int T = 3, RP = 57;
  __shared__ u32 tab[(POS3_ELEMS + 2) * 8];  // C | S | M | P | ONE | ZERO
  for (int i = threadIdx.x; i < POS3_ELEMS * 8; i += 32) tab[i] = tab3[i];
  if (threadIdx.x < 8) {
    tab[POS3_ELEMS * 8 + threadIdx.x] = FR_ONE[threadIdx.x];
    tab[(POS3_ELEMS + 1) * 8 + threadIdx.x] = 0;
  }
  __syncwarp();
  const u32* C = tab;
  const u32* S = C + (8 * T + RP) * 8;
  const u32* M = S + (2 * T - 1) * RP * 8;
  const u32* P = M + T * T * 8;
  const u32* ONE = tab + POS3_ELEMS * 8;
  const u32* ZERO = ONE + 8;
  const u32 P2[8] = GCP_2P_LIMBS;

  const int lane = threadIdx.x;
  const int grp = min(lane / 3, PL_GROUPS_PER_WARP - 1);
  const int r = lane < 30 ? lane % 3 : lane - 29;     // lanes 30, 31 shadow elements 1, 2 of group 9
  const int base = grp * 3;
  const size_t idx = (size_t)blockIdx.x * PL_GROUPS_PER_WARP + grp;
  const bool valid = idx < g.total;
  const size_t item = valid ? idx / g.chunks_per_item : 0;
  const int chunk = valid ? (int)(idx - item * g.chunks_per_item) : 0;

  u32 s[8];
  bool ok = true;
  fr_set_zero(s);
  if (valid && r > 0) {
    u32 x[8];
    load_fr(x, in + (item * g.in_item_stride + (size_t)chunk * g.in_chunk_stride + (r - 1)) * 8);
    ok = fr_is_canonical(x);
    if (g.in_mont)
      fr_copy(s, x);
    else
      fr_to_mont(s, x);
  }
  {
    u32 c[8];
    load_const(c, C + r * 8);
    fr_add(s, s, c);  // ark(C, 0)
  }
#pragma unroll 1
  for (int rnd = 0; rnd < 8 + RP; rnd++) {
    const bool full = (rnd < 4) || (rnd >= 4 + RP);
    const bool last = (rnd == 7 + RP);
    const u32* crow = (rnd < 4) ? C + (rnd + 1) * T * 8 : (full ? C + ((rnd - RP + 1) * T + RP) * 8 : C + (5 * T + (rnd - 4)) * 8);
    // S-box where this lane's element takes one, then the round constant
    u32 y[8];
    fr_copy(y, s);
    if (full || r == 0) {
      fr_sqr(y, y);
      fr_sqr(y, y);
      fr_mul(y, y, s);
      if (!last) {
        u32 c[8];
        load_const(c, crow + (full ? r : 0) * 8);
        fr_add(y, y, c);
      }
    }
    // everyone gets the three post-S-box elements
    u32 v[3][8];
#pragma unroll
    for (int j = 0; j < 3; j++)
#pragma unroll
      for (int l = 0; l < 8; l++) v[j][l] = __shfl_sync(0xffffffffu, y[l], base + j);
    // this lane's coefficient row
    const u32* w0;
    const u32* w1;
    const u32* w2;
    if (full) {
      const u32* coef = (rnd == 3) ? P : M;
      const int col = last ? 0 : r;            // the digest is column 0 of M (poseidon.go:180)
      w0 = coef + (0 * T + col) * 8;
      w1 = coef + (1 * T + col) * 8;
      w2 = coef + (2 * T + col) * 8;
    } else {
      const u32* srow = S + (2 * T - 1) * (rnd - 4) * 8;
      w0 = r == 0 ? srow : srow + (T + r - 1) * 8;
      w1 = r == 0 ? srow + 8 : (r == 1 ? ONE : ZERO);
      w2 = r == 0 ? srow + 16 : (r == 2 ? ONE : ZERO);
    }
    Wide w;
    wide_zero(w);
    u32 c = 0;
#define GCP_LANE_ROW(I)          \
  mac_row<I>(w, v[0], w0[I]);    \
  mac_row<I>(w, v[1], w1[I]);    \
  mac_row<I>(w, v[2], w2[I]);    \
  redc_row<I>(w, c);
    GCP_LANE_ROW(0) GCP_LANE_ROW(1) GCP_LANE_ROW(2) GCP_LANE_ROW(3) GCP_LANE_ROW(4) GCP_LANE_ROW(5) GCP_LANE_ROW(6) GCP_LANE_ROW(7)
#undef GCP_LANE_ROW
    wide_redc_finish(w, c, s);
    cond_sub(s, P2);
  }
  // all inputs of the hash canonical?
  const unsigned bad = __ballot_sync(0xffffffffu, !ok);
  const bool hash_ok = ((bad >> base) & 7u) == 0;
  if (valid && lane < 30 && r == 0) {
    u32 h[8];
    fr_copy(h, s);
    if (g.out_mont) {
      fr_canon(h);
    } else {
      u32 t[8];
      fr_from_mont(t, h);
      fr_copy(h, t);
    }
    if (status) {
      if (!hash_ok) status[item] = GCP_STATUS_NONCANONICAL;
      if (g.final_level && (!hash_ok || status[item] != GCP_STATUS_OK)) fr_set_zero(h);
    }
    store_fr(out + (item * g.out_item_stride + chunk) * 8, h);
  }
}

}  // namespace gcp
