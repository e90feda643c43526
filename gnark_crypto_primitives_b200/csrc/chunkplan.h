// Chunk schedule of the chunked host-buffer pipelines (pure host C++: unit-tested without a GPU, tests/cpp/test_host_logic.cpp).
#pragma once
#include <algorithm>
#include <cstddef>
#include <cstdlib>

// Chunk schedule of the host-buffer pipelines whose kernels are resident-wave shaped (smt_path_kernel,
// varbase_window_kernel: one thread per item, registers set the residency).  The first chunk is ONE wave, so the only
// copy no kernel hides is short (at 2^18 census-like proofs the old ~1 GB first chunk left 20 ms of copy exposed against
// 51 ms of compute); for the SMT row pipelines the second chunk is one wave as well, so that its copy is over before the
// first chunk's kernels are (a 2.46-wave second chunk of census-like proofs takes 18.7 ms to arrive while the first wave
// computes for 15 ms: census-like rows 4.18 -> 4.43 M proofs/s; the per-item proof pipelines, whose copies are short,
// measured better without it); later chunks are `cap_waves` whole waves, and a remainder shorter than a wave joins the
// chunk before it instead of running as a thin launch of its own.  GCP_B200_SMT_CHUNK overrides the size (tests).
namespace gcp {

struct ChunkPlan {
  size_t wave, cap, n, off = 0, taken = 0;
  bool forced = false;
  bool second_single;  // the second chunk is one wave too (pipelines whose copy per wave is long against its compute: SMT rows)
  ChunkPlan(size_t n_items, size_t wave_items, size_t cap_items, bool second_chunk_single = false)
      : wave(std::max<size_t>(1, wave_items)), n(n_items), second_single(second_chunk_single) {
    cap = std::max(wave, cap_items - cap_items % wave);
    if (const char* env = getenv("GCP_B200_SMT_CHUNK")) {
      long v = atol(env);
      if (v > 0) {
        cap = wave = (size_t)v;
        forced = true;
      }
    }
  }
  size_t largest() const { return std::min(n, cap + wave); }
  size_t next() {  // items of the next chunk (0: done); advances
    const size_t left = n - off;
    if (left == 0) return 0;
    static const size_t first_div = [] {
      const char* env = getenv("GCP_B200_FIRST_DIV");
      long v = env ? atol(env) : 1;
      return (size_t)(v >= 1 && v <= 64 ? v : 1);
    }();
    size_t take = off == 0 ? std::min(left, std::max<size_t>(1, wave / first_div))
                           : (taken == 1 && second_single ? std::min(left, wave) : std::min(left, cap));
    if (!forced && left - take < wave) take = left <= cap + wave ? left : take;
    off += take;
    taken++;
    return take;
  }
};

// Points per thread of normalize_kernel (= points that share one Fermat inversion, ~325 multiplications): per_min until the
// batch is big enough for ~640 threads on every SM at that ratio, then as many as keep that many threads busy, up to
// per_max.  GCP_B200_NORM_PER forces a ratio in [1, per_max] (measurements, tests).
inline size_t normalize_points_per_thread(size_t n_points, int sm_count, size_t per_min, size_t per_max) {
  const size_t fill = (size_t)std::max(1, sm_count) * 640;
  size_t per = (n_points + fill - 1) / fill;
  per = per < per_min ? per_min : (per > per_max ? per_max : per);
  if (const char* e = getenv("GCP_B200_NORM_PER")) {
    long v = atol(e);
    if (v >= 1 && (size_t)v <= per_max) per = (size_t)v;
  }
  return per;
}

}  // namespace gcp
