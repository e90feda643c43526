// CPU-only unit test of the host-side logic under the C ABI that needs no GPU: the chunk schedule of the host-buffer
// pipelines (csrc/chunkplan.h) and the process-wide copy pool (csrc/hostcopy.h).  Built and run by tests/test_host_logic.py.
#include <cstdio>
#include <cstring>
#include <random>
#include <thread>
#include <vector>

#include "chunkplan.h"
#include "hostcopy.h"

#define CHECK(cond)                                                      \
  do {                                                                   \
    if (!(cond)) {                                                       \
      fprintf(stderr, "%s:%d: check failed: %s\n", __FILE__, __LINE__, #cond); \
      return 1;                                                          \
    }                                                                    \
  } while (0)

static int test_chunk_plan() {
  const size_t waves[] = {1, 7, 128, 75776, 94720};
  const size_t caps[] = {1, 100, 209715, 262144, 1u << 20};
  const size_t ns[] = {0, 1, 2, 127, 128, 129, 75775, 75776, 75777, 151552, 200000, 262144, 524288, 1u << 20, (1u << 20) + 3};
  for (int second = 0; second < 2; second++)
  for (size_t wave : waves)
    for (size_t cap : caps)
      for (size_t n : ns) {
        gcp::ChunkPlan plan(n, wave, cap, second != 0);
        const size_t largest = plan.largest();
        const size_t cap_eff = std::max(wave, cap - cap % wave);
        std::vector<size_t> sizes;
        size_t total = 0;
        for (size_t c = plan.next(); c != 0; c = plan.next()) {
          sizes.push_back(c);
          total += c;
          CHECK(sizes.size() <= n + 1);
        }
        CHECK(total == n);                                        // every item exactly once
        CHECK(plan.next() == 0);                                  // and the plan stays finished
        for (size_t i = 0; i < sizes.size(); i++) {
          CHECK(sizes[i] >= 1 && sizes[i] <= largest);            // the slots are sized for largest()
          if (i == 0 && sizes.size() > 1) CHECK(sizes[0] == wave);                    // one wave first
          if (i == 1 && sizes.size() > 2 && second) CHECK(sizes[1] == wave);          // and one more where asked: its copy hides under the first
          if (i > (second ? 1u : 0u) && i + 1 < sizes.size()) CHECK(sizes[i] == cap_eff);   // whole waves in the middle
          if (i + 1 == sizes.size() && sizes.size() > 1) CHECK(sizes[i] >= wave);    // no thin launch at the end
        }
        if (n <= wave) CHECK(sizes.size() == (n ? 1u : 0u));
      }
  return 0;
}

static int test_copy_pool() {
  // without a pool: plain memcpy
  {
    std::vector<unsigned char> a(5 << 20, 3), b(5 << 20, 0);
    gcp::CopyPool::copy(b.data(), a.data(), a.size());
    CHECK(a == b && gcp::CopyPool::workers() == 0);
  }
  gcp::CopyPool::acquire();
  gcp::CopyPool::acquire();  // two contexts
  CHECK(gcp::CopyPool::workers() >= 1 && gcp::CopyPool::workers() <= 16);
  std::vector<std::thread> callers;
  std::vector<int> bad(6, 0);
  for (int t = 0; t < 6; t++)
    callers.emplace_back([t, &bad] {
      std::mt19937_64 rng(1234 + t);
      for (int it = 0; it < 12; it++) {
        const size_t bytes = (size_t)(rng() % (40u << 20)) + 1;   // 1 B .. 40 MB: below and above the piece size
        std::vector<unsigned char> src(bytes), dst(bytes, 0);
        for (size_t i = 0; i < bytes; i += 4093) src[i] = (unsigned char)(rng() >> 7);
        src[bytes - 1] = (unsigned char)t;
        gcp::CopyPool::copy(dst.data(), src.data(), bytes);
        if (memcmp(src.data(), dst.data(), bytes) != 0) bad[t]++;
      }
    });
  for (auto& c : callers) c.join();
  for (int t = 0; t < 6; t++) CHECK(bad[t] == 0);
  gcp::CopyPool::release();
  CHECK(gcp::CopyPool::workers() >= 1);  // one context still holds it
  gcp::CopyPool::release();
  CHECK(gcp::CopyPool::workers() == 0);  // the last release joined the workers
  gcp::CopyPool::acquire();              // and it can come back
  CHECK(gcp::CopyPool::workers() >= 1);
  gcp::CopyPool::release();
  return 0;
}

// normalize_kernel's points per Fermat inversion (chunkplan.h): 32 on small batches, then what keeps ~640 threads per SM
// busy, at most 128; the forced ratio of the measurements and the parity tests
static int test_normalize_points_per_thread() {
  unsetenv("GCP_B200_NORM_PER");
  const int sms = 148;
  const size_t fill = (size_t)sms * 640;
  CHECK(gcp::normalize_points_per_thread(0, sms, 32, 128) == 32);
  CHECK(gcp::normalize_points_per_thread(1, sms, 32, 128) == 32);
  CHECK(gcp::normalize_points_per_thread(32 * fill, sms, 32, 128) == 32);
  CHECK(gcp::normalize_points_per_thread(32 * fill + 1, sms, 32, 128) == 33);
  CHECK(gcp::normalize_points_per_thread((size_t)1 << 23, sms, 32, 128) == 89);  // 2^22 ciphertexts: 88.6 -> 89
  CHECK(gcp::normalize_points_per_thread(128 * fill, sms, 32, 128) == 128);
  CHECK(gcp::normalize_points_per_thread((size_t)1 << 30, sms, 32, 128) == 128);
  CHECK(gcp::normalize_points_per_thread(1000, 0, 32, 128) == 32);               // a silly SM count does not divide by 0
  for (size_t n : {(size_t)1, (size_t)4096, (size_t)1 << 22, (size_t)1 << 27}) {
    const size_t per = gcp::normalize_points_per_thread(n, sms, 32, 128);
    CHECK(per >= 32 && per <= 128);
    CHECK(per * ((n + per - 1) / per) >= n);  // the launch covers every point
  }
  setenv("GCP_B200_NORM_PER", "7", 1);
  CHECK(gcp::normalize_points_per_thread((size_t)1 << 23, sms, 32, 128) == 7);
  setenv("GCP_B200_NORM_PER", "128", 1);
  CHECK(gcp::normalize_points_per_thread(10, sms, 32, 128) == 128);
  setenv("GCP_B200_NORM_PER", "129", 1);  // out of range: ignored
  CHECK(gcp::normalize_points_per_thread(10, sms, 32, 128) == 32);
  setenv("GCP_B200_NORM_PER", "0", 1);
  CHECK(gcp::normalize_points_per_thread(10, sms, 32, 128) == 32);
  unsetenv("GCP_B200_NORM_PER");
  return 0;
}

int main() {
  if (test_normalize_points_per_thread()) return 1;
  if (test_chunk_plan()) return 1;
  if (test_copy_pool()) return 1;
  printf("host logic ok\n");
  return 0;
}
