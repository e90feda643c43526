// Process-wide pool of host threads that copy caller memory into page-locked staging buffers.
//
// Why: a cudaMemcpyAsync from PAGEABLE memory (a Go heap slice, a numpy array) is staged by the driver on the calling
// thread at ~5 GB/s; the engine instead copies such sources into a ring of page-locked buffers with several threads and
// sends them from there (h2d_copy in capi.cu).  Round 1 spawned up to 8 std::threads per context per 32 MB slice, so a
// single-process host driving 8 GPUs (gcp_group_*) put 64 short-lived threads on a 32-vCPU machine.  This pool is
// created once per process (first context) and sized to the host: min(16, max(1, hardware threads / 2)) workers in
// total, whatever the number of contexts; a caller splits its slice into pieces, queues them and helps drain the queue
// (its own pieces or another context's) until its slice is done.  GCP_B200_COPY_THREADS overrides the size.
#pragma once
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <mutex>
#include <thread>
#include <vector>

namespace gcp {

class CopyPool {
 public:
  // reference-counted by contexts: the first acquire starts the workers, the last release joins them
  static void acquire() {
    std::lock_guard<std::mutex> lk(life_mu());
    if (refs()++ == 0) instance() = new CopyPool();
  }
  static void release() {
    CopyPool* dead = nullptr;
    {
      std::lock_guard<std::mutex> lk(life_mu());
      if (refs() > 0 && --refs() == 0) {
        dead = instance();
        instance() = nullptr;
      }
    }
    delete dead;
  }
  static int workers() {
    std::lock_guard<std::mutex> lk(life_mu());
    return instance() ? (int)instance()->threads_.size() : 0;
  }

  // memcpy(dst, src, bytes) spread over the pool; returns when every byte is copied.  Safe from any number of threads.
  static void copy(void* dst, const void* src, size_t bytes) {
    CopyPool* p = nullptr;
    {
      std::lock_guard<std::mutex> lk(life_mu());
      p = instance();
    }
    constexpr size_t kPiece = (size_t)2 << 20;
    if (!p || p->threads_.empty() || bytes < 2 * kPiece) {
      memcpy(dst, src, bytes);
      return;
    }
    Job job;
    const size_t n_pieces = (bytes + kPiece - 1) / kPiece;
    job.left.store(n_pieces);
    {
      std::lock_guard<std::mutex> lk(p->mu_);
      for (size_t i = 0; i < n_pieces; i++) {
        const size_t lo = i * kPiece, len = std::min(kPiece, bytes - lo);
        p->queue_.push_back(Piece{(char*)dst + lo, (const char*)src + lo, len, &job});
      }
    }
    p->cv_.notify_all();
    // help: run pieces (of any job) until this job is complete
    while (job.left.load(std::memory_order_acquire) != 0) {
      Piece pc;
      if (p->try_pop(pc)) {
        run(pc);
      } else {
        std::unique_lock<std::mutex> lk(job.mu);
        job.cv.wait_for(lk, std::chrono::microseconds(200), [&] { return job.left.load(std::memory_order_acquire) == 0; });
      }
    }
    // the last worker may still hold job.mu (it decrements and notifies under it): wait for it to let go
    std::lock_guard<std::mutex> lk(job.mu);
  }

 private:
  struct Job {
    std::atomic<size_t> left{0};
    std::mutex mu;
    std::condition_variable cv;
  };
  struct Piece {
    char* dst;
    const char* src;
    size_t len;
    Job* job;
  };

  static std::mutex& life_mu() {
    static std::mutex m;
    return m;
  }
  static int& refs() {
    static int r = 0;
    return r;
  }
  static CopyPool*& instance() {
    static CopyPool* p = nullptr;
    return p;
  }

  CopyPool() {
    unsigned hw = std::thread::hardware_concurrency();
    int n = (int)std::min(16u, std::max(1u, (hw ? hw : 8u) / 2));
    if (const char* env = getenv("GCP_B200_COPY_THREADS")) {
      int v = atoi(env);
      if (v >= 0 && v <= 256) n = v;
    }
    for (int i = 0; i < n; i++) threads_.emplace_back([this] { worker(); });
  }
  ~CopyPool() {
    {
      std::lock_guard<std::mutex> lk(mu_);
      stop_ = true;
    }
    cv_.notify_all();
    for (auto& t : threads_) t.join();
  }

  static void run(const Piece& pc) {
    memcpy(pc.dst, pc.src, pc.len);
    // decrement under the job's mutex: the owner takes that mutex once more before it lets the job go out of scope
    std::lock_guard<std::mutex> lk(pc.job->mu);
    if (pc.job->left.fetch_sub(1, std::memory_order_acq_rel) == 1) pc.job->cv.notify_all();
  }
  bool try_pop(Piece& out) {
    std::lock_guard<std::mutex> lk(mu_);
    if (queue_.empty()) return false;
    out = queue_.front();
    queue_.pop_front();
    return true;
  }
  void worker() {
    for (;;) {
      Piece pc;
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return stop_ || !queue_.empty(); });
        if (queue_.empty()) {
          if (stop_) return;
          continue;
        }
        pc = queue_.front();
        queue_.pop_front();
      }
      run(pc);
    }
  }

  std::mutex mu_;
  std::condition_variable cv_;
  std::deque<Piece> queue_;
  std::vector<std::thread> threads_;
  bool stop_ = false;
};

}  // namespace gcp
