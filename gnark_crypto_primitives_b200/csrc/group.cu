// Several GPUs of one box behind one handle, for a single-process caller (the Go host of north_star: one cgo
// handle, goroutines on top).  The reference has no multi-device code; the units of its path are
// independent, so a group shards every batch by contiguous index range over its devices, one host thread and one
// context per device, with NO data-path collective.  The one exchange step is the ElGamal tally: every device folds its
// slice to n_fields partial ciphertexts (the ordinary 128-byte wire format), the partials are all-gathered as bytes
// with ncclAllGather over NVLink / NVSwitch (1 KiB per device for 8 fields: latency-bound), and every device folds the
// gathered (devices x n_fields) array.  Edwards addition is exact and associative, so the result does not depend on
// the device count (tests/test_gpu_group.py compares 1 and 2 devices bit for bit).
//
// NCCL is bound at run time (dlopen "libnccl.so.2": inside a torch process that is torch's own copy, elsewhere the
// system one), so the library has no link-time dependency on it; a group of more than one device cannot be created
// without it.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <condition_variable>
#include <cstring>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "internal.h"

namespace {

// the few NCCL entry points used, with their public prototypes (nccl.h); ncclUint8 = 1, ncclSuccess = 0
typedef struct ncclComm* ncclComm_t;
typedef int (*ncclCommInitAll_t)(ncclComm_t*, int, const int*);
typedef int (*ncclCommDestroy_t)(ncclComm_t);
typedef int (*ncclAllGather_t)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t);
typedef const char* (*ncclGetErrorString_t)(int);
constexpr int kNcclUint8 = 1;

struct Nccl {
  void* handle = nullptr;
  ncclCommInitAll_t CommInitAll = nullptr;
  ncclCommDestroy_t CommDestroy = nullptr;
  ncclAllGather_t AllGather = nullptr;
  ncclGetErrorString_t GetErrorString = nullptr;
  bool load(std::string& err) {
    handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!handle) handle = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!handle) {
      err = std::string("NCCL not found: ") + dlerror();
      return false;
    }
    CommInitAll = (ncclCommInitAll_t)dlsym(handle, "ncclCommInitAll");
    CommDestroy = (ncclCommDestroy_t)dlsym(handle, "ncclCommDestroy");
    AllGather = (ncclAllGather_t)dlsym(handle, "ncclAllGather");
    GetErrorString = (ncclGetErrorString_t)dlsym(handle, "ncclGetErrorString");
    if (!CommInitAll || !CommDestroy || !AllGather || !GetErrorString) {
      err = "NCCL symbols missing";
      return false;
    }
    return true;
  }
};

}  // namespace

struct gcp_group {
  std::vector<int> devices;
  std::vector<gcp_ctx*> ctx;
  std::vector<cudaStream_t> stream;       // exchange stream per device
  std::vector<unsigned char*> d_send;     // per device: this device's partial ciphertexts | statuses (kSendStatusOff)
  std::vector<unsigned char*> d_recv;     // per device: gathered partial ciphertexts | gathered statuses
  std::vector<unsigned char*> d_out;      // per device: final tally + status
  std::vector<ncclComm_t> comm;
  Nccl nccl;
  bool have_nccl = false;
  std::mutex mu;  // one group call at a time (the collective must be entered by all devices together)
  std::mutex err_mu;
  std::string err;
  // one persistent host thread per device (round 1 spawned and joined three std::thread fan-outs per call): a call
  // publishes a job, every worker runs it for its device, the caller waits for the last one
  std::vector<std::thread> workers;
  std::mutex wmu;
  std::condition_variable wcv, dcv;
  const std::function<int(int)>* job = nullptr;
  unsigned long long generation = 0;
  int pending = 0;
  std::vector<int> rcs;
  bool stop = false;
};

// message of the last failed gcp_group_create: process-wide, mutex-guarded, read through a per-thread copy (a goroutine may
// run the create call and the gcp_group_last_error(NULL) call on different OS threads)
static std::mutex g_group_create_mu;
static std::string g_group_create_text;
static struct GroupCreateError {
  GroupCreateError& operator=(const std::string& m) {
    std::lock_guard<std::mutex> lk(g_group_create_mu);
    g_group_create_text = m;
    return *this;
  }
  GroupCreateError& operator=(const char* m) { return *this = std::string(m); }
  const char* c_str() const {
    thread_local std::string copy;
    std::lock_guard<std::mutex> lk(g_group_create_mu);
    copy = g_group_create_text;
    return copy.c_str();
  }
} g_group_create_error;

namespace {

struct Shard {
  size_t lo, hi;
};
Shard shard_of(size_t n, int world, int rank) { return {n * (size_t)rank / world, n * (size_t)(rank + 1) / world}; }

constexpr size_t kMaxFields = 64;
constexpr size_t kCtBytes = 128;
constexpr size_t kSendStatusOff = kMaxFields * kCtBytes;  // statuses sit behind the largest ciphertext block

void worker_main(gcp_group* g, int i) {
  cudaSetDevice(g->devices[i]);
  unsigned long long seen = 0;
  for (;;) {
    const std::function<int(int)>* job;
    {
      std::unique_lock<std::mutex> lk(g->wmu);
      g->wcv.wait(lk, [&] { return g->stop || g->generation != seen; });
      if (g->stop) return;
      seen = g->generation;
      job = g->job;
    }
    const int rc = (*job)(i);
    std::lock_guard<std::mutex> lk(g->wmu);
    g->rcs[i] = rc;
    if (--g->pending == 0) g->dcv.notify_all();
  }
}

// run fn(i) for every device on its persistent host thread; returns the first non-zero code and records that device's message
int for_each_device(gcp_group* g, const std::function<int(int)>& fn) {
  const int w = (int)g->ctx.size();
  if (w == 1) {
    g->rcs[0] = fn(0);
  } else {
    std::unique_lock<std::mutex> lk(g->wmu);
    g->job = &fn;
    g->pending = w;
    g->generation++;
    g->wcv.notify_all();
    g->dcv.wait(lk, [&] { return g->pending == 0; });
    g->job = nullptr;
  }
  for (int i = 0; i < w; i++)
    if (g->rcs[i] != GCP_OK) {
      const char* m = gcp_last_error(g->ctx[i]);
      if (g->err.empty()) g->err = "device " + std::to_string(g->devices[i]) + ": " + (m ? m : "error");
      return g->rcs[i];
    }
  return GCP_OK;
}

const char* off(const void* p, size_t bytes) { return p ? (const char*)p + bytes : nullptr; }
char* off(void* p, size_t bytes) { return p ? (char*)p + bytes : nullptr; }
const uint8_t* offb(const uint8_t* p, size_t n) { return p ? p + n : nullptr; }
uint8_t* offb(uint8_t* p, size_t n) { return p ? p + n : nullptr; }

// Exchange of the partial tallies.  Phase 1 (no communication) leaves device i's partial ciphertexts and their statuses in
// d_send[i], in DEVICE memory (round 1 brought them to the host and back).  Phase 2 is entered only when phase 1
// succeeded everywhere, so a device that fails before the collective never leaves the others waiting inside it:
// all-gather of the ciphertext bytes and of the status bytes (ncclAllGather, NVLink), fold of the gathered
// (devices x n_fields) array on every device, statuses merged on the device, result read back from device 0.
// the calling thread's current device is put back on every path (a one-device group runs on the caller's thread)
struct DevGuard {
  int prev = -1;
  DevGuard() {
    if (cudaGetDevice(&prev) != cudaSuccess) {
      cudaGetLastError();
      prev = -1;
    }
  }
  ~DevGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

int gather_and_fold(gcp_group* g, int i, int n_fields, void* out, uint8_t* status, int fmt) {
  const int w = (int)g->ctx.size();
  const size_t pb = (size_t)n_fields * kCtBytes;
  DevGuard dev_guard;
  if (cudaSetDevice(g->devices[i]) != cudaSuccess) return GCP_ERR_CUDA;
  cudaStream_t st = g->stream[i];
  unsigned char* recv_status = g->d_recv[i] + (size_t)w * kSendStatusOff;
  if (w > 1) {
    int nrc = g->nccl.AllGather(g->d_send[i], g->d_recv[i], pb, kNcclUint8, g->comm[i], st);
    if (nrc == 0) nrc = g->nccl.AllGather(g->d_send[i] + kSendStatusOff, recv_status, (size_t)n_fields, kNcclUint8, g->comm[i], st);
    if (nrc != 0) {
      std::lock_guard<std::mutex> lk(g->err_mu);
      g->err = std::string("ncclAllGather: ") + g->nccl.GetErrorString(nrc);
      return GCP_ERR_CUDA;
    }
  } else {
    if (cudaMemcpyAsync(g->d_recv[i], g->d_send[i], pb, cudaMemcpyDeviceToDevice, st) != cudaSuccess ||
        cudaMemcpyAsync(recv_status, g->d_send[i] + kSendStatusOff, (size_t)n_fields, cudaMemcpyDeviceToDevice, st) != cudaSuccess)
      return GCP_ERR_CUDA;
  }
  int rc = gcp_elgamal_tally_dev(g->ctx[i], g->d_recv[i], (size_t)w, n_fields, g->d_out[i], g->d_out[i] + pb, fmt, st);
  if (rc != GCP_OK) return rc;
  rc = gcp_internal_merge_status_dev(g->ctx[i], recv_status, w, n_fields, g->d_out[i], g->d_out[i] + pb, st);
  if (rc != GCP_OK) return rc;
  if (i == 0) {
    if (cudaMemcpyAsync(out, g->d_out[0], pb, cudaMemcpyDeviceToHost, st) != cudaSuccess) return GCP_ERR_CUDA;
    if (cudaMemcpyAsync(status, g->d_out[0] + pb, n_fields, cudaMemcpyDeviceToHost, st) != cudaSuccess) return GCP_ERR_CUDA;
  }
  if (cudaStreamSynchronize(st) != cudaSuccess) return GCP_ERR_CUDA;
  return GCP_OK;
}

}  // namespace

extern "C" {

const char* gcp_group_last_error(const gcp_group* g) { return g ? g->err.c_str() : g_group_create_error.c_str(); }

void gcp_group_destroy(gcp_group* g) {
  if (!g) return;
  DevGuard dev_guard;
  {
    std::lock_guard<std::mutex> lk(g->wmu);
    g->stop = true;
  }
  g->wcv.notify_all();
  for (auto& t : g->workers) t.join();
  for (size_t i = 0; i < g->ctx.size(); i++) {
    if (!g->ctx[i]) continue;  // never created: nothing on that device
    cudaSetDevice(g->devices[i]);
    if (i < g->comm.size() && g->comm[i]) g->nccl.CommDestroy(g->comm[i]);
    if (i < g->d_send.size() && g->d_send[i]) cudaFree(g->d_send[i]);
    if (i < g->d_recv.size() && g->d_recv[i]) cudaFree(g->d_recv[i]);
    if (i < g->d_out.size() && g->d_out[i]) cudaFree(g->d_out[i]);
    if (i < g->stream.size() && g->stream[i]) cudaStreamDestroy(g->stream[i]);
    gcp_ctx_destroy(g->ctx[i]);
  }
  cudaGetLastError();
  delete g;
}

int gcp_group_create(const int* devices, int n_devices, const char* constants_path, gcp_group** out) {
  if (!out) return GCP_ERR_BAD_ARG;
  *out = nullptr;
  if (!devices || n_devices < 1 || n_devices > 64) {
    g_group_create_error = "devices must name 1..64 GPUs";
    return GCP_ERR_BAD_ARG;
  }
  const int visible = gcp_device_count();
  for (int i = 0; i < n_devices; i++)
    if (devices[i] < 0 || devices[i] >= visible) {
      g_group_create_error = visible > 0 ? "device index out of range"
                                         : "no CUDA device visible (this engine has no CPU fallback)";
      return visible > 0 ? GCP_ERR_BAD_ARG : GCP_ERR_NO_DEVICE;
    }
  for (int i = 0; i < n_devices; i++)
    for (int j = 0; j < i; j++)
      if (devices[i] == devices[j]) {
        g_group_create_error = "a device may appear only once in a group";
        return GCP_ERR_BAD_ARG;
      }
  DevGuard dev_guard;
  gcp_group* g = new gcp_group;
  g->devices.assign(devices, devices + n_devices);
  g->ctx.assign(n_devices, nullptr);
  g->stream.assign(n_devices, nullptr);
  g->d_send.assign(n_devices, nullptr);
  g->d_recv.assign(n_devices, nullptr);
  g->d_out.assign(n_devices, nullptr);
  g->comm.assign(n_devices, nullptr);
  g->rcs.assign(n_devices, GCP_OK);
  for (int i = 0; i < n_devices; i++) {
    int rc = gcp_ctx_create(devices[i], constants_path, &g->ctx[i]);
    if (rc != GCP_OK) {
      g_group_create_error = std::string("device ") + std::to_string(devices[i]) + ": " + gcp_last_error(nullptr);
      gcp_group_destroy(g);
      return rc;
    }
    const size_t pb = kMaxFields * kCtBytes;
    if (cudaSetDevice(devices[i]) != cudaSuccess || cudaStreamCreateWithFlags(&g->stream[i], cudaStreamNonBlocking) != cudaSuccess ||
        cudaMalloc(&g->d_send[i], pb + kMaxFields) != cudaSuccess ||
        cudaMalloc(&g->d_recv[i], (pb + kMaxFields) * n_devices) != cudaSuccess ||
        cudaMalloc(&g->d_out[i], pb + kMaxFields) != cudaSuccess) {
      g_group_create_error = std::string("device ") + std::to_string(devices[i]) + ": " + cudaGetErrorString(cudaGetLastError());
      gcp_group_destroy(g);
      return GCP_ERR_CUDA;
    }
  }
  if (n_devices > 1) {
    std::string e;
    if (!g->nccl.load(e)) {
      g_group_create_error = e + " (a group of several GPUs exchanges its partial tallies with ncclAllGather)";
      gcp_group_destroy(g);
      return GCP_ERR_CUDA;
    }
    int nrc = g->nccl.CommInitAll(g->comm.data(), n_devices, g->devices.data());
    if (nrc != 0) {
      g_group_create_error = std::string("ncclCommInitAll: ") + g->nccl.GetErrorString(nrc);
      for (auto& c : g->comm) c = nullptr;
      gcp_group_destroy(g);
      return GCP_ERR_CUDA;
    }
    g->have_nccl = true;
    for (int i = 0; i < n_devices; i++) g->workers.emplace_back(worker_main, g, i);
  }
  *out = g;
  return GCP_OK;
}

int gcp_group_size(const gcp_group* g) { return g ? (int)g->ctx.size() : 0; }
gcp_ctx* gcp_group_ctx(gcp_group* g, int i) { return (g && i >= 0 && i < (int)g->ctx.size()) ? g->ctx[i] : nullptr; }
int gcp_group_uses_nccl(const gcp_group* g) { return (g && g->have_nccl) ? 1 : 0; }

int gcp_group_poseidon_hash(gcp_group* g, const void* in, int arity, size_t n, void* out, uint8_t* status, int fmt) {
  if (!g) return GCP_ERR_BAD_ARG;
  std::lock_guard<std::mutex> lk(g->mu);
  g->err.clear();
  const int w = (int)g->ctx.size();
  const size_t row = (size_t)(arity > 0 ? arity : 0) * 32;
  return for_each_device(g, [&](int i) {
    Shard s = shard_of(n, w, i);
    if (s.hi == s.lo && n != 0) return (int)GCP_OK;
    return gcp_poseidon_hash(g->ctx[i], off(in, s.lo * row), arity, s.hi - s.lo, off(out, s.lo * 32), offb(status, s.lo), fmt);
  });
}

static int group_smt(gcp_group* g, int n_levels, size_t n, const void* roots, int shared_root, const void* siblings,
                     const uint8_t* packed, const uint64_t* offsets, const void* old_keys, const void* old_values,
                     const uint8_t* is_old0, const void* keys, const void* values, const uint8_t* fnc,
                     const uint8_t* enabled, uint8_t* out_flags, uint8_t* out_status, void* out_roots, int fmt) {
  if (!g) return GCP_ERR_BAD_ARG;
  std::lock_guard<std::mutex> lk(g->mu);
  g->err.clear();
  const int w = (int)g->ctx.size();
  const size_t sib_row = (size_t)(n_levels > 0 ? n_levels : 0) * 32;
  return for_each_device(g, [&](int i) {
    Shard s = shard_of(n, w, i);
    const size_t m = s.hi - s.lo;
    if (m == 0 && n != 0) return (int)GCP_OK;
    const void* r = shared_root ? roots : off(roots, s.lo * 32);
    if (siblings)
      return gcp_smt_verify(g->ctx[i], n_levels, m, r, shared_root, off(siblings, s.lo * sib_row), off(old_keys, s.lo * 32),
                            off(old_values, s.lo * 32), offb(is_old0, s.lo), off(keys, s.lo * 32), off(values, s.lo * 32),
                            offb(fnc, s.lo), offb(enabled, s.lo), offb(out_flags, s.lo), offb(out_status, s.lo),
                            off(out_roots, s.lo * 32), fmt);
    // packed offsets are absolute into `packed`: the slice keeps them
    return gcp_smt_verify_packed(g->ctx[i], n_levels, m, r, shared_root, packed, offsets ? offsets + s.lo : nullptr,
                                 off(old_keys, s.lo * 32), off(old_values, s.lo * 32), offb(is_old0, s.lo),
                                 off(keys, s.lo * 32), off(values, s.lo * 32), offb(fnc, s.lo), offb(enabled, s.lo),
                                 offb(out_flags, s.lo), offb(out_status, s.lo), off(out_roots, s.lo * 32), fmt);
  });
}

int gcp_group_smt_verify(gcp_group* g, int n_levels, size_t n, const void* roots, int shared_root, const void* siblings,
                         const void* old_keys, const void* old_values, const uint8_t* is_old0, const void* keys,
                         const void* values, const uint8_t* fnc, const uint8_t* enabled, uint8_t* out_flags,
                         uint8_t* out_status, void* out_roots, int fmt) {
  if (g && n && !siblings) {
    g->err = "null buffer";
    return GCP_ERR_BAD_ARG;
  }
  return group_smt(g, n_levels, n, roots, shared_root, siblings, nullptr, nullptr, old_keys, old_values, is_old0, keys,
                   values, fnc, enabled, out_flags, out_status, out_roots, fmt);
}

int gcp_group_smt_verify_packed(gcp_group* g, int n_levels, size_t n, const void* roots, int shared_root,
                                const uint8_t* packed, const uint64_t* offsets, const void* old_keys,
                                const void* old_values, const uint8_t* is_old0, const void* keys, const void* values,
                                const uint8_t* fnc, const uint8_t* enabled, uint8_t* out_flags, uint8_t* out_status,
                                void* out_roots, int fmt) {
  if (g && n && (!packed || !offsets)) {
    g->err = "null buffer";
    return GCP_ERR_BAD_ARG;
  }
  return group_smt(g, n_levels, n, roots, shared_root, nullptr, packed, offsets, old_keys, old_values, is_old0, keys,
                   values, fnc, enabled, out_flags, out_status, out_roots, fmt);
}

int gcp_group_elgamal_encrypt(gcp_group* g, const void* pub_key, int pk_per_item, const void* k, const void* m, size_t n,
                              void* out_ct, uint8_t* status, int fmt) {
  if (!g) return GCP_ERR_BAD_ARG;
  std::lock_guard<std::mutex> lk(g->mu);
  g->err.clear();
  const int w = (int)g->ctx.size();
  return for_each_device(g, [&](int i) {
    Shard s = shard_of(n, w, i);
    if (s.hi == s.lo && n != 0) return (int)GCP_OK;
    return gcp_elgamal_encrypt(g->ctx[i], pk_per_item ? off(pub_key, s.lo * 64) : pub_key, pk_per_item, off(k, s.lo * 32),
                               off(m, s.lo * 32), s.hi - s.lo, off(out_ct, s.lo * kCtBytes), offb(status, s.lo), fmt);
  });
}

// shared body of the tallies: `partial_fn(i, shard, device partial, device status)` folds device i's slice into d_send[i]
static int group_tally(gcp_group* g, size_t n_ballots, int n_fields, void* out, uint8_t* status, int fmt,
                       const std::function<int(int, Shard, unsigned char*, uint8_t*)>& partial_fn) {
  if (!g) return GCP_ERR_BAD_ARG;
  std::lock_guard<std::mutex> lk(g->mu);
  g->err.clear();
  if (n_fields < 1 || n_fields > (int)kMaxFields) {
    g->err = "n_fields must be in [1, 64]";
    return GCP_ERR_BAD_ARG;
  }
  if (!out || !status) {
    g->err = "null buffer";
    return GCP_ERR_BAD_ARG;
  }
  const int w = (int)g->ctx.size();
  // phase 1: every device folds its slice (no communication); the partial stays on the device
  int rc = for_each_device(g, [&](int i) {
    return partial_fn(i, shard_of(n_ballots, w, i), g->d_send[i], g->d_send[i] + kSendStatusOff);
  });
  if (rc != GCP_OK) return rc;
  // phase 2: all-gather of the partials (bytes) and the final fold on every device
  return for_each_device(g, [&](int i) { return gather_and_fold(g, i, n_fields, out, status, fmt & ~GCP_MSG_U64); });  // partials are ciphertexts
}

int gcp_group_elgamal_tally(gcp_group* g, const void* ct, size_t n_ballots, int n_fields, void* out, uint8_t* status,
                            int fmt) {
  return group_tally(g, n_ballots, n_fields, out, status, fmt, [&](int i, Shard s, unsigned char* p, uint8_t* ps) {
    return gcp_internal_tally_to_dev(g->ctx[i], off(ct, s.lo * (size_t)n_fields * kCtBytes), s.hi - s.lo, n_fields, p, ps, fmt);
  });
}

int gcp_group_elgamal_encrypt_tally(gcp_group* g, const void* pub_key, const void* k, const void* m, size_t n_ballots,
                                    int n_fields, void* out, uint8_t* status, int fmt) {
  return group_tally(g, n_ballots, n_fields, out, status, fmt, [&](int i, Shard s, unsigned char* p, uint8_t* ps) {
    const size_t row = (size_t)n_fields * 32, mrow = (size_t)n_fields * ((fmt & GCP_MSG_U64) ? 8 : 32);
    return gcp_internal_encrypt_tally_to_dev(g->ctx[i], pub_key, off(k, s.lo * row), off(m, s.lo * mrow), s.hi - s.lo, n_fields,
                                             p, ps, fmt);
  });
}

int gcp_group_ballot_batch(gcp_group* g, int n_levels, size_t n_voters, const void* roots, int shared_root,
                           const void* siblings, const uint8_t* packed, const uint64_t* offsets, const void* keys,
                           const void* values, const void* pub_key, const void* k, const void* m, int n_fields,
                           uint8_t* out_flags, uint8_t* out_status, void* out_tally, uint8_t* out_tally_status, int fmt) {
  const size_t sib_row = (size_t)(n_levels > 0 ? n_levels : 0) * 32;
  return group_tally(g, n_voters, n_fields, out_tally, out_tally_status, fmt, [&](int i, Shard s, unsigned char* p, uint8_t* ps) {
    const size_t row = (size_t)n_fields * 32, mrow = (size_t)n_fields * ((fmt & GCP_MSG_U64) ? 8 : 32);
    return gcp_internal_ballot_batch_to_dev(g->ctx[i], n_levels, s.hi - s.lo, shared_root ? roots : off(roots, s.lo * 32),
                                            shared_root, off(siblings, s.lo * sib_row), packed,
                                            offsets ? offsets + s.lo : nullptr, off(keys, s.lo * 32), off(values, s.lo * 32),
                                            pub_key, off(k, s.lo * row), off(m, s.lo * mrow), n_fields, offb(out_flags, s.lo),
                                            offb(out_status, s.lo), p, ps, fmt);
  });
}

}  // extern "C"
