// C ABI of the engine (include/gcp_b200.h): context, constant tables, device memory pools, and the
// host-buffer entry points (chunked, double-buffered H2D -> kernels -> D2H pipelines on two streams).
// There is no CPU compute path here: every value is produced by the kernels in kernels.cu.
#include "../../include/gcp_b200.h"
#include "chunkplan.h"
#include "hostcopy.h"
#include "internal.h"
#include "kernels.h"

#include <dlfcn.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

using namespace gcp;

namespace {

constexpr uint32_t BLOB_MAGIC = 0x32425350u;  // 'PSB2', written by oracle/gen_constants.py
constexpr int N_SLOTS = 128;

// message of the last failed gcp_ctx_create, process-wide and mutex-guarded: a Go caller reads it with a second cgo call
// that may run on another OS thread than the create call did (a goroutine can migrate between the two)
std::mutex g_create_mu;
std::string g_create_error_text;
struct CreateError {
  CreateError& operator=(const std::string& m) {
    std::lock_guard<std::mutex> lk(g_create_mu);
    g_create_error_text = m;
    return *this;
  }
  CreateError& operator=(const char* m) { return *this = std::string(m); }
  const char* c_str() const {  // a per-thread copy, so the pointer stays valid while another thread fails a create
    thread_local std::string copy;
    std::lock_guard<std::mutex> lk(g_create_mu);
    copy = g_create_error_text;
    return copy.c_str();
  }
} g_create_error;

std::string library_dir() {
  Dl_info info;
  if (dladdr((void*)&gcp_ctx_create, &info) && info.dli_fname) {
    std::string p(info.dli_fname);
    size_t k = p.find_last_of('/');
    return k == std::string::npos ? std::string(".") : p.substr(0, k);
  }
  return ".";
}

}  // namespace

struct gcp_ctx {
  int device = 0;
  std::recursive_mutex mu;  // recursive: host entry points hold it for the whole call and call the *_dev entry points
  std::string err;
  cudaStream_t stream[2] = {nullptr, nullptr};
  u32* d_tables = nullptr;
  u32* d_pair_tables = nullptr;  // 18 x 40 elements: PoseidonTable::D
  // small host-buffer hash batches: page-locked memory mapped into the device's address space (inputs | results | status)
  char* small_h = nullptr;
  char* small_d = nullptr;
  PoseidonTable tab[18];  // index by t
  struct Buf {
    void* p = nullptr;
    size_t cap = 0;
  } slot[N_SLOTS];
  uint64_t launches = 0;
  int sm_count = 0;
  // ElGamal: Niels tables of G and of the cached shared public key
  // (d_tabG / d_tabPK point at the ENTRIES of the tables, which is what the kernels take; fb[] owns the allocations)
  u32* d_tabG = nullptr;
  u32* d_tabPK = nullptr;
  struct FbTable {
    u32* alloc = nullptr;     // header + entries
    int cap_bits = 0;         // window width the allocation can hold
    int wbits = 0;            // window width of the table it holds now
    uint64_t uses = 0;        // scalar multiplications served by the base it holds
  } fb[2];                    // 0: G, 1: the cached shared public key
  int fb_forced_bits = 0;     // gcp_ctx_set_fixed_base_window; 0: widen a table once its base has repaid the build
  int fb_wide_bits = 24;      // GCP_B200_FB_WBITS
  uint64_t fb_widen_at = (uint64_t)1 << 27;  // GCP_B200_FB_WIDEN_AT
  bool fb_wide_failed = false;               // the wide allocation did not fit: stay narrow
  u32* d_fb_small = nullptr;  // table-construction scratch (fb_small_scratch_bytes)
  u32* d_base_xy = nullptr;   // 16 words: base point being tabulated
  u32* d_flagG = nullptr;     // 1 word
  u32* d_flagPK = nullptr;    // 1 word: cached key is canonical and on the curve
  unsigned char pk_cached[64];
  int pk_cached_fmt = -1;     // -1: no key cached
  bool have_mimc7 = false;
  u32* d_p2_keys = nullptr;   // 62 Poseidon2 round keys, Montgomery form (poseidon2.cuh)
  bool have_p2_keys = false;
  int smt_hasher = GCP_HASHER_POSEIDON;  // the utils.Hasher plug of the tree/smt gadgets (gcp_ctx_set_smt_hasher)
  // staging ring for large host->device copies from PAGEABLE memory (see h2d_copy)
  static constexpr int STAGE_SLOTS = 4;
  static constexpr size_t STAGE_BYTES = (size_t)32 << 20;
  void* stage_buf[STAGE_SLOTS] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t stage_ev[STAGE_SLOTS] = {nullptr, nullptr, nullptr, nullptr};
  bool stage_used[STAGE_SLOTS] = {false, false, false, false};
  int stage_next = 0;
  bool pool_ref = false;      // this context holds a reference on the process-wide copy pool
  // results on their way to PAGEABLE caller memory (see d2h_copy): small ones park in a page-locked arena, large ones go
  // through a ring of page-locked slices; both are copied out by flush_outputs() once the streams have drained
  static constexpr size_t OUT_ARENA_BYTES = (size_t)64 << 20;
  static constexpr size_t OUT_SMALL_MAX = (size_t)4 << 20;
  static constexpr int OUT_SLOTS = 4;
  static constexpr size_t OUT_SLOT_BYTES = (size_t)32 << 20;
  char* out_arena = nullptr;
  size_t out_arena_used = 0;
  struct ParkedOut {
    void* dst;
    const void* src;
    size_t bytes;
  };
  std::vector<ParkedOut> parked;
  struct OutSlot {
    void* buf = nullptr;
    cudaEvent_t ev = nullptr;
    void* dst = nullptr;  // non-null: holds `bytes` for dst, sent on a stream, event recorded
    size_t bytes = 0;
  } out_slot[OUT_SLOTS];
  int out_next = 0;

  int fail(int code, const std::string& msg) {
    err = msg;
    return code;
  }
  int cuda_fail(cudaError_t e, const char* what) {
    err = std::string(what) + ": " + cudaGetErrorString(e);
    return GCP_ERR_CUDA;
  }
  // grow-only device buffer
  void* buf(int s, size_t bytes) {
    if (bytes == 0) bytes = 16;
    if (slot[s].cap >= bytes) return slot[s].p;
    if (slot[s].p) cudaFree(slot[s].p);
    slot[s].p = nullptr;
    slot[s].cap = 0;
    size_t cap = bytes + bytes / 8;
    if (cudaMalloc(&slot[s].p, cap) != cudaSuccess) {
      cudaGetLastError();
      if (cudaMalloc(&slot[s].p, bytes) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
      }
      cap = bytes;
    }
    slot[s].cap = cap;
    return slot[s].p;
  }
};

// The thread's current CUDA device is the caller's business (torch, another library, a Go thread that also drives other
// GPUs): every entry point selects its context's device for the duration of the call and puts the previous one back.
struct DeviceGuard {
  int prev = -1;
  bool ok = false;
  explicit DeviceGuard(int device) {
    if (cudaGetDevice(&prev) != cudaSuccess) {
      cudaGetLastError();
      prev = -1;
    }
    ok = cudaSetDevice(device) == cudaSuccess;
    if (!ok) cudaGetLastError();
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

#define CU(call, what)                                   \
  do {                                                   \
    cudaError_t e_ = (call);                             \
    if (e_ != cudaSuccess) return ctx->cuda_fail(e_, what); \
  } while (0)

// Host -> device copy of a caller buffer on `st`.  Page-locked sources (gcp_host_alloc, cudaHostRegister, torch pinned
// memory) go straight to cudaMemcpyAsync.  PAGEABLE sources - a Go heap slice, a numpy array - would be staged by the
// driver on the calling thread at ~5 GB/s (measured: 583 k instead of 738 k proofs/s end to end at 2^19 dense proofs);
// from 1 MB up they are copied into a ring of four 32 MB page-locked buffers by the process-wide copy pool
// (hostcopy.h) at memory bandwidth and sent from there, so the copy of chunk k+1 keeps pace with the kernels of chunk k
// and the calling thread is back at its launch loop as soon as the last slice is queued.
static bool staging_disabled() {
  static const bool off = getenv("GCP_B200_NO_STAGING") != nullptr;
  return off;
}

static int h2d_copy(gcp_ctx* ctx, void* dst, const void* src, size_t bytes, cudaStream_t st) {
  if (bytes == 0) return GCP_OK;
  bool pageable = false;
  if (bytes >= ((size_t)1 << 20) && !staging_disabled()) {
    cudaPointerAttributes attr;
    cudaError_t e = cudaPointerGetAttributes(&attr, src);
    if (e != cudaSuccess) cudaGetLastError();
    pageable = (e != cudaSuccess) || attr.type == cudaMemoryTypeUnregistered;
  }
  if (!pageable) {
    CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st), "H2D");
    return GCP_OK;
  }
  // the whole ring is allocated at its first use (page-locking 32 MB takes milliseconds: not inside a later call's pipeline)
  for (int slot = 0; slot < gcp_ctx::STAGE_SLOTS; slot++) {
    if (ctx->stage_buf[slot]) continue;
    if (cudaHostAlloc(&ctx->stage_buf[slot], gcp_ctx::STAGE_BYTES, cudaHostAllocDefault) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->stage_ev[slot], cudaEventDisableTiming) != cudaSuccess) {
      cudaGetLastError();
      if (ctx->stage_buf[slot]) cudaFreeHost(ctx->stage_buf[slot]);
      ctx->stage_buf[slot] = nullptr;
      // no page-locked memory to be had: let the driver stage the copy
      CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st), "H2D");
      return GCP_OK;
    }
  }
  for (size_t off = 0; off < bytes; off += gcp_ctx::STAGE_BYTES) {
    const size_t m = std::min(gcp_ctx::STAGE_BYTES, bytes - off);
    const int slot = ctx->stage_next;
    ctx->stage_next = (slot + 1) % gcp_ctx::STAGE_SLOTS;
    if (ctx->stage_used[slot]) CU(cudaEventSynchronize(ctx->stage_ev[slot]), "staging event");  // its last send is done
    CopyPool::copy(ctx->stage_buf[slot], (const char*)src + off, m);
    CU(cudaMemcpyAsync((char*)dst + off, ctx->stage_buf[slot], m, cudaMemcpyHostToDevice, st), "H2D");
    CU(cudaEventRecord(ctx->stage_ev[slot], st), "staging event");
    ctx->stage_used[slot] = true;
  }
  return GCP_OK;
}

// Device -> host copy of results into a caller buffer on `st`.  A cudaMemcpyAsync into PAGEABLE memory does not return
// before the copy has run, i.e. before every kernel queued ahead of it on that stream has finished: the launch loop of a
// chunked pipeline then stalls after each chunk and the next chunk's upload is no longer hidden (measured: census-like
// proofs from host rows 3.3 M/s with numpy outputs against 5.0 M/s resident).  Such results are therefore sent to
// page-locked memory first and copied to the caller by flush_outputs(), which every host entry point runs (through its
// StreamGuard) after its streams have drained: results up to 4 MB park in a 64 MB arena, larger ones stream through a ring
// of four 32 MB slices that the copy pool empties as they come back.
static void release_out_slot(gcp_ctx* ctx, gcp_ctx::OutSlot& o) {
  if (!o.dst) return;
  cudaEventSynchronize(o.ev);
  CopyPool::copy(o.dst, o.buf, o.bytes);
  o.dst = nullptr;
}

static void flush_outputs(gcp_ctx* ctx) {  // call with both pipeline streams drained
  for (auto& o : ctx->out_slot) release_out_slot(ctx, o);
  for (const auto& p : ctx->parked) CopyPool::copy(p.dst, p.src, p.bytes);
  ctx->parked.clear();
  ctx->out_arena_used = 0;
}

static int d2h_copy(gcp_ctx* ctx, void* dst, const void* d_src, size_t bytes, cudaStream_t st) {
  if (bytes == 0) return GCP_OK;
  bool pageable = false;
  if (!staging_disabled()) {
    cudaPointerAttributes attr;
    cudaError_t e = cudaPointerGetAttributes(&attr, dst);
    if (e != cudaSuccess) cudaGetLastError();
    pageable = (e != cudaSuccess) || attr.type == cudaMemoryTypeUnregistered;
  }
  if (pageable && bytes <= gcp_ctx::OUT_SMALL_MAX) {
    if (!ctx->out_arena && cudaHostAlloc((void**)&ctx->out_arena, gcp_ctx::OUT_ARENA_BYTES, cudaHostAllocDefault) != cudaSuccess) {
      cudaGetLastError();
      ctx->out_arena = nullptr;
    }
    const size_t need = (bytes + 255) & ~(size_t)255;
    if (ctx->out_arena && ctx->out_arena_used + need <= gcp_ctx::OUT_ARENA_BYTES) {
      char* p = ctx->out_arena + ctx->out_arena_used;
      ctx->out_arena_used += need;
      CU(cudaMemcpyAsync(p, d_src, bytes, cudaMemcpyDeviceToHost, st), "D2H");
      ctx->parked.push_back({dst, p, bytes});
      return GCP_OK;
    }
    pageable = false;  // arena full or unavailable: plain (blocking) copy
  }
  if (!pageable) {
    CU(cudaMemcpyAsync(dst, d_src, bytes, cudaMemcpyDeviceToHost, st), "D2H");
    return GCP_OK;
  }
  // the whole ring is allocated at its first use (page-locking 32 MB takes milliseconds: not inside a later call's pipeline)
  for (auto& o : ctx->out_slot) {
    if (o.buf) continue;
    if (cudaHostAlloc(&o.buf, gcp_ctx::OUT_SLOT_BYTES, cudaHostAllocDefault) != cudaSuccess ||
        cudaEventCreateWithFlags(&o.ev, cudaEventDisableTiming) != cudaSuccess) {
      cudaGetLastError();
      if (o.buf) cudaFreeHost(o.buf);
      o.buf = nullptr;
      CU(cudaMemcpyAsync(dst, d_src, bytes, cudaMemcpyDeviceToHost, st), "D2H");  // no page-locked memory: plain copy
      return GCP_OK;
    }
  }
  for (size_t off = 0; off < bytes; off += gcp_ctx::OUT_SLOT_BYTES) {
    const size_t m = std::min(gcp_ctx::OUT_SLOT_BYTES, bytes - off);
    gcp_ctx::OutSlot& o = ctx->out_slot[ctx->out_next];
    ctx->out_next = (ctx->out_next + 1) % gcp_ctx::OUT_SLOTS;
    release_out_slot(ctx, o);
    CU(cudaMemcpyAsync(o.buf, (const char*)d_src + off, m, cudaMemcpyDeviceToHost, st), "D2H");
    CU(cudaEventRecord(o.ev, st), "staging event");
    o.dst = (char*)dst + off;
    o.bytes = m;
  }
  return GCP_OK;
}

// Every host-buffer entry point holds one of these: whatever path it returns by (an allocation failure in the middle of a
// chunk loop included), both pipeline streams have drained, so no copy into or out of the caller's buffers is still in
// flight (the header's promise that no host pointer is used after the call returns: the cgo pointer rules).
struct StreamGuard {
  gcp_ctx* c;
  ~StreamGuard();
};

StreamGuard::~StreamGuard() {
  cudaStreamSynchronize(c->stream[0]);
  cudaStreamSynchronize(c->stream[1]);
  flush_outputs(c);  // results parked in page-locked memory reach the caller's buffers before the call returns
}

#define GCP_TRY(call)            \
  do {                           \
    int rc_ = (call);            \
    if (rc_ != GCP_OK) return rc_; \
  } while (0)



extern "C" {

int gcp_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

const char* gcp_last_error(const gcp_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }
int gcp_ctx_device(const gcp_ctx* ctx) { return ctx ? ctx->device : -1; }
uint64_t gcp_ctx_launch_count(const gcp_ctx* ctx) { return ctx ? ctx->launches : 0; }

int gcp_ctx_smt_hasher(const gcp_ctx* ctx) { return ctx ? ctx->smt_hasher : -1; }

int gcp_ctx_set_smt_hasher(gcp_ctx* ctx, int hasher) {
  if (!ctx) return GCP_ERR_BAD_ARG;
  std::lock_guard<std::recursive_mutex> lk(ctx->mu);
  if (hasher != GCP_HASHER_POSEIDON && hasher != GCP_HASHER_POSEIDON2)
    return ctx->fail(GCP_ERR_BAD_ARG, "unknown hasher (GCP_HASHER_POSEIDON or GCP_HASHER_POSEIDON2)");
  if (hasher == GCP_HASHER_POSEIDON2 && !ctx->have_p2_keys)
    return ctx->fail(GCP_ERR_CONSTANTS,
                     "Poseidon2 round keys missing (data/poseidon2_bn254_t2.bin not found and gcp_poseidon2_set_round_keys not called)");
  // queued work of the two pipeline streams was launched with the previous plug: nothing to wait for, the plug is a launch argument
  ctx->smt_hasher = hasher;
  return GCP_OK;
}

void gcp_ctx_destroy(gcp_ctx* ctx) {
  if (!ctx) return;
  DeviceGuard device_guard(ctx->device);
  for (auto& s : ctx->stream)
    if (s) {
      cudaStreamSynchronize(s);
      cudaStreamDestroy(s);
    }
  for (auto& b : ctx->slot)
    if (b.p) cudaFree(b.p);
  for (int i = 0; i < gcp_ctx::STAGE_SLOTS; i++) {
    if (ctx->stage_ev[i]) cudaEventDestroy(ctx->stage_ev[i]);
    if (ctx->stage_buf[i]) cudaFreeHost(ctx->stage_buf[i]);
  }
  if (ctx->d_tables) cudaFree(ctx->d_tables);
  if (ctx->d_pair_tables) cudaFree(ctx->d_pair_tables);
  if (ctx->small_h) cudaFreeHost(ctx->small_h);
  for (u32* p : {ctx->fb[0].alloc, ctx->fb[1].alloc, ctx->d_fb_small, ctx->d_base_xy, ctx->d_flagG, ctx->d_flagPK, ctx->d_p2_keys})
    if (p) cudaFree(p);
  if (ctx->out_arena) cudaFreeHost(ctx->out_arena);
  for (auto& o : ctx->out_slot) {
    if (o.ev) cudaEventDestroy(o.ev);
    if (o.buf) cudaFreeHost(o.buf);
  }
  if (ctx->pool_ref) CopyPool::release();
  delete ctx;
}

int gcp_copy_threads(void) { return CopyPool::workers(); }

// Bandwidth of the staging path on this host: pageable memory -> page-locked memory through the copy pool, GB/s (what
// bounds the host-buffer calls when their inputs live in ordinary memory and the kernels outrun it).
int gcp_copy_probe(size_t bytes, double* gb_per_s) {
  if (!gb_per_s || bytes < ((size_t)1 << 20)) return GCP_ERR_BAD_ARG;
  *gb_per_s = 0;
  if (gcp_device_count() <= 0) return GCP_ERR_NO_DEVICE;
  void* dst = nullptr;
  if (cudaHostAlloc(&dst, bytes, cudaHostAllocDefault) != cudaSuccess) {
    cudaGetLastError();
    return GCP_ERR_ALLOC;
  }
  char* src = (char*)malloc(bytes);
  if (!src) {
    cudaFreeHost(dst);
    return GCP_ERR_ALLOC;
  }
  memset(src, 1, bytes);
  CopyPool::acquire();
  CopyPool::copy(dst, src, bytes);  // warm: faults the destination in
  double best = 0;
  for (int it = 0; it < 3; it++) {
    auto t0 = std::chrono::steady_clock::now();
    CopyPool::copy(dst, src, bytes);
    double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (dt > 0) best = std::max(best, (double)bytes / dt / 1e9);
  }
  CopyPool::release();
  free(src);
  cudaFreeHost(dst);
  *gb_per_s = best;
  return GCP_OK;
}

int gcp_host_alloc(size_t bytes, void** out) {
  if (!out) return GCP_ERR_BAD_ARG;
  *out = nullptr;
  if (gcp_device_count() <= 0) return GCP_ERR_NO_DEVICE;
  if (cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocPortable) != cudaSuccess) {
    cudaGetLastError();
    *out = nullptr;
    return GCP_ERR_ALLOC;
  }
  return GCP_OK;
}

void gcp_host_free(void* p) {
  if (p) cudaFreeHost(p);
}

// ---- fixed-base tables (elgamal.cuh) ---------------------------------------------------------------------------------
// Every table starts FB_NARROW_BITS wide (654 MB, ~2 ms to build).  A base that has served fb_widen_at scalar
// multiplications (default 2^27: the 8.9 GB, ~45 ms build of the 24-bit table is repaid by its 12-15 % shorter
// multiplications at about that point) gets the wide table; gcp_ctx_set_fixed_base_window forces one width.
static constexpr int FB_NARROW_BITS = 20;

// Builds the table of the point in ctx->d_base_xy (which = 1) or of the generator (which = 0) at `wbits`; the streams must
// be idle.  A wider table than the allocation holds is built beside the old one and swapped in.
static int fb_build_table(gcp_ctx* ctx, int which, int wbits, int base_mont, int te, cudaStream_t st) {
  gcp_ctx::FbTable& t = ctx->fb[which];
  cudaError_t e;
  u32* fresh = nullptr;
  if (wbits > t.cap_bits) {
    if ((e = cudaMalloc(&fresh, fb_table_bytes(wbits))) != cudaSuccess) {
      cudaGetLastError();
      if (!t.alloc) return ctx->cuda_fail(e, "cudaMalloc fixed-base table");
      ctx->fb_wide_failed = true;  // no room for the wide table: keep multiplying out of the one we have
      if (which == 0) return GCP_OK;
      wbits = t.cap_bits;
    }
  }
  u32* dst = fresh ? fresh : t.alloc;
  if (which == 0 && (e = upload_generator(ctx->d_base_xy, st)) != cudaSuccess) return ctx->cuda_fail(e, "upload generator");
  if ((e = launch_fb_table_build(ctx->d_base_xy, base_mont, ctx->d_fb_small, dst, which ? ctx->d_flagPK : ctx->d_flagG, st, te,
                                 wbits)) != cudaSuccess ||
      (e = cudaStreamSynchronize(st)) != cudaSuccess) {
    if (fresh) cudaFree(fresh);
    return ctx->cuda_fail(e, "fixed-base table build");
  }
  ctx->launches += 4;
  if (fresh) {
    if (t.alloc) cudaFree(t.alloc);
    t.alloc = fresh;
    t.cap_bits = wbits;
  }
  t.wbits = wbits;
  (which ? ctx->d_tabPK : ctx->d_tabG) = fb_table_entries(t.alloc);
  return GCP_OK;
}

static int fb_wanted_bits(const gcp_ctx* ctx, uint64_t uses) {
  if (ctx->fb_forced_bits) return ctx->fb_forced_bits;
  return (uses >= ctx->fb_widen_at && !ctx->fb_wide_failed) ? ctx->fb_wide_bits : FB_NARROW_BITS;
}

// `n` more multiplications by G are about to be queued on `st`: widen G's table if it has earned it
static int fb_note_g_use(gcp_ctx* ctx, uint64_t n, cudaStream_t st) {
  gcp_ctx::FbTable& t = ctx->fb[0];
  t.uses += n;
  const int want = fb_wanted_bits(ctx, t.uses);
  if (want == t.wbits || (want < t.wbits && !ctx->fb_forced_bits)) return GCP_OK;
  cudaError_t e;  // work queued earlier may still read the table
  if ((e = cudaStreamSynchronize(st)) != cudaSuccess || (e = cudaStreamSynchronize(ctx->stream[0])) != cudaSuccess ||
      (e = cudaStreamSynchronize(ctx->stream[1])) != cudaSuccess)
    return ctx->cuda_fail(e, "stream sync");
  return fb_build_table(ctx, 0, want, 0, 0, st);
}

int gcp_ctx_create(int device, const char* constants_path, gcp_ctx** out) {
  if (!out) {
    g_create_error = "out is NULL";
    return GCP_ERR_BAD_ARG;
  }
  *out = nullptr;
  int ndev = gcp_device_count();
  if (ndev <= 0) {
    g_create_error = "no CUDA device visible (this engine has no CPU fallback)";
    return GCP_ERR_NO_DEVICE;
  }
  if (device < 0 || device >= ndev) {
    g_create_error = "device index out of range";
    return GCP_ERR_BAD_ARG;
  }
  // ---- read the constant blob ----------------------------------------------------------------
  std::string path;
  if (constants_path && *constants_path)
    path = constants_path;
  else if (const char* env = getenv("GCP_B200_CONSTANTS"))
    path = env;
  else
    path = library_dir() + "/data/poseidon_bn254.bin";
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) {
    g_create_error = "cannot open Poseidon constant blob: " + path;
    return GCP_ERR_CONSTANTS;
  }
  std::vector<unsigned char> blob;
  {
    fseek(f, 0, SEEK_END);
    long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    blob.resize(sz > 0 ? (size_t)sz : 0);
    size_t got = blob.empty() ? 0 : fread(blob.data(), 1, blob.size(), f);
    fclose(f);
    if (got != blob.size() || blob.size() < 16 + 16 * 40) {
      g_create_error = "short read on constant blob: " + path;
      return GCP_ERR_CONSTANTS;
    }
  }
  uint32_t hdr[4];
  memcpy(hdr, blob.data(), 16);
  if (hdr[0] != BLOB_MAGIC || hdr[1] != 1 || hdr[2] != 16) {
    g_create_error = "bad constant blob header: " + path;
    return GCP_ERR_CONSTANTS;
  }
  const size_t base = 16 + 16 * 40;
  const size_t n_elems = (blob.size() - base) / 32;

  gcp_ctx* ctx = new gcp_ctx();
  ctx->device = device;
  CopyPool::acquire();
  ctx->pool_ref = true;
  auto bail = [&](int code) {
    g_create_error = ctx->err;
    gcp_ctx_destroy(ctx);
    return code;
  };
  cudaError_t e;
  DeviceGuard device_guard(device);
  if (!device_guard.ok) return bail(ctx->fail(GCP_ERR_CUDA, "cudaSetDevice failed"));
  for (auto& s : ctx->stream)
    if ((e = cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking)) != cudaSuccess)
      return bail(ctx->cuda_fail(e, "cudaStreamCreate"));
  if ((e = cudaMalloc(&ctx->d_tables, n_elems * 32)) != cudaSuccess) return bail(ctx->cuda_fail(e, "cudaMalloc tables"));
  if ((e = cudaMemcpyAsync(ctx->d_tables, blob.data() + base, n_elems * 32, cudaMemcpyHostToDevice, ctx->stream[0])) !=
      cudaSuccess)
    return bail(ctx->cuda_fail(e, "upload tables"));
  if ((e = launch_to_mont(ctx->d_tables, n_elems, ctx->stream[0])) != cudaSuccess)
    return bail(ctx->cuda_fail(e, "to_mont kernel (is this an sm_100a device?)"));
  ctx->launches++;
  for (int i = 0; i < 16; i++) {
    uint32_t d[10];
    memcpy(d, blob.data() + 16 + 40 * i, 40);
    int t = (int)d[0];
    if (t != i + 2 || (size_t)d[8] + d[9] > n_elems) return bail(ctx->fail(GCP_ERR_CONSTANTS, "bad constant directory"));
    PoseidonTable& pt = ctx->tab[t];
    pt.t = t;
    pt.RP = (int)d[1];
    pt.C = ctx->d_tables + (size_t)d[2] * 8;
    pt.S = ctx->d_tables + (size_t)d[4] * 8;
    pt.M = ctx->d_tables + (size_t)d[6] * 8;
    pt.P = ctx->d_tables + (size_t)d[8] * 8;
    // the constant-memory kernels assume C | S | M | P are contiguous in that order
    if (d[4] != d[2] + d[3] || d[6] != d[4] + d[5] || d[8] != d[6] + d[7])
      return bail(ctx->fail(GCP_ERR_CONSTANTS, "constant tables not contiguous"));
  }
  if ((e = upload_const_tables(ctx->tab[3].C, ctx->tab[4].C, ctx->stream[0])) != cudaSuccess)
    return bail(ctx->cuda_fail(e, "constant-memory upload"));
  // derived constants of the partial-round pairs, one table of RP / 2 elements per t (poseidon.cuh)
  if ((e = cudaMalloc(&ctx->d_pair_tables, (size_t)18 * 40 * 32)) != cudaSuccess)
    return bail(ctx->cuda_fail(e, "cudaMalloc pair constants"));
  for (int t = 2; t <= 17; t++) {
    PoseidonTable& pt = ctx->tab[t];
    if (pt.RP / 2 > 40) return bail(ctx->fail(GCP_ERR_CONSTANTS, "more partial rounds than the pair table holds"));
    u32* d = ctx->d_pair_tables + (size_t)t * 40 * 8;
    if ((e = launch_poseidon_pair_constants(pt, d, ctx->stream[0])) != cudaSuccess)
      return bail(ctx->cuda_fail(e, "pair-constant kernel"));
    pt.D = d;
    ctx->launches++;
  }
  // fixed-base table of the generator G (elgamal/mul.go:26-72 restated with wider windows), built on the device
  if ((e = cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device)) != cudaSuccess)
    return bail(ctx->cuda_fail(e, "device attribute"));
  if (const char* v = getenv("GCP_B200_FB_WBITS")) {
    const int b = atoi(v);
    if (b >= FB_NARROW_BITS && b <= 26) ctx->fb_wide_bits = b;
  }
  if (const char* v = getenv("GCP_B200_FB_WIDEN_AT")) ctx->fb_widen_at = strtoull(v, nullptr, 10);
  if ((e = cudaMalloc(&ctx->d_fb_small, fb_small_scratch_bytes())) != cudaSuccess ||
      (e = cudaMalloc(&ctx->d_base_xy, 64)) != cudaSuccess || (e = cudaMalloc(&ctx->d_flagG, 4)) != cudaSuccess ||
      (e = cudaMalloc(&ctx->d_flagPK, 4)) != cudaSuccess)
    return bail(ctx->cuda_fail(e, "cudaMalloc fixed-base scratch"));
  {
    int rc = fb_build_table(ctx, 0, FB_NARROW_BITS, 0, 0, ctx->stream[0]);
    if (rc == GCP_OK) rc = fb_build_table(ctx, 1, FB_NARROW_BITS, 0, 0, ctx->stream[0]);  // a valid table (of G) until a key arrives
    if (rc != GCP_OK) return bail(rc);
  }
  {
    u32 flag = 0;
    if ((e = cudaMemcpy(&flag, ctx->d_flagG, 4, cudaMemcpyDeviceToHost)) != cudaSuccess)
      return bail(ctx->cuda_fail(e, "read generator flag"));
    if (!flag) return bail(ctx->fail(GCP_ERR_CONSTANTS, "generator self-check failed (not on curve)"));
  }
  // MiMC7 round constants (data/mimc7_bn254.bin next to the Poseidon blob); optional: without the file MiMC7 is disabled
  {
    std::string mpath = path;
    size_t k = mpath.find_last_of('/');
    mpath = (k == std::string::npos ? std::string("") : mpath.substr(0, k + 1)) + "mimc7_bn254.bin";
    FILE* mf = fopen(mpath.c_str(), "rb");
    if (mf) {
      unsigned char mb[16 + 91 * 32];
      size_t got = fread(mb, 1, sizeof(mb), mf);
      fclose(mf);
      uint32_t mh[4];
      memcpy(mh, mb, 16);
      if (got == sizeof(mb) && mh[0] == 0x374D494Du && mh[1] == 1 && mh[2] == 91) {
        u32* d_m = (u32*)ctx->buf(47, 91 * 32);
        if (!d_m) return bail(ctx->fail(GCP_ERR_ALLOC, "device allocation failed"));
        if ((e = cudaMemcpyAsync(d_m, mb + 16, 91 * 32, cudaMemcpyHostToDevice, ctx->stream[0])) != cudaSuccess ||
            (e = launch_to_mont(d_m, 91, ctx->stream[0])) != cudaSuccess ||
            (e = upload_mimc7_constants(d_m, ctx->stream[0])) != cudaSuccess ||
            (e = cudaStreamSynchronize(ctx->stream[0])) != cudaSuccess)
          return bail(ctx->cuda_fail(e, "MiMC7 constants"));
        ctx->launches++;
        ctx->have_mimc7 = true;
      }
    }
  }
  // Poseidon2 (t = 2) round keys (data/poseidon2_bn254_t2.bin); optional, and replaceable with
  // gcp_poseidon2_set_round_keys: without either the Poseidon2 entry points fail with GCP_ERR_CONSTANTS
  {
    if ((e = cudaMalloc(&ctx->d_p2_keys, 62 * 32)) != cudaSuccess) return bail(ctx->cuda_fail(e, "cudaMalloc Poseidon2 keys"));
    std::string ppath = path;
    size_t k = ppath.find_last_of('/');
    ppath = (k == std::string::npos ? std::string("") : ppath.substr(0, k + 1)) + "poseidon2_bn254_t2.bin";
    FILE* pf = fopen(ppath.c_str(), "rb");
    if (pf) {
      unsigned char pb[16 + 62 * 32];
      size_t got = fread(pb, 1, sizeof(pb), pf);
      fclose(pf);
      uint32_t ph[4];
      memcpy(ph, pb, 16);
      if (got == sizeof(pb) && ph[0] == 0x32534F50u && ph[1] == 1 && ph[2] == 62) {
        if ((e = cudaMemcpyAsync(ctx->d_p2_keys, pb + 16, 62 * 32, cudaMemcpyHostToDevice, ctx->stream[0])) != cudaSuccess ||
            (e = launch_to_mont(ctx->d_p2_keys, 62, ctx->stream[0])) != cudaSuccess ||
            (e = cudaStreamSynchronize(ctx->stream[0])) != cudaSuccess)
          return bail(ctx->cuda_fail(e, "Poseidon2 round keys"));
        ctx->launches++;
        ctx->have_p2_keys = true;
      }
    }
  }
  *out = ctx;
  return GCP_OK;
}

int gcp_probe_imad_wide(gcp_ctx* ctx, double* wide_mul_per_s) {
  if (!ctx || !wide_mul_per_s) return GCP_ERR_BAD_ARG;
  std::lock_guard<std::recursive_mutex> lk(ctx->mu);
  DeviceGuard device_guard(ctx->device);
  if (!device_guard.ok) return ctx->fail(GCP_ERR_CUDA, "cudaSetDevice failed");
  int sms = 0;
  CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device), "device attribute");
  const int blocks = sms * 8;
  u32* d_out = (u32*)ctx->buf(42, (size_t)blocks * 256 * 4);
  if (!d_out) return ctx->fail(GCP_ERR_ALLOC, "device allocation failed");
  cudaStream_t st = ctx->stream[0];
  cudaEvent_t e0, e1;
  CU(cudaEventCreate(&e0), "event");
  CU(cudaEventCreate(&e1), "event");
  double best = 0;
  for (int it = 0; it < 10; it++) {  // ~9 ms per launch; the first launches warm the clocks up, best of the rest
    CU(cudaEventRecord(e0, st), "event record");
    double ops = launch_imad_probe(d_out, blocks, 17u + it, st);
    ctx->launches++;
    CU(cudaEventRecord(e1, st), "event record");
    CU(cudaEventSynchronize(e1), "imad probe kernel");
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, e0, e1), "event elapsed");
    if (it >= 2 && ms > 0) best = std::max(best, ops / (ms * 1e-3));
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  *wide_mul_per_s = best;
  return GCP_OK;
}

// ---------------------------------------------------------------------------------------------------
// Poseidon
// ---------------------------------------------------------------------------------------------------
static int poseidon_hash_dev_locked(gcp_ctx* ctx, const void* d_in, int arity, size_t n, void* d_out,
                                    uint8_t* d_status, int fmt, cudaStream_t st) {
  if (arity < 1 || arity > 16) return ctx->fail(GCP_ERR_BAD_ARG, "bad inputs provided");  // poseidon.go:41-43
  if (fmt != GCP_FMT_CANONICAL && fmt != GCP_FMT_MONTGOMERY) return ctx->fail(GCP_ERR_BAD_ARG, "bad element format");
  if (n == 0) return GCP_OK;
  if (!d_in || !d_out) return ctx->fail(GCP_ERR_BAD_ARG, "null buffer");
  if (d_status) CU(cudaMemsetAsync(d_status, 0, n, st), "memset status");
  CU(launch_poseidon(ctx->tab[arity + 1], (const u32*)d_in, (u32*)d_out, d_status, n, 1, (size_t)arity, 0, 1, fmt, fmt,
                     1, st),
     "poseidon kernel");
  ctx->launches++;
  return GCP_OK;
}

// MultiHash (poseidon.go:54-91): 16-wide chunks, then the hash of the chunk hashes, recursively.
// scratch slots 0/1 hold the intermediate levels (Montgomery form).
static int poseidon_multihash_dev_locked(gcp_ctx* ctx, const void* d_in, int len, size_t n, void* d_out,
                                         uint8_t* d_status, int fmt, cudaStream_t st, int slot_base) {
  if (len < 1) return ctx->fail(GCP_ERR_BAD_ARG, "bad inputs provided");
  if (len > 4096) return ctx->fail(GCP_ERR_BAD_ARG, "the maximum number of inputs supported is 4096");
  if (len <= 16) return poseidon_hash_dev_locked(ctx, d_in, len, n, d_out, d_status, fmt, st);
  if (fmt != GCP_FMT_CANONICAL && fmt != GCP_FMT_MONTGOMERY) return ctx->fail(GCP_ERR_BAD_ARG, "bad element format");
  if (n == 0) return GCP_OK;
  if (!d_in || !d_out) return ctx->fail(GCP_ERR_BAD_ARG, "null buffer");
  if (d_status) CU(cudaMemsetAsync(d_status, 0, n, st), "memset status");
  const u32* cur = (const u32*)d_in;
  int cur_len = len;
  int cur_fmt = fmt;
  int level = 0;
  while (cur_len > 16) {
    int full = cur_len / 16, rem = cur_len % 16;
    int nchunks = full + (rem ? 1 : 0);
    u32* nxt = (u32*)ctx->buf(slot_base + (level & 1), n * (size_t)nchunks * 32);
    if (!nxt) return ctx->fail(GCP_ERR_ALLOC, "device allocation failed (multihash scratch)");
    CU(launch_poseidon(ctx->tab[17], cur, nxt, d_status, n, full, (size_t)cur_len, 16, (size_t)nchunks, cur_fmt,
                       GCP_FMT_MONTGOMERY, 0, st),
       "poseidon kernel (chunks)");
    ctx->launches++;
    if (rem) {
      CU(launch_poseidon(ctx->tab[rem + 1], cur + (size_t)full * 16 * 8, nxt + (size_t)full * 8, d_status, n, 1,
                         (size_t)cur_len, 0, (size_t)nchunks, cur_fmt, GCP_FMT_MONTGOMERY, 0, st),
         "poseidon kernel (tail chunk)");
      ctx->launches++;
    }
    cur = nxt;
    cur_len = nchunks;
    cur_fmt = GCP_FMT_MONTGOMERY;
    level++;
  }
  CU(launch_poseidon(ctx->tab[cur_len + 1], cur, (u32*)d_out, d_status, n, 1, (size_t)cur_len, 0, 1, cur_fmt, fmt, 1, st),
     "poseidon kernel (final)");
  ctx->launches++;
  return GCP_OK;
}

int gcp_poseidon_hash_dev(gcp_ctx* ctx, const void* d_in, int arity, size_t n, void* d_out, uint8_t* d_status, int fmt,
                          void* stream) {
  if (!ctx) return GCP_ERR_BAD_ARG;
  std::lock_guard<std::recursive_mutex> lk(ctx->mu);
  DeviceGuard device_guard(ctx->device);
  if (!device_guard.ok) return ctx->fail(GCP_ERR_CUDA, "cudaSetDevice failed");
  return poseidon_hash_dev_locked(ctx, d_in, arity, n, d_out, d_status, fmt, (cudaStream_t)stream);
}

int gcp_poseidon_multihash_dev(gcp_ctx* ctx, const void* d_in, int len, size_t n, void* d_out, uint8_t* d_status,
                               int fmt, void* stream) {
  if (!ctx) return GCP_ERR_BAD_ARG;
  std::lock_guard<std::recursive_mutex> lk(ctx->mu);
  DeviceGuard device_guard(ctx->device);
  if (!device_guard.ok) return ctx->fail(GCP_ERR_CUDA, "cudaSetDevice failed");
  return poseidon_multihash_dev_locked(ctx, d_in, len, n, d_out, d_status, fmt, (cudaStream_t)stream, 0);
}

// Host-buffer form: chunks of items, two streams alternate so that the copy of chunk k+1 overlaps the kernels of chunk k.
constexpr size_t SMALL_HASH_IN_BYTES = (size_t)256 << 10;
constexpr size_t SMALL_HASH_MAX_N = SMALL_HASH_IN_BYTES / 32;  // arity 1
constexpr size_t SMALL_HASH_OUT_BYTES = SMALL_HASH_MAX_N * 32;
static bool small_path_disabled() {
  static const bool off = getenv("GCP_B200_NO_ZEROCOPY") != nullptr;  // measurements
  return off;
}

static int poseidon_host(gcp_ctx* ctx, const void* in, int len, size_t n, void* out, uint8_t* status, int fmt,
                         bool multi) {
  if (!ctx) return GCP_ERR_BAD_ARG;
  std::lock_guard<std::recursive_mutex> lk(ctx->mu);
  DeviceGuard device_guard(ctx->device);
  if (!device_guard.ok) return ctx->fail(GCP_ERR_CUDA, "cudaSetDevice failed");
  StreamGuard guard{ctx};
  if (len < 1 || len > (multi ? 4096 : 16))
    return ctx->fail(GCP_ERR_BAD_ARG, multi ? "the maximum number of inputs supported is 4096" : "bad inputs provided");
  if (n == 0) return GCP_OK;
  if (!in || !out) return ctx->fail(GCP_ERR_BAD_ARG, "null buffer");
  const size_t item_bytes = (size_t)len * 32;
  // Small batches (BASELINE config 1: 1024 Hash2 = 64 KB in, 32 KB out) are latency, not bandwidth: three staged copies, two
  // stream synchronisations and the status memset cost ~75 us around a 154 us kernel.  Up to 256 KB of inputs the kernel
  // reads them from, and writes its results to, page-locked host memory mapped into its address space (zero-copy): one
  // CPU memcpy each way, one launch, one synchronisation.
  if (!multi && n * item_bytes <= SMALL_HASH_IN_BYTES && !small_path_disabled()) {
    if (!ctx->small_h) {
      cudaError_t e = cudaHostAlloc((void**)&ctx->small_h, SMALL_HASH_IN_BYTES + SMALL_HASH_OUT_BYTES + SMALL_HASH_MAX_N,
                                    cudaHostAllocMapped);
      if (e == cudaSuccess) e = cudaHostGetDevicePointer((void**)&ctx->small_d, ctx->small_h, 0);
      if (e != cudaSuccess) {
        cudaGetLastError();
        if (ctx->small_h) cudaFreeHost(ctx->small_h);
        ctx->small_h = ctx->small_d = nullptr;
      }
    }
    if (ctx->small_h) {
      if (fmt != GCP_FMT_CANONICAL && fmt != GCP_FMT_MONTGOMERY) return ctx->fail(GCP_ERR_BAD_ARG, "bad element format");
      char* h_out = ctx->small_h + SMALL_HASH_IN_BYTES;
      char* h_st = h_out + SMALL_HASH_OUT_BYTES;
      memcpy(ctx->small_h, in, n * item_bytes);
      memset(h_st, 0, n);
      cudaStream_t st = ctx->stream[0];
      CU(launch_poseidon(ctx->tab[len + 1], (const u32*)ctx->small_d, (u32*)(ctx->small_d + SMALL_HASH_IN_BYTES),
                         (uint8_t*)(ctx->small_d + SMALL_HASH_IN_BYTES + SMALL_HASH_OUT_BYTES), n, 1, (size_t)len, 0, 1, fmt, fmt,
                         1, st),
         "poseidon kernel");
      ctx->launches++;
      CU(cudaStreamSynchronize(st), "stream sync");
      memcpy(out, h_out, n * 32);
      if (status) memcpy(status, h_st, n);
      return GCP_OK;
    }
  }
  size_t chunk = std::max<size_t>(1, std::min<size_t>(n, ((size_t)256 << 20) / item_bytes));
  int rc = GCP_OK;
  size_t k = 0;
  for (size_t off = 0; off < n && rc == GCP_OK; off += chunk, k++) {
    size_t m = std::min(chunk, n - off);
    int s = (int)(k & 1);
    cudaStream_t st = ctx->stream[s];
    void* d_in = ctx->buf(4 + s * 3 + 0, m * item_bytes);
    void* d_out = ctx->buf(4 + s * 3 + 1, m * 32);
    uint8_t* d_st = (uint8_t*)ctx->buf(4 + s * 3 + 2, m);
    if (!d_in || !d_out || !d_st) return ctx->fail(GCP_ERR_ALLOC, "device allocation failed");
    GCP_TRY(h2d_copy(ctx, d_in, (const char*)in + off * item_bytes, m * item_bytes, st));
    rc = multi ? poseidon_multihash_dev_locked(ctx, d_in, len, m, d_out, d_st, fmt, st, 38 + s * 2)
               : poseidon_hash_dev_locked(ctx, d_in, len, m, d_out, d_st, fmt, st);
    if (rc != GCP_OK) break;
    GCP_TRY(d2h_copy(ctx, (char*)out + off * 32, d_out, m * 32, st));
    if (status) GCP_TRY(d2h_copy(ctx, status + off, d_st, m, st));
  }
  cudaError_t e0 = cudaStreamSynchronize(ctx->stream[0]);
  cudaError_t e1 = cudaStreamSynchronize(ctx->stream[1]);
  if (rc != GCP_OK) return rc;
  if (e0 != cudaSuccess) return ctx->cuda_fail(e0, "stream sync");
  if (e1 != cudaSuccess) return ctx->cuda_fail(e1, "stream sync");
  return GCP_OK;
}

int gcp_poseidon_hash(gcp_ctx* ctx, const void* in, int arity, size_t n, void* out, uint8_t* status, int fmt) {
  return poseidon_host(ctx, in, arity, n, out, status, fmt, false);
}

int gcp_poseidon_multihash(gcp_ctx* ctx, const void* in, int len, size_t n, void* out, uint8_t* status, int fmt) {
  return poseidon_host(ctx, in, len, n, out, status, fmt, true);
}

// ---------------------------------------------------------------------------------------------------
// SMT
// ---------------------------------------------------------------------------------------------------
static int check_fmt(gcp_ctx* ctx, int fmt);

static int smt_check_args(gcp_ctx* ctx, int n_levels, size_t n, const void* roots, const void* siblings,
                          const void* old_keys, const void* old_values, const void* keys, const void* values,
                          const uint8_t* flags, const uint8_t* status, int fmt) {
  if (n_levels < 2 || n_levels > 253) return ctx->fail(GCP_ERR_BAD_ARG, "n_levels must be in [2, 253]");
  if (fmt != GCP_FMT_CANONICAL && fmt != GCP_FMT_MONTGOMERY) return ctx->fail(GCP_ERR_BAD_ARG, "bad element format");
  if (n == 0) return GCP_OK;
  if (!roots || !siblings || !keys || !values || !flags || !status) return ctx->fail(GCP_ERR_BAD_ARG, "null buffer");
  if ((old_keys == nullptr) != (old_values == nullptr))
    return ctx->fail(GCP_ERR_BAD_ARG, "old_keys and old_values must both be given or both be NULL");
  return GCP_OK;
}

// one scratch slot per verifier call: leaf hashes | perm | lidx | info | hist | cursor
static size_t smt_scratch_bytes(size_t n) {
  const size_t off_perm = n * 32, off_lidx = off_perm + n * 4, off_info = off_lidx + ((n * 2 + 15) & ~(size_t)15);
  return off_info + ((n + 15) & ~(size_t)15) + 1024 + 1040;
}

static int smt_verify_dev_locked(gcp_ctx* ctx, int n_levels, size_t n, const void* d_roots, int shared_root,
                                 const void* d_siblings, const void* d_old_keys, const void* d_old_values,
                                 const uint8_t* d_is_old0, const void* d_keys, const void* d_values,
                                 const uint8_t* d_fnc, const uint8_t* d_enabled, uint8_t* d_flags, uint8_t* d_status,
                                 void* d_out_roots, int fmt, cudaStream_t st, int leaf_slot, int leaf_form = 0) {
  int rc = smt_check_args(ctx, n_levels, n, d_roots, d_siblings, d_old_keys, d_old_values, d_keys, d_values, d_flags,
                          d_status, fmt);
  if (rc != GCP_OK || n == 0) return rc;
  SmtArgs a;
  a.n_levels = n_levels;
  a.n = n;
  a.roots = (const u32*)d_roots;
  a.root_stride = shared_root ? 0 : 8;
  a.siblings = (const u32*)d_siblings;
  a.old_keys = (const u32*)d_old_keys;
  a.old_values = (const u32*)d_old_values;
  a.is_old0 = d_is_old0;
  a.keys = (const u32*)d_keys;
  a.values = (const u32*)d_values;
  a.fnc = d_fnc;
  a.enabled = d_enabled;
  // one scratch slot: leaf hashes | perm | lidx | info | hist | cursor
  const size_t off_perm = n * 32, off_lidx = off_perm + n * 4, off_info = off_lidx + ((n * 2 + 15) & ~(size_t)15);
  const size_t off_hist = off_info + ((n + 15) & ~(size_t)15), off_cur = off_hist + 1024, total = off_cur + 1040;
  char* scratch = (char*)ctx->buf(leaf_slot, total);
  if (!scratch) return ctx->fail(GCP_ERR_ALLOC, "device allocation failed (smt scratch)");
  a.leaf = (u32*)scratch;
  SmtScratch sc;
  sc.perm = (u32*)(scratch + off_perm);
  sc.lidx = (u16*)(scratch + off_lidx);
  sc.info = (u8*)(scratch + off_info);
  sc.hist = (u32*)(scratch + off_hist);
  sc.cursor = (u32*)(scratch + off_cur);
  if (n > 0xffffffffull) return ctx->fail(GCP_ERR_BAD_ARG, "at most 2^32 - 1 proofs per call");
  a.flags = d_flags;
  a.status = d_status;
  a.out_roots = (u32*)d_out_roots;
  a.mont = fmt;
  a.leaf_hash_form = leaf_form;
  a.hasher = ctx->smt_hasher;
  a.hkeys = ctx->d_p2_keys;
  CU(launch_smt_verify(a, sc, ctx->sm_count, st), "smt kernels");
  ctx->launches += 5;
  return GCP_OK;
}

// Proof-streaming scan on its own (the HBM-bound pass of the verifier): per proof lidx = number of path levels that
// carry a hash (1 + index of the last non-zero sibling among [0, n-2]) and an info byte (bit 0: siblings[n-1] == 0,
// bit 1: every sibling canonical).
int gcp_smt_scan_dev(gcp_ctx* ctx, int n_levels, size_t n, const void* d_siblings, uint16_t* d_lidx, uint8_t* d_info,
                     void* stream) {
  if (!ctx) return GCP_ERR_BAD_ARG;
  std::lock_guard<std::recursive_mutex> lk(ctx->mu);
  DeviceGuard device_guard(ctx->device);
  if (!device_guard.ok) return ctx->fail(GCP_ERR_CUDA, "cudaSetDevice failed");
  if (n_levels < 2 || n_levels > 253) return ctx->fail(GCP_ERR_BAD_ARG, "n_levels must be in [2, 253]");
  if (n == 0) return GCP_OK;
  if (!d_siblings || !d_lidx || !d_info) return ctx->fail(GCP_ERR_BAD_ARG, "null buffer");
  CU(launch_smt_scan((const u32*)d_siblings, n, n_levels, d_lidx, d_info, nullptr, ctx->sm_count, (cudaStream_t)stream),
     "smt scan kernel");
  ctx->launches++;
  return GCP_OK;
}

int gcp_smt_verify_dev(gcp_ctx* ctx, int n_levels, size_t n, const void* d_roots, int shared_root,
                       const void* d_siblings, const void* d_old_keys, const void* d_old_values,
                       const uint8_t* d_is_old0, const void* d_keys, const void* d_values, const uint8_t* d_fnc,
                       const uint8_t* d_enabled, uint8_t* d_out_flags, uint8_t* d_out_status, void* d_out_roots,
                       int fmt, void* stream) {
  if (!ctx) return GCP_ERR_BAD_ARG;
  std::lock_guard<std::recursive_mutex> lk(ctx->mu);
  DeviceGuard device_guard(ctx->device);
  if (!device_guard.ok) return ctx->fail(GCP_ERR_CUDA, "cudaSetDevice failed");
  return smt_verify_dev_locked(ctx, n_levels, n, d_roots, shared_root, d_siblings, d_old_keys, d_old_values, d_is_old0,
                               d_keys, d_values, d_fnc, d_enabled, d_out_flags, d_out_status, d_out_roots, fmt,
                               (cudaStream_t)stream, 2);
}

int gcp_smt_verify_with_leaf_hash_dev(gcp_ctx* ctx, int n_levels, size_t n, const void* d_roots, int shared_root,
                                      const void* d_siblings, const void* d_old_keys, const void* d_hash1_old,
                                      const uint8_t* d_is_old0, const void* d_keys, const void* d_hash1_new,
                                      const uint8_t* d_fnc, const uint8_t* d_enabled, uint8_t* d_out_flags,
                                      uint8_t* d_out_status, void* d_out_roots, int fmt, void* stream) {
  if (!ctx) return GCP_ERR_BAD_ARG;
  std::lock_guard<std::recursive_mutex> lk(ctx->mu);
  DeviceGuard device_guard(ctx->device);
  if (!device_guard.ok) return ctx->fail(GCP_ERR_CUDA, "cudaSetDevice failed");
  return smt_verify_dev_locked(ctx, n_levels, n, d_roots, shared_root, d_siblings, d_old_keys, d_hash1_old, d_is_old0,
                               d_keys, d_hash1_new, d_fnc, d_enabled, d_out_flags, d_out_status, d_out_roots, fmt,
                               (cudaStream_t)stream, 2, 1);
}

// Hash1 of the tree (tree/smt/hash.go:10-19): H(key, values..., 1); rows are assembled on the device (scratch slot),
// the hash is the batch Poseidon kernel with arity n_values + 2.
static int smt_leaf_hash_dev_locked(gcp_ctx* ctx, const void* d_keys, const void* d_values, int n_values, size_t n,
                                    void* d_out, uint8_t* d_status, int fmt, cudaStream_t st, int rows_slot) {
  if (n_values < 0 || n_values > 14) return ctx->fail(GCP_ERR_BAD_ARG, "bad inputs provided");  // poseidon.go:41-43 via hFn
  int rc = check_fmt(ctx, fmt);
  if (rc != GCP_OK || n == 0) return rc;
  if (!d_keys || !d_out || (n_values && !d_values)) return ctx->fail(GCP_ERR_BAD_ARG, "null buffer");
  const int arity = n_values + 2;
  u32* rows = (u32*)ctx->buf(rows_slot, n * (size_t)arity * 32);
  if (!rows) return ctx->fail(GCP_ERR_ALLOC, "device allocation failed");
  if (ctx->smt_hasher == GCP_HASHER_POSEIDON2 && (arity != 2 && arity != 3))
    return ctx->fail(GCP_ERR_BAD_ARG, "poseidon2: need 2 or 3 limbs");  // native.go:31-33 through utils.Poseidon2Hasher
  CU(launch_smt_leaf_rows((const u32*)d_keys, (const u32*)d_values, n_values, n, rows, fmt, st), "leaf rows kernel");
  ctx->launches++;
  if (ctx->smt_hasher == GCP_HASHER_POSEIDON2) {
    if (!d_status) return ctx->fail(GCP_ERR_BAD_ARG, "null buffer");
    CU(launch_poseidon2_hash(ctx->d_p2_keys, rows, arity, n, (u32*)d_out, d_status, fmt, st), "poseidon2 hash kernel");
    ctx->launches++;
    return GCP_OK;
  }
  return poseidon_hash_dev_locked(ctx, rows, arity, n, d_out, d_status, fmt, st);
}

int gcp_smt_leaf_hash_dev(gcp_ctx* ctx, const void* d_keys, const void* d_values, int n_values, size_t n, void* d_out,
                          uint8_t* d_status, int fmt, void* stream) {
  if (!ctx) return GCP_ERR_BAD_ARG;
  std::lock_guard<std::recursive_mutex> lk(ctx->mu);
  DeviceGuard device_guard(ctx->device);
  if (!device_guard.ok) return ctx->fail(GCP_ERR_CUDA, "cudaSetDevice failed");
  return smt_leaf_hash_dev_locked(ctx, d_keys, d_values, n_values, n, d_out, d_status, fmt, (cudaStream_t)stream, 100);
}

int gcp_smt_leaf_hash(gcp_ctx* ctx, const void* keys, const void* values, int n_values, size_t n, void* out, uint8_t* status,
                      int fmt) {
  if (!ctx) return GCP_ERR_BAD_ARG;
  std::lock_guard<std::recursive_mutex> lk(ctx->mu);
  DeviceGuard device_guard(ctx->device);
  if (!device_guard.ok) return ctx->fail(GCP_ERR_CUDA, "cudaSetDevice failed");
  StreamGuard guard{ctx};
  if (n_values < 0 || n_values > 14) return ctx->fail(GCP_ERR_BAD_ARG, "bad inputs provided");
  if (n == 0) return check_fmt(ctx, fmt);
  if (!keys || !out || (n_values && !values)) return ctx->fail(GCP_ERR_BAD_ARG, "null buffer");
  const size_t chunk = std::min<size_t>(n, (size_t)1 << 20), vb = (size_t)n_values * 32;
  size_t k = 0;
  for (size_t off = 0; off < n; off += chunk, k++) {
    const size_t m = std::min(chunk, n - off);
    const int s = (int)(k & 1);
    cudaStream_t st = ctx->stream[s];
    void* d_k = ctx->buf(4 + s * 3 + 0, chunk * 32);
    void* d_v = ctx->buf(4 + s * 3 + 1, std::max<size_t>(chunk * vb, 16));
    void* d_o = ctx->buf(4 + s * 3 + 2, chunk * 33);  // digests, then the status bytes
    if (!d_k || !d_v || !d_o) return ctx->fail(GCP_ERR_ALLOC, "device allocation failed");
    uint8_t* d_st = (uint8_t*)d_o + chunk * 32;
    GCP_TRY(h2d_copy(ctx, d_k, (const char*)keys + off * 32, m * 32, st));
    if (vb) GCP_TRY(h2d_copy(ctx, d_v, (const char*)values + off * vb, m * vb, st));
    GCP_TRY(smt_leaf_hash_dev_locked(ctx, d_k, d_v, n_values, m, d_o, d_st, fmt, st, 101 + s));
    GCP_TRY(d2h_copy(ctx, (char*)out + off * 32, d_o, m * 32, st));
    if (status) GCP_TRY(d2h_copy(ctx, status + off, d_st, m, st));
  }
  CU(cudaStreamSynchronize(ctx->stream[0]), "stream sync");
  CU(cudaStreamSynchronize(ctx->stream[1]), "stream sync");
  return GCP_OK;
}

// Host-buffer verifier: `siblings` dense (Assignment.Siblings rows), or NULL with arbo-packed proofs in
// `packed` / `offsets` that are expanded on the device (smt_unpack_kernel) chunk by chunk.
static int smt_verify_host(gcp_ctx* ctx, int n_levels, size_t n, const void* roots, int shared_root, const void* siblings,
                           const uint8_t* packed, const uint64_t* offsets, const void* old_keys, const void* old_values,
                           const uint8_t* is_old0, const void* keys, const void* values, const uint8_t* fnc,
                           const uint8_t* enabled, uint8_t* out_flags, uint8_t* out_status, void* out_roots, int fmt,
                           int leaf_form = 0) {
  if (!ctx) return GCP_ERR_BAD_ARG;
  std::lock_guard<std::recursive_mutex> lk(ctx->mu);
  DeviceGuard device_guard(ctx->device);
  if (!device_guard.ok) return ctx->fail(GCP_ERR_CUDA, "cudaSetDevice failed");
  StreamGuard guard{ctx};
  const bool is_packed = siblings == nullptr;
  int rc = smt_check_args(ctx, n_levels, n, roots, is_packed ? (const void*)packed : siblings, old_keys, old_values, keys,
                          values, out_flags, out_status, fmt);
  if (rc != GCP_OK || n == 0) return rc;
  if (is_packed && !offsets) return ctx->fail(GCP_ERR_BAD_ARG, "null buffer");
  const size_t sib_bytes = (size_t)n_levels * 32;
  ChunkPlan plan(n, smt_path_wave_items(ctx->sm_count), ((size_t)1 << 30) / sib_bytes, true);
  // a shared root is uploaded once
  void* d_shared_root = nullptr;
  if (shared_root) {
    d_shared_root = ctx->buf(3, 32);
    if (!d_shared_root) return ctx->fail(GCP_ERR_ALLOC, "device allocation failed");
    CU(cudaMemcpy(d_shared_root, roots, 32, cudaMemcpyHostToDevice), "H2D root");
  }
  size_t k = 0, off = 0;
  for (size_t m = plan.next(); m != 0 && rc == GCP_OK; off += m, m = plan.next(), k++) {
    int s = (int)(k & 1);
    cudaStream_t st = ctx->stream[s];
    const int b = 10 + s * 14;
    // sized for the largest chunk of the plan from the start: growing a slot later would cudaFree under running work
    const size_t cap = plan.largest();
    if (!ctx->buf(b + 12, smt_scratch_bytes(cap))) return ctx->fail(GCP_ERR_ALLOC, "device allocation failed");
    void* d_sib = ctx->buf(b + 0, cap * sib_bytes);
    void* d_roots = shared_root ? d_shared_root : ctx->buf(b + 1, cap * 32);
    void* d_keys = ctx->buf(b + 2, cap * 32);
    void* d_vals = ctx->buf(b + 3, cap * 32);
    void* d_okeys = old_keys ? ctx->buf(b + 4, cap * 32) : nullptr;
    void* d_ovals = old_keys ? ctx->buf(b + 5, cap * 32) : nullptr;
    uint8_t* d_is0 = is_old0 ? (uint8_t*)ctx->buf(b + 6, cap) : nullptr;
    uint8_t* d_fnc = fnc ? (uint8_t*)ctx->buf(b + 7, cap) : nullptr;
    uint8_t* d_en = enabled ? (uint8_t*)ctx->buf(b + 8, cap) : nullptr;
    uint8_t* d_flags = (uint8_t*)ctx->buf(b + 9, cap);
    uint8_t* d_status = (uint8_t*)ctx->buf(b + 10, cap);
    void* d_oroots = out_roots ? ctx->buf(b + 11, cap * 32) : nullptr;
    if (!d_sib || !d_roots || !d_keys || !d_vals || !d_flags || !d_status || (old_keys && (!d_okeys || !d_ovals)) ||
        (is_old0 && !d_is0) || (fnc && !d_fnc) || (enabled && !d_en) || (out_roots && !d_oroots))
      return ctx->fail(GCP_ERR_ALLOC, "device allocation failed");
    uint8_t* d_bad = nullptr;
    if (is_packed) {
      const uint64_t pbeg = offsets[off], pend = offsets[off + m];
      if (pend < pbeg) {
        rc = ctx->fail(GCP_ERR_BAD_ARG, "packed offsets must be non-decreasing");
        break;
      }
      const size_t pbytes = (size_t)(pend - pbeg);
      uint8_t* d_packed = (uint8_t*)ctx->buf(80 + s * 3 + 0, pbytes + 4);  // + 4: the aligned word of a last odd byte
      uint64_t* d_off = (uint64_t*)ctx->buf(80 + s * 3 + 1, (m + 1) * 8);
      d_bad = (uint8_t*)ctx->buf(80 + s * 3 + 2, m);
      if (!d_packed || !d_off || !d_bad) return ctx->fail(GCP_ERR_ALLOC, "device allocation failed");
      if (pbytes) GCP_TRY(h2d_copy(ctx, d_packed, packed + pbeg, pbytes, st));
      GCP_TRY(h2d_copy(ctx, d_off, offsets + off, (m + 1) * 8, st));
      CU(launch_smt_unpack(d_packed, d_off, pbeg, pbytes, m, n_levels, (u32*)d_sib, d_bad, fmt, st), "smt unpack kernel");
      ctx->launches++;
    } else {
      rc = h2d_copy(ctx, d_sib, (const char*)siblings + off * sib_bytes, m * sib_bytes, st);
      if (rc != GCP_OK) break;
    }
    if (!shared_root) GCP_TRY(h2d_copy(ctx, d_roots, (const char*)roots + off * 32, m * 32, st));
    GCP_TRY(h2d_copy(ctx, d_keys, (const char*)keys + off * 32, m * 32, st));
    GCP_TRY(h2d_copy(ctx, d_vals, (const char*)values + off * 32, m * 32, st));
    if (old_keys) {
      GCP_TRY(h2d_copy(ctx, d_okeys, (const char*)old_keys + off * 32, m * 32, st));
      GCP_TRY(h2d_copy(ctx, d_ovals, (const char*)old_values + off * 32, m * 32, st));
    }
    if (is_old0) GCP_TRY(h2d_copy(ctx, d_is0, is_old0 + off, m, st));
    if (fnc) GCP_TRY(h2d_copy(ctx, d_fnc, fnc + off, m, st));
    if (enabled) GCP_TRY(h2d_copy(ctx, d_en, enabled + off, m, st));
    rc = smt_verify_dev_locked(ctx, n_levels, m, d_roots, shared_root, d_sib, d_okeys, d_ovals, d_is0, d_keys, d_vals,
                               d_fnc, d_en, d_flags, d_status, d_oroots, fmt, st, b + 12, leaf_form);
    if (rc != GCP_OK) break;
    if (is_packed) {
      CU(launch_smt_apply_bad(d_bad, m, d_flags, d_status, (u32*)d_oroots, st), "smt apply-bad kernel");
      ctx->launches++;
    }
    GCP_TRY(d2h_copy(ctx, out_flags + off, d_flags, m, st));
    GCP_TRY(d2h_copy(ctx, out_status + off, d_status, m, st));
    if (out_roots) GCP_TRY(d2h_copy(ctx, (char*)out_roots + off * 32, d_oroots, m * 32, st));
  }
  cudaError_t e0 = cudaStreamSynchronize(ctx->stream[0]);
  cudaError_t e1 = cudaStreamSynchronize(ctx->stream[1]);
  if (rc != GCP_OK) return rc;
  if (e0 != cudaSuccess) return ctx->cuda_fail(e0, "stream sync");
  if (e1 != cudaSuccess) return ctx->cuda_fail(e1, "stream sync");
  return GCP_OK;
}

int gcp_smt_verify(gcp_ctx* ctx, int n_levels, size_t n, const void* roots, int shared_root, const void* siblings,
                   const void* old_keys, const void* old_values, const uint8_t* is_old0, const void* keys,
                   const void* values, const uint8_t* fnc, const uint8_t* enabled, uint8_t* out_flags,
                   uint8_t* out_status, void* out_roots, int fmt) {
  if (ctx && n && !siblings) return ctx->fail(GCP_ERR_BAD_ARG, "null buffer");
  return smt_verify_host(ctx, n_levels, n, roots, shared_root, siblings, nullptr, nullptr, old_keys, old_values, is_old0,
                         keys, values, fnc, enabled, out_flags, out_status, out_roots, fmt);
}

int gcp_smt_verify_with_leaf_hash(gcp_ctx* ctx, int n_levels, size_t n, const void* roots, int shared_root,
                                  const void* siblings, const void* old_keys, const void* hash1_old, const uint8_t* is_old0,
                                  const void* keys, const void* hash1_new, const uint8_t* fnc, const uint8_t* enabled,
                                  uint8_t* out_flags, uint8_t* out_status, void* out_roots, int fmt) {
  if (ctx && n && !siblings) return ctx->fail(GCP_ERR_BAD_ARG, "null buffer");
  return smt_verify_host(ctx, n_levels, n, roots, shared_root, siblings, nullptr, nullptr, old_keys, hash1_old, is_old0, keys,
                         hash1_new, fnc, enabled, out_flags, out_status, out_roots, fmt, 1);
}

int gcp_smt_verify_packed(gcp_ctx* ctx, int n_levels, size_t n, const void* roots, int shared_root,
                          const uint8_t* packed, const uint64_t* offsets, const void* old_keys, const void* old_values,
                          const uint8_t* is_old0, const void* keys, const void* values, const uint8_t* fnc,
                          const uint8_t* enabled, uint8_t* out_flags, uint8_t* out_status, void* out_roots, int fmt) {
  if (ctx && n && (!packed || !offsets)) return ctx->fail(GCP_ERR_BAD_ARG, "null buffer");
  return smt_verify_host(ctx, n_levels, n, roots, shared_root, nullptr, packed, offsets, old_keys, old_values, is_old0,
                         keys, values, fnc, enabled, out_flags, out_status, out_roots, fmt);
}

int gcp_smt_unpack_siblings_dev(gcp_ctx* ctx, int n_levels, size_t n, const uint8_t* d_packed, size_t packed_bytes,
                                const uint64_t* d_offsets, void* d_siblings, uint8_t* d_bad, int fmt, void* stream) {
  if (!ctx) return GCP_ERR_BAD_ARG;
  std::lock_guard<std::recursive_mutex> lk(ctx->mu);
  DeviceGuard device_guard(ctx->device);
  if (!device_guard.ok) return ctx->fail(GCP_ERR_CUDA, "cudaSetDevice failed");
  if (n_levels < 2 || n_levels > 253) return ctx->fail(GCP_ERR_BAD_ARG, "n_levels must be in [2, 253]");
  if (fmt != GCP_FMT_CANONICAL && fmt != GCP_FMT_MONTGOMERY) return ctx->fail(GCP_ERR_BAD_ARG, "bad element format");
  if (n == 0) return GCP_OK;
  if (!d_packed || !d_offsets || !d_siblings || !d_bad) return ctx->fail(GCP_ERR_BAD_ARG, "null buffer");
  CU(launch_smt_unpack(d_packed, d_offsets, 0, packed_bytes, n, n_levels, (u32*)d_siblings, d_bad, fmt,
                       (cudaStream_t)stream),
     "smt unpack kernel");
  ctx->launches++;
  return GCP_OK;
}

int gcp_smt_verify_inclusion(gcp_ctx* ctx, int n_levels, size_t n, const void* roots, int shared_root,
                             const void* siblings, const void* keys, const void* values, uint8_t* out_flags,
                             uint8_t* out_status, void* out_roots, int fmt) {
  // verifier.go:29-43: Verifier(enabled=1, root, siblings, key, value, isOld0=0, key, value, fnc=0)
  return gcp_smt_verify(ctx, n_levels, n, roots, shared_root, siblings, nullptr, nullptr, nullptr, keys, values, nullptr,
                        nullptr, out_flags, out_status, out_roots, fmt);
}

int gcp_smt_verify_exclusion(gcp_ctx* ctx, int n_levels, size_t n, const void* roots, int shared_root,
                             const void* siblings, const void* old_keys, const void* old_values,
                             const uint8_t* is_old0, const void* keys, uint8_t* out_flags, uint8_t* out_status,
                             void* out_roots, int fmt) {
  // verifier.go:66-81: Verifier(enabled=1, root, siblings, oldKey, oldValue, isOld0, key, value=0, fnc=1)
  if (!ctx) return GCP_ERR_BAD_ARG;
  if (n == 0) return GCP_OK;
  std::vector<uint8_t> ones(n, 1);
  std::vector<uint8_t> zeros(n * 32, 0);  // value = 0 in either element format
  return gcp_smt_verify(ctx, n_levels, n, roots, shared_root, siblings, old_keys, old_values, is_old0, keys, zeros.data(),
                        ones.data(), nullptr, out_flags, out_status, out_roots, fmt);
}


// smt.Processor (tree/smt/processor.go:10-72) over a batch
static int smt_process_check(gcp_ctx* ctx, int n_levels, size_t n, const void* a, const void* b, const void* c,
                             const void* d, const void* e, const void* f, const void* g, const void* h, const void* i,
                             const void* j, const void* k, int fmt) {
  if (n_levels < 2 || n_levels > 253) return ctx->fail(GCP_ERR_BAD_ARG, "n_levels must be in [2, 253]");
  if (fmt != GCP_FMT_CANONICAL && fmt != GCP_FMT_MONTGOMERY) return ctx->fail(GCP_ERR_BAD_ARG, "bad element format");
  if (n && (!a || !b || !c || !d || !e || !f || !g || !h || !i || !j || !k)) return ctx->fail(GCP_ERR_BAD_ARG, "null buffer");
  return GCP_OK;
}

static int smt_process_dev_locked(gcp_ctx* ctx, int n_levels, size_t n, const void* d_old_roots, const void* d_siblings,
                                  const void* d_old_keys, const void* d_old_values, const uint8_t* d_is_old0,
                                  const void* d_new_keys, const void* d_new_values, const uint8_t* d_fnc0,
                                  const uint8_t* d_fnc1, void* d_new_roots, uint8_t* d_status, int fmt, cudaStream_t st,
                                  int leaf_form, int scratch_slot) {
  int rc = smt_process_check(ctx, n_levels, n, d_old_roots, d_siblings, d_old_keys, d_old_values, d_is_old0, d_new_keys,
                             d_new_values, d_fnc0, d_fnc1, d_new_roots, d_status, fmt);
  if (rc != GCP_OK || n == 0) return rc;
  SmtProcessArgs a;
  a.n_levels = n_levels;
  a.n = n;
  a.old_roots = (const u32*)d_old_roots;
  a.siblings = (const u32*)d_siblings;
  a.old_keys = (const u32*)d_old_keys;
  a.old_values = (const u32*)d_old_values;
  a.is_old0 = d_is_old0;
  a.new_keys = (const u32*)d_new_keys;
  a.new_values = (const u32*)d_new_values;
  a.fnc0 = d_fnc0;
  a.fnc1 = d_fnc1;
  a.new_roots = (u32*)d_new_roots;
  a.status = d_status;
  a.mont = fmt;
  a.leaf_hash_form = leaf_form;
  a.hasher = ctx->smt_hasher;
  a.hkeys = ctx->d_p2_keys;
  // from 1024 transitions up: the verifier's pipeline (scan, sort by path length, prep, two-chain path kernel); below that
  // (and with GCP_B200_PROCESS_NAIVE set, for measurements) one thread per transition
  static const bool naive = getenv("GCP_B200_PROCESS_NAIVE") != nullptr;
  if (n < 1024 || naive) {
    CU(launch_smt_process(a, nullptr, nullptr, nullptr, ctx->sm_count, st), "smt process kernel");
    ctx->launches++;
    return GCP_OK;
  }
  if (n > 0xffffffffull) return ctx->fail(GCP_ERR_BAD_ARG, "at most 2^32 - 1 transitions per call");
  const size_t base_bytes = smt_scratch_bytes(n);
  char* scratch = (char*)ctx->buf(scratch_slot, base_bytes + n * 32);
  if (!scratch) return ctx->fail(GCP_ERR_ALLOC, "device allocation failed (smt scratch)");
  const size_t off_perm = n * 32, off_lidx = off_perm + n * 4, off_info = off_lidx + ((n * 2 + 15) & ~(size_t)15);
  const size_t off_hist = off_info + ((n + 15) & ~(size_t)15), off_cur = off_hist + 1024;
  SmtScratch sc;
  sc.perm = (u32*)(scratch + off_perm);
  sc.lidx = (u16*)(scratch + off_lidx);
  sc.info = (u8*)(scratch + off_info);
  sc.hist = (u32*)(scratch + off_hist);
  sc.cursor = (u32*)(scratch + off_cur);
  CU(launch_smt_process(a, &sc, (u32*)scratch, (u32*)(scratch + base_bytes), ctx->sm_count, st), "smt process kernels");
  ctx->launches += 5;
  return GCP_OK;
}

int gcp_smt_process_dev(gcp_ctx* ctx, int n_levels, size_t n, const void* d_old_roots, const void* d_siblings,
                        const void* d_old_keys, const void* d_old_values, const uint8_t* d_is_old0,
                        const void* d_new_keys, const void* d_new_values, const uint8_t* d_fnc0, const uint8_t* d_fnc1,
                        void* d_new_roots, uint8_t* d_status, int fmt, void* stream) {
  if (!ctx) return GCP_ERR_BAD_ARG;
  std::lock_guard<std::recursive_mutex> lk(ctx->mu);
  DeviceGuard device_guard(ctx->device);
  if (!device_guard.ok) return ctx->fail(GCP_ERR_CUDA, "cudaSetDevice failed");
  return smt_process_dev_locked(ctx, n_levels, n, d_old_roots, d_siblings, d_old_keys, d_old_values, d_is_old0, d_new_keys,
                                d_new_values, d_fnc0, d_fnc1, d_new_roots, d_status, fmt, (cudaStream_t)stream, 0, 103);
}

int gcp_smt_process_with_leaf_hash_dev(gcp_ctx* ctx, int n_levels, size_t n, const void* d_old_roots, const void* d_siblings,
                                       const void* d_old_keys, const void* d_hash1_old, const uint8_t* d_is_old0,
                                       const void* d_new_keys, const void* d_hash1_new, const uint8_t* d_fnc0,
                                       const uint8_t* d_fnc1, void* d_new_roots, uint8_t* d_status, int fmt, void* stream) {
  if (!ctx) return GCP_ERR_BAD_ARG;
  std::lock_guard<std::recursive_mutex> lk(ctx->mu);
  DeviceGuard device_guard(ctx->device);
  if (!device_guard.ok) return ctx->fail(GCP_ERR_CUDA, "cudaSetDevice failed");
  return smt_process_dev_locked(ctx, n_levels, n, d_old_roots, d_siblings, d_old_keys, d_hash1_old, d_is_old0, d_new_keys,
                                d_hash1_new, d_fnc0, d_fnc1, d_new_roots, d_status, fmt, (cudaStream_t)stream, 1, 103);
}

// Host-buffer processor: dense sibling rows, or (siblings == NULL) arbo packed proofs expanded on the device.
// post_insert: the packed strings come from GenProof AFTER the add, as in the reference's own flow
// (wrapper_arbo.go:152-172): the last unpacked sibling is dropped where isOld0 == 0 and fnc1 == 0.
static int smt_process_host(gcp_ctx* ctx, int n_levels, size_t n, const void* old_roots, const void* siblings,
                            const uint8_t* packed, const uint64_t* offsets, const void* old_keys,
                            const void* old_values, const uint8_t* is_old0, const void* new_keys, const void* new_values,
                            const uint8_t* fnc0, const uint8_t* fnc1, void* new_roots, uint8_t* status, int fmt,
                            bool post_insert, int leaf_form = 0) {
  if (!ctx) return GCP_ERR_BAD_ARG;
  const bool is_packed = siblings == nullptr;
  std::lock_guard<std::recursive_mutex> call_lk(ctx->mu);  // the scratch slots belong to this call until it returns
  DeviceGuard device_guard(ctx->device);
  if (!device_guard.ok) return ctx->fail(GCP_ERR_CUDA, "cudaSetDevice failed");
  StreamGuard guard{ctx};
  {
    int rc = smt_process_check(ctx, n_levels, n, old_roots, is_packed ? (const void*)packed : siblings, old_keys,
                               old_values, is_old0, new_keys, new_values, fnc0, fnc1, new_roots, status, fmt);
    if (rc != GCP_OK || n == 0) return rc;
    if (is_packed && !offsets) return ctx->fail(GCP_ERR_BAD_ARG, "null buffer");
  }
  const size_t sib_bytes = (size_t)n_levels * 32;
  // chunks alternate between the two streams (slot sets 10.. and 24..), so the copies of one overlap the kernel of the other
  const size_t chunk = std::max<size_t>(1, std::min<size_t>(n, ((size_t)256 << 20) / sib_bytes));
  size_t k = 0;
  for (size_t off = 0; off < n; off += chunk, k++) {
    const size_t m = std::min(chunk, n - off);
    const int s = (int)(k & 1);
    const int b = 10 + s * 14;
    cudaStream_t st = ctx->stream[s];
    void* d_sib = ctx->buf(b + 0, chunk * sib_bytes);
    void* d_e[5];
    uint8_t* d_b[4];
    for (int q = 0; q < 5; q++) d_e[q] = ctx->buf(b + 1 + q, chunk * 32);
    void* d_out = ctx->buf(b + 6, chunk * 32);
    for (int q = 0; q < 4; q++) d_b[q] = (uint8_t*)ctx->buf(b + 7 + q, chunk);
    if (!d_sib || !d_out || !d_e[0] || !d_e[1] || !d_e[2] || !d_e[3] || !d_e[4] || !d_b[0] || !d_b[1] || !d_b[2] || !d_b[3])
      return ctx->fail(GCP_ERR_ALLOC, "device allocation failed");
    const void* src_e[5] = {old_roots, old_keys, old_values, new_keys, new_values};
    const uint8_t* src_b[3] = {is_old0, fnc0, fnc1};
    for (int q = 0; q < 3; q++) GCP_TRY(h2d_copy(ctx, d_b[q], src_b[q] + off, m, st));
    uint8_t* d_bad = nullptr;
    if (is_packed) {
      const uint64_t pbeg = offsets[off], pend = offsets[off + m];
      if (pend < pbeg) return ctx->fail(GCP_ERR_BAD_ARG, "packed offsets must be non-decreasing");
      const size_t pbytes = (size_t)(pend - pbeg);
      uint8_t* d_packed = (uint8_t*)ctx->buf(80 + s * 3 + 0, pbytes + 4);
      uint64_t* d_off = (uint64_t*)ctx->buf(80 + s * 3 + 1, (chunk + 1) * 8);
      d_bad = (uint8_t*)ctx->buf(80 + s * 3 + 2, chunk);
      if (!d_packed || !d_off || !d_bad) return ctx->fail(GCP_ERR_ALLOC, "device allocation failed");
      if (pbytes) GCP_TRY(h2d_copy(ctx, d_packed, packed + pbeg, pbytes, st));
      GCP_TRY(h2d_copy(ctx, d_off, offsets + off, (m + 1) * 8, st));
      CU(launch_smt_unpack(d_packed, d_off, pbeg, pbytes, m, n_levels, (u32*)d_sib, d_bad, fmt, st,
                           post_insert ? d_b[0] : nullptr, post_insert ? d_b[2] : nullptr),
         "smt unpack kernel");
      ctx->launches++;
    } else {
      GCP_TRY(h2d_copy(ctx, d_sib, (const char*)siblings + off * sib_bytes, m * sib_bytes, st));
    }
    for (int q = 0; q < 5; q++) GCP_TRY(h2d_copy(ctx, d_e[q], (const char*)src_e[q] + off * 32, m * 32, st));
    int rc = smt_process_dev_locked(ctx, n_levels, m, d_e[0], d_sib, d_e[1], d_e[2], d_b[0], d_e[3], d_e[4], d_b[1], d_b[2],
                                    d_out, d_b[3], fmt, st, leaf_form, b + 12);
    if (rc != GCP_OK) return rc;
    if (is_packed) {
      // a proof arbo.UnpackSiblings rejects never reaches the gadget: status 7, new root 0 (flags: the status array twice)
      CU(launch_smt_apply_bad(d_bad, m, d_b[3], d_b[3], (u32*)d_out, st), "smt apply-bad kernel");
      ctx->launches++;
    }
    GCP_TRY(d2h_copy(ctx, (char*)new_roots + off * 32, d_out, m * 32, st));
    GCP_TRY(d2h_copy(ctx, status + off, d_b[3], m, st));
  }
  CU(cudaStreamSynchronize(ctx->stream[0]), "stream sync");
  CU(cudaStreamSynchronize(ctx->stream[1]), "stream sync");
  return GCP_OK;
}

int gcp_smt_process(gcp_ctx* ctx, int n_levels, size_t n, const void* old_roots, const void* siblings,
                    const void* old_keys, const void* old_values, const uint8_t* is_old0, const void* new_keys,
                    const void* new_values, const uint8_t* fnc0, const uint8_t* fnc1, void* new_roots, uint8_t* status,
                    int fmt) {
  if (ctx && n && !siblings) return ctx->fail(GCP_ERR_BAD_ARG, "null buffer");
  return smt_process_host(ctx, n_levels, n, old_roots, siblings, nullptr, nullptr, old_keys, old_values, is_old0,
                          new_keys, new_values, fnc0, fnc1, new_roots, status, fmt, false);
}

int gcp_smt_process_with_leaf_hash(gcp_ctx* ctx, int n_levels, size_t n, const void* old_roots, const void* siblings,
                                   const void* old_keys, const void* hash1_old, const uint8_t* is_old0,
                                   const void* new_keys, const void* hash1_new, const uint8_t* fnc0, const uint8_t* fnc1,
                                   void* new_roots, uint8_t* status, int fmt) {
  if (ctx && n && !siblings) return ctx->fail(GCP_ERR_BAD_ARG, "null buffer");
  return smt_process_host(ctx, n_levels, n, old_roots, siblings, nullptr, nullptr, old_keys, hash1_old, is_old0, new_keys,
                          hash1_new, fnc0, fnc1, new_roots, status, fmt, false, 1);
}

int gcp_smt_process_packed(gcp_ctx* ctx, int n_levels, size_t n, const void* old_roots, const uint8_t* packed,
                           const uint64_t* offsets, const void* old_keys, const void* old_values, const uint8_t* is_old0,
                           const void* new_keys, const void* new_values, const uint8_t* fnc0, const uint8_t* fnc1,
                           void* new_roots, uint8_t* status, int fmt) {
  if (ctx && n && (!packed || !offsets)) return ctx->fail(GCP_ERR_BAD_ARG, "null buffer");
  return smt_process_host(ctx, n_levels, n, old_roots, nullptr, packed, offsets, old_keys, old_values, is_old0, new_keys,
                          new_values, fnc0, fnc1, new_roots, status, fmt, false);
}

int gcp_smt_process_arbo(gcp_ctx* ctx, int n_levels, size_t n, const void* old_roots, const uint8_t* packed,
                         const uint64_t* offsets, const void* old_keys, const void* old_values, const uint8_t* is_old0,
                         const void* new_keys, const void* new_values, const uint8_t* fnc0, const uint8_t* fnc1,
                         void* new_roots, uint8_t* status, int fmt) {
  if (ctx && n && (!packed || !offsets)) return ctx->fail(GCP_ERR_BAD_ARG, "null buffer");
  return smt_process_host(ctx, n_levels, n, old_roots, nullptr, packed, offsets, old_keys, old_values, is_old0, new_keys,
                          new_values, fnc0, fnc1, new_roots, status, fmt, true);
}

// ---------------------------------------------------------------------------------------------------
// ElGamal
// ---------------------------------------------------------------------------------------------------
static int check_fmt(gcp_ctx* ctx, int fmt) {
  if (fmt != GCP_FMT_CANONICAL && fmt != GCP_FMT_MONTGOMERY) return ctx->fail(GCP_ERR_BAD_ARG, "bad element format");
  return GCP_OK;
}
// The entry points that carry curve points also take GCP_COORDS_TE or-ed into the format: points on the wire are in
// iden3 / circom twisted-Edwards coordinates (ecc/format/twistededwards.go:29-48) and are converted inside the kernels.
static int check_fmt_points(gcp_ctx* ctx, int fmt) {
  if (fmt & ~(GCP_FMT_MONTGOMERY | GCP_COORDS_TE)) return ctx->fail(GCP_ERR_BAD_ARG, "bad element format");
  return GCP_OK;
}
static inline int elem_fmt(int fmt) { return fmt & GCP_FMT_MONTGOMERY; }
static inline size_t msg_bytes(int fmt) { return (fmt & GCP_MSG_U64) ? 8 : 32; }  // per message of the fused tallies
static inline int coords_te(int fmt) { return (fmt & GCP_COORDS_TE) ? 1 : 0; }

// Make d_tabPK the table of the given shared public key (64 bytes, host or device memory); `uses` multiplications by it
// (and as many by G) are about to be queued.
static int ensure_pk_table(gcp_ctx* ctx, const void* pk, bool pk_on_device, int fmt, cudaStream_t st, uint64_t uses) {
  fmt &= GCP_FMT_MONTGOMERY | GCP_COORDS_TE;  // what the key's table depends on
  unsigned char host_pk[64];
  if (pk_on_device) {
    CU(cudaMemcpyAsync(host_pk, pk, 64, cudaMemcpyDeviceToHost, st), "D2H public key");
    CU(cudaStreamSynchronize(st), "D2H public key");
  } else {
    memcpy(host_pk, pk, 64);
  }
  int rc = fb_note_g_use(ctx, uses, st);
  if (rc != GCP_OK) return rc;
  gcp_ctx::FbTable& t = ctx->fb[1];
  const bool same = ctx->pk_cached_fmt == fmt && memcmp(host_pk, ctx->pk_cached, 64) == 0;
  t.uses = same ? t.uses + uses : uses;
  const int want = fb_wanted_bits(ctx, t.uses);
  if (same && (want == t.wbits || (want < t.wbits && !ctx->fb_forced_bits))) return GCP_OK;
  ctx->pk_cached_fmt = -1;
  // the previous table may still be in use by work queued on the other stream
  CU(cudaStreamSynchronize(st), "stream sync");
  CU(cudaStreamSynchronize(ctx->stream[0]), "stream sync");
  CU(cudaStreamSynchronize(ctx->stream[1]), "stream sync");
  CU(cudaMemcpyAsync(ctx->d_base_xy, host_pk, 64, cudaMemcpyHostToDevice, st), "H2D public key");
  rc = fb_build_table(ctx, 1, want, elem_fmt(fmt), coords_te(fmt), st);
  if (rc != GCP_OK) return rc;
  memcpy(ctx->pk_cached, host_pk, 64);
  ctx->pk_cached_fmt = fmt;
  return GCP_OK;
}

int gcp_ctx_set_fixed_base_window(gcp_ctx* ctx, int window_bits) {
  if (!ctx) return GCP_ERR_BAD_ARG;
  std::lock_guard<std::recursive_mutex> lk(ctx->mu);
  if (window_bits != 0 && (window_bits < 8 || window_bits > 26))
    return ctx->fail(GCP_ERR_BAD_ARG, "fixed-base window must be 0 (automatic) or 8..26 bits");
  DeviceGuard device_guard(ctx->device);
  if (!device_guard.ok) return ctx->fail(GCP_ERR_CUDA, "cudaSetDevice failed");
  ctx->fb_forced_bits = window_bits;
  ctx->fb_wide_failed = false;
  if (window_bits == 0) return GCP_OK;  // the tables in place stay until their use counts ask for another width
  return fb_note_g_use(ctx, 0, ctx->stream[0]);  // the key's table follows at its next use (ensure_pk_table)
}

int gcp_ctx_fixed_base_window(const gcp_ctx* ctx, int which) {
  if (!ctx || which < 0 || which > 1) return GCP_ERR_BAD_ARG;
  return ctx->fb[which].wbits;
}

static int fixed_base_dev_locked(gcp_ctx* ctx, const void* d_scalars, size_t n, void* d_out, uint8_t* d_status, int fmt,
                                 cudaStream_t st, int xyz_slot) {
  int rc = check_fmt_points(ctx, fmt);
  if (rc != GCP_OK || n == 0) return rc;
  if (!d_scalars || !d_out || !d_status) return ctx->fail(GCP_ERR_BAD_ARG, "null buffer");
  u32* xyz = (u32*)ctx->buf(xyz_slot, n * 96);
  if (!xyz) return ctx->fail(GCP_ERR_ALLOC, "device allocation failed");
  rc = fb_note_g_use(ctx, n, st);
  if (rc != GCP_OK) return rc;
  CU(launch_fixed_base_mul(ctx->d_tabG, (const u32*)d_scalars, n, xyz, d_status, elem_fmt(fmt), st), "fixed-base kernel");
  CU(launch_normalize(xyz, n, (u32*)d_out, d_status, 1, elem_fmt(fmt), st, 24, coords_te(fmt)), "normalize kernel");
  ctx->launches += 2;
  return GCP_OK;
}

static int encrypt_dev_locked(gcp_ctx* ctx, const void* d_pk, int pk_per_item, const void* d_k, const void* d_m, size_t n,
                              void* d_out, uint8_t* d_status, int fmt, cudaStream_t st, int xyz_slot, int vb_slot) {
  int rc = check_fmt_points(ctx, fmt);
  if (rc != GCP_OK || n == 0) return rc;
  if (!d_pk || !d_k || !d_m || !d_out || !d_status) return ctx->fail(GCP_ERR_BAD_ARG, "null buffer");
  u32* xyz = (u32*)ctx->buf(xyz_slot, n * 192);
  if (!xyz) return ctx->fail(GCP_ERR_ALLOC, "device allocation failed");
  if (pk_per_item) {
    u32* scratch = (u32*)ctx->buf(vb_slot, varbase_scratch_bytes(0, n));
    if (!scratch) return ctx->fail(GCP_ERR_ALLOC, "device allocation failed");
    int nl = 0;
    CU(launch_encrypt_per_key(ctx->d_tabG, (const u32*)d_pk, (const u32*)d_k, (const u32*)d_m, n, xyz, d_status, elem_fmt(fmt),
                              scratch, &nl, st, coords_te(fmt)),
       "encrypt kernels");
    ctx->launches += nl - 1;
  } else {
    CU(launch_encrypt_shared(ctx->d_tabG, ctx->d_tabPK, ctx->d_flagPK, (const u32*)d_k, (const u32*)d_m, n, xyz, d_status,
                             elem_fmt(fmt), st),
       "encrypt kernel");
  }
  CU(launch_normalize(xyz, 2 * n, (u32*)d_out, d_status, 2, elem_fmt(fmt), st, 24, coords_te(fmt)), "normalize kernel");
  ctx->launches += 2;
  return GCP_OK;
}

static int add_dev_locked(gcp_ctx* ctx, const void* d_a, const void* d_b, size_t n, void* d_out, uint8_t* d_status,
                          int fmt, cudaStream_t st, int xyz_slot) {
  int rc = check_fmt_points(ctx, fmt);
  if (rc != GCP_OK || n == 0) return rc;
  if (!d_a || !d_b || !d_out || !d_status) return ctx->fail(GCP_ERR_BAD_ARG, "null buffer");
  u32* xyz = (u32*)ctx->buf(xyz_slot, n * 192);
  if (!xyz) return ctx->fail(GCP_ERR_ALLOC, "device allocation failed");
  CU(cudaMemsetAsync(d_status, 0, n, st), "memset status");
  CU(launch_ct_add((const u32*)d_a, (const u32*)d_b, n, xyz, d_status, elem_fmt(fmt), st, coords_te(fmt)), "ciphertext add kernel");
  CU(launch_normalize(xyz, 2 * n, (u32*)d_out, d_status, 2, elem_fmt(fmt), st, 24, coords_te(fmt)), "normalize kernel");
  ctx->launches += 2;
  return GCP_OK;
}

static int tally_dev_locked(gcp_ctx* ctx, const void* d_ct, size_t n_ballots, int n_fields, void* d_out,
                            uint8_t* d_status, int fmt, cudaStream_t st, int slot_base) {
  int rc = check_fmt_points(ctx, fmt);
  if (rc != GCP_OK) return rc;
  if (n_fields < 1 || n_fields > 64) return ctx->fail(GCP_ERR_BAD_ARG, "n_fields must be in [1, 64]");
  if (!d_out || !d_status || (n_ballots && !d_ct)) return ctx->fail(GCP_ERR_BAD_ARG, "null buffer");
  int n_blocks = tally_max_blocks(n_ballots, n_fields, ctx->sm_count);
  const int cols = n_fields * 2;
  u32* partials = (u32*)ctx->buf(slot_base, (size_t)n_blocks * cols * 128);
  u32* bad = (u32*)ctx->buf(slot_base + 1, (size_t)n_fields * 4);
  u32* xyz = (u32*)ctx->buf(slot_base + 2, (size_t)cols * 96);
  if (!partials || !bad || !xyz) return ctx->fail(GCP_ERR_ALLOC, "device allocation failed");
  CU(launch_tally((const u32*)d_ct, n_ballots, n_fields, n_blocks, partials, bad, xyz, d_status, elem_fmt(fmt), st, coords_te(fmt)),
     "tally kernels");
  CU(launch_normalize(xyz, (size_t)cols, (u32*)d_out, d_status, 2, elem_fmt(fmt), st, 24, coords_te(fmt)), "normalize kernel");
  ctx->launches += 3;
  return GCP_OK;
}

int gcp_elgamal_fixed_base_mul_dev(gcp_ctx* ctx, const void* d_scalars, size_t n, void* d_out_points, uint8_t* d_status,
                                   int fmt, void* stream) {
  if (!ctx) return GCP_ERR_BAD_ARG;
  std::lock_guard<std::recursive_mutex> lk(ctx->mu);
  DeviceGuard device_guard(ctx->device);
  if (!device_guard.ok) return ctx->fail(GCP_ERR_CUDA, "cudaSetDevice failed");
  return fixed_base_dev_locked(ctx, d_scalars, n, d_out_points, d_status, fmt, (cudaStream_t)stream, 43);
}

int gcp_elgamal_encrypt_dev(gcp_ctx* ctx, const void* d_pub_key, int pk_per_item, const void* d_k, const void* d_m,
                            size_t n, void* d_out_ct, uint8_t* d_status, int fmt, void* stream) {
  if (!ctx) return GCP_ERR_BAD_ARG;
  std::lock_guard<std::recursive_mutex> lk(ctx->mu);
  DeviceGuard device_guard(ctx->device);
  if (!device_guard.ok) return ctx->fail(GCP_ERR_CUDA, "cudaSetDevice failed");
  if (n && !pk_per_item) {
    if (!d_pub_key) return ctx->fail(GCP_ERR_BAD_ARG, "null buffer");
    int rc = ensure_pk_table(ctx, d_pub_key, true, fmt, (cudaStream_t)stream, n);
    if (rc != GCP_OK) return rc;
  }
  return encrypt_dev_locked(ctx, d_pub_key, pk_per_item, d_k, d_m, n, d_out_ct, d_status, fmt, (cudaStream_t)stream, 43, 98);
}

int gcp_elgamal_add_dev(gcp_ctx* ctx, const void* d_a, const void* d_b, size_t n, void* d_out, uint8_t* d_status, int fmt,
                        void* stream) {
  if (!ctx) return GCP_ERR_BAD_ARG;
  std::lock_guard<std::recursive_mutex> lk(ctx->mu);
  DeviceGuard device_guard(ctx->device);
  if (!device_guard.ok) return ctx->fail(GCP_ERR_CUDA, "cudaSetDevice failed");
  return add_dev_locked(ctx, d_a, d_b, n, d_out, d_status, fmt, (cudaStream_t)stream, 43);
}

int gcp_elgamal_tally_dev(gcp_ctx* ctx, const void* d_ct, size_t n_ballots, int n_fields, void* d_out, uint8_t* d_status,
                          int fmt, void* stream) {
  if (!ctx) return GCP_ERR_BAD_ARG;
  std::lock_guard<std::recursive_mutex> lk(ctx->mu);
  DeviceGuard device_guard(ctx->device);
  if (!device_guard.ok) return ctx->fail(GCP_ERR_CUDA, "cudaSetDevice failed");
  return tally_dev_locked(ctx, d_ct, n_ballots, n_fields, d_out, d_status, fmt, (cudaStream_t)stream, 44);
}

static int encrypt_tally_dev_locked(gcp_ctx* ctx, const void* d_k, const void* d_m, const uint8_t* d_mask, size_t n_ballots, int n_fields,
                                    void* d_out, uint8_t* d_status, int fmt, cudaStream_t st, int slot_base) {
  int n_blocks = tally_max_blocks(n_ballots, n_fields, ctx->sm_count);
  const int cols = n_fields * 2;
  u32* partials = (u32*)ctx->buf(slot_base, (size_t)n_blocks * cols * 128);
  u32* bad = (u32*)ctx->buf(slot_base + 1, (size_t)n_fields * 4);
  u32* xyz = (u32*)ctx->buf(slot_base + 2, (size_t)cols * 96);
  if (!partials || !bad || !xyz) return ctx->fail(GCP_ERR_ALLOC, "device allocation failed");
  CU(launch_encrypt_tally(ctx->d_tabG, ctx->d_tabPK, (const u32*)d_k, (const u32*)d_m, d_mask, n_ballots, n_fields, n_blocks,
                          partials, bad, xyz, d_status, elem_fmt(fmt), st, (int)(msg_bytes(fmt) / 4)),
     "encrypt-tally kernels");
  CU(launch_normalize(xyz, (size_t)cols, (u32*)d_out, d_status, 2, elem_fmt(fmt), st, 24, coords_te(fmt)), "normalize kernel");
  ctx->launches += 3;
  return GCP_OK;
}

static int encrypt_tally_check(gcp_ctx* ctx, const void* pk, const void* k, const void* m, size_t n_ballots,
                               int n_fields, const void* out, const void* status, int fmt) {
  int rc = check_fmt_points(ctx, fmt & ~GCP_MSG_U64);  // the fused tallies also take GCP_MSG_U64
  if (rc != GCP_OK) return rc;
  if (n_fields < 1 || n_fields > 64) return ctx->fail(GCP_ERR_BAD_ARG, "n_fields must be in [1, 64]");
  if (!pk || !out || !status || (n_ballots && (!k || !m))) return ctx->fail(GCP_ERR_BAD_ARG, "null buffer");
  return GCP_OK;
}

int gcp_elgamal_encrypt_tally_dev(gcp_ctx* ctx, const void* d_pub_key, const void* d_k, const void* d_m,
                                  size_t n_ballots, int n_fields, void* d_out, uint8_t* d_status, int fmt,
                                  void* stream) {
  if (!ctx) return GCP_ERR_BAD_ARG;
  std::lock_guard<std::recursive_mutex> lk(ctx->mu);
  DeviceGuard device_guard(ctx->device);
  if (!device_guard.ok) return ctx->fail(GCP_ERR_CUDA, "cudaSetDevice failed");
  int rc = encrypt_tally_check(ctx, d_pub_key, d_k, d_m, n_ballots, n_fields, d_out, d_status, fmt);
  if (rc != GCP_OK) return rc;
  rc = ensure_pk_table(ctx, d_pub_key, true, fmt, (cudaStream_t)stream, (uint64_t)n_ballots * n_fields);
  if (rc != GCP_OK) return rc;
  rc = encrypt_tally_dev_locked(ctx, d_k, d_m, nullptr, n_ballots, n_fields, d_out, d_status, fmt, (cudaStream_t)stream, 44);
  if (rc != GCP_OK) return rc;
  // AssertIsOnCurve(pubKey) (encrypt.go:49): an off-curve or non-canonical key gives status 4 and no result on every field
  CU(launch_tally_status_merge(nullptr, 0, n_fields, 1, ctx->d_flagPK, (u32*)d_out, d_status, (cudaStream_t)stream),
     "tally status kernel");
  ctx->launches++;
  return GCP_OK;
}

// Last step of the three chunked folds (tally, encrypt-tally, ballot batch): the per-chunk partial ciphertexts
// (n_chunks x n_fields, with their statuses, on the device; both pipeline streams drained) are folded into the result.
// `out` / `status` are host buffers, or - for the group exchange (group.cu) - device buffers on this context's device,
// so that a partial tally never leaves the GPU before the all-gather.  pk_flag: the cached key's on-curve flag, or
// nullptr for a plain tally.
static int finish_partials(gcp_ctx* ctx, u32* d_parts, uint8_t* d_part_status, size_t n_chunks, int n_fields, int fmt,
                           const u32* pk_flag, void* out, uint8_t* status, bool out_on_device) {
  fmt &= ~GCP_MSG_U64;  // the partials are ciphertexts
  const size_t ballot_ct = (size_t)n_fields * 128;
  cudaStream_t st = ctx->stream[0];
  u32* d_res = out_on_device ? (u32*)out : (u32*)ctx->buf(66, ballot_ct);
  uint8_t* d_res_status = out_on_device ? status : (uint8_t*)ctx->buf(67, n_fields);
  if (!d_res || !d_res_status) return ctx->fail(GCP_ERR_ALLOC, "device allocation failed");
  if (n_chunks > 1) {
    int rc = tally_dev_locked(ctx, d_parts, n_chunks, n_fields, d_res, d_res_status, fmt, st, 49);
    if (rc != GCP_OK) return rc;
  } else {
    CU(cudaMemcpyAsync(d_res, d_parts, ballot_ct, cudaMemcpyDeviceToDevice, st), "D2D");
  }
  CU(launch_tally_status_merge(d_part_status, (int)n_chunks, n_fields, n_chunks > 1 ? 1 : 0, pk_flag, d_res, d_res_status, st),
     "tally status kernel");
  ctx->launches++;
  if (!out_on_device) {
    GCP_TRY(d2h_copy(ctx, out, d_res, ballot_ct, st));
    GCP_TRY(d2h_copy(ctx, status, d_res_status, n_fields, st));
  }
  CU(cudaStreamSynchronize(st), "stream sync");
  return GCP_OK;
}

// Host form: scalars are streamed in chunks, each chunk is encrypted and reduced on the device, and the per-chunk
// partial ciphertexts are tallied at the end.  status[f] = 4 for every field when the key is off the curve.
static int encrypt_tally_host(gcp_ctx* ctx, const void* pub_key, const void* k, const void* m, size_t n_ballots, int n_fields,
                              void* out, uint8_t* status, int fmt, bool out_on_device) {
  if (!ctx) return GCP_ERR_BAD_ARG;
  std::lock_guard<std::recursive_mutex> lk(ctx->mu);
  DeviceGuard device_guard(ctx->device);
  if (!device_guard.ok) return ctx->fail(GCP_ERR_CUDA, "cudaSetDevice failed");
  StreamGuard guard{ctx};
  int rc = encrypt_tally_check(ctx, pub_key, k, m, n_ballots, n_fields, out, status, fmt);
  if (rc != GCP_OK) return rc;
  rc = ensure_pk_table(ctx, pub_key, false, fmt, ctx->stream[0], (uint64_t)n_ballots * n_fields);
  if (rc != GCP_OK) return rc;
  const size_t ballot_in = (size_t)n_fields * 32, ballot_m = (size_t)n_fields * msg_bytes(fmt), ballot_ct = (size_t)n_fields * 128;
  // Chunk = 1/16 of the call, between 64 MB and 256 MB of k (and as much of m).  Every chunk costs ~0.4 ms of small kernels
  // (second-stage fold, normalisation) and the first chunk's copy is the one nothing hides, so large calls want large
  // chunks and small calls small ones; measured at 2^23 ballots x 8 from page-locked memory
  // (profiles/r02_encrypt_tally_chunk_sweep.jsonl): 16 / 64 / 256 / 512 MB chunks -> 253 / 276 / 286 / 282 M enc/s (309 M resident).
  static const size_t forced_mb = [] {
    const char* env = getenv("GCP_B200_ET_CHUNK_MB");
    long v = env ? atol(env) : 0;
    return (size_t)(v >= 1 && v <= 4096 ? v : 0);
  }();
  const size_t total_bytes = n_ballots * ballot_in;
  const size_t chunk_bytes = forced_mb ? (forced_mb << 20)
                                       : std::min<size_t>((size_t)256 << 20, std::max<size_t>((size_t)64 << 20, total_bytes / 16));
  size_t chunk = std::max<size_t>(1, std::min<size_t>(std::max<size_t>(n_ballots, 1), chunk_bytes / ballot_in));
  size_t n_chunks = n_ballots ? (n_ballots + chunk - 1) / chunk : 1;
  u32* d_parts = (u32*)ctx->buf(64, n_chunks * ballot_ct);
  uint8_t* d_part_status = (uint8_t*)ctx->buf(65, n_chunks * n_fields);
  if (!d_parts || !d_part_status) return ctx->fail(GCP_ERR_ALLOC, "device allocation failed");
  for (size_t c = 0; c < n_chunks; c++) {
    size_t off = c * chunk, cnt = n_ballots ? std::min(chunk, n_ballots - off) : 0;
    int s = (int)(c & 1);
    cudaStream_t st = ctx->stream[s];
    void* dk = ctx->buf(48 + s * 8, std::max<size_t>(std::min(chunk, n_ballots), 1) * ballot_in);
    void* dm = ctx->buf(48 + s * 8 + 4, std::max<size_t>(std::min(chunk, n_ballots), 1) * ballot_m);
    if (!dk || !dm) return ctx->fail(GCP_ERR_ALLOC, "device allocation failed");
    if (cnt) {
      GCP_TRY(h2d_copy(ctx, dk, (const char*)k + off * ballot_in, cnt * ballot_in, st));
      GCP_TRY(h2d_copy(ctx, dm, (const char*)m + off * ballot_m, cnt * ballot_m, st));
    }
    rc = encrypt_tally_dev_locked(ctx, dk, dm, nullptr, cnt, n_fields, (char*)d_parts + c * ballot_ct, d_part_status + c * n_fields,
                                  fmt, st, 48 + s * 8 + 1);
    if (rc != GCP_OK) return rc;
  }
  CU(cudaStreamSynchronize(ctx->stream[0]), "stream sync");
  CU(cudaStreamSynchronize(ctx->stream[1]), "stream sync");
  return finish_partials(ctx, d_parts, d_part_status, n_chunks, n_fields, fmt, ctx->d_flagPK, out, status, out_on_device);
}

int gcp_elgamal_encrypt_tally(gcp_ctx* ctx, const void* pub_key, const void* k, const void* m, size_t n_ballots,
                              int n_fields, void* out, uint8_t* status, int fmt) {
  return encrypt_tally_host(ctx, pub_key, k, m, n_ballots, n_fields, out, status, fmt, false);
}

// Generic host-buffer pipeline for the per-item ElGamal calls.
//   kind 0: fixed-base (in0 = scalars 32 B)            -> 64 B
//   kind 1: encrypt    (in0 = k 32 B, in1 = m 32 B, optional per-item pk 64 B) -> 128 B
//   kind 2: add        (in0 = a 128 B, in1 = b 128 B)  -> 128 B
//   kind 3: neg        (in0 = a 128 B)                 -> 128 B
static int elgamal_host(gcp_ctx* ctx, int kind, const void* pk, int pk_per_item, const void* in0, const void* in1,
                        size_t n, void* out, uint8_t* status, int fmt) {
  if (!ctx) return GCP_ERR_BAD_ARG;
  std::lock_guard<std::recursive_mutex> lk(ctx->mu);
  DeviceGuard device_guard(ctx->device);
  if (!device_guard.ok) return ctx->fail(GCP_ERR_CUDA, "cudaSetDevice failed");
  StreamGuard guard{ctx};
  int rc = check_fmt_points(ctx, fmt);
  if (rc != GCP_OK || n == 0) return rc;
  if (!in0 || !out || !status || ((kind == 1 || kind == 2) && !in1) || (kind == 1 && !pk))
    return ctx->fail(GCP_ERR_BAD_ARG, "null buffer");
  const size_t in0_b = (kind <= 1) ? 32 : 128, in1_b = (kind == 1) ? 32 : (kind == 2 ? 128 : 0);
  const size_t out_b = (kind == 0) ? 64 : 128;
  if (kind == 1 && !pk_per_item) {
    rc = ensure_pk_table(ctx, pk, false, fmt, ctx->stream[0], n);
    if (rc != GCP_OK) return rc;
  }
  // per-item keys run the variable-base window kernel: whole resident waves of it; the other kinds are PCIe-bound and
  // only want a short first chunk (the one copy nothing hides)
  const bool varbase = kind == 1 && pk_per_item;
  ChunkPlan plan(n, varbase ? varbase_wave_items(ctx->sm_count) : (size_t)1 << 18, varbase ? (size_t)1 << 18 : (size_t)1 << 20);
  const size_t cap = plan.largest();
  size_t k = 0, off = 0;
  for (size_t m = plan.next(); m != 0 && rc == GCP_OK; off += m, m = plan.next(), k++) {
    int s = (int)(k & 1);
    cudaStream_t st = ctx->stream[s];
    const int b = 48 + s * 8;
    void* d0 = ctx->buf(b + 0, cap * in0_b);
    void* d1 = in1_b ? ctx->buf(b + 1, cap * in1_b) : nullptr;
    void* dpk = (kind == 1 && pk_per_item) ? ctx->buf(b + 2, cap * 64) : nullptr;
    void* dout = ctx->buf(b + 3, cap * out_b);
    uint8_t* dst = (uint8_t*)ctx->buf(b + 4, cap);
    if (!d0 || (in1_b && !d1) || (kind == 1 && pk_per_item && !dpk) || !dout || !dst || !ctx->buf(b + 5, cap * 192) ||
        (varbase && !ctx->buf(99 + s, varbase_scratch_bytes(0, cap))))
      return ctx->fail(GCP_ERR_ALLOC, "device allocation failed");
    GCP_TRY(h2d_copy(ctx, d0, (const char*)in0 + off * in0_b, m * in0_b, st));
    if (in1_b) GCP_TRY(h2d_copy(ctx, d1, (const char*)in1 + off * in1_b, m * in1_b, st));
    if (dpk) GCP_TRY(h2d_copy(ctx, dpk, (const char*)pk + off * 64, m * 64, st));
    switch (kind) {
      case 0: rc = fixed_base_dev_locked(ctx, d0, m, dout, dst, fmt, st, b + 5); break;
      case 1: rc = encrypt_dev_locked(ctx, pk_per_item ? dpk : (const void*)ctx->d_base_xy, pk_per_item, d0, d1, m, dout, dst, fmt, st, b + 5, 99 + s); break;
      case 2: rc = add_dev_locked(ctx, d0, d1, m, dout, dst, fmt, st, b + 5); break;
      default:
        CU(cudaMemsetAsync(dst, 0, m, st), "memset status");
        CU(launch_ct_neg((const u32*)d0, m, (u32*)dout, dst, st), "neg kernel");
        ctx->launches++;
        break;
    }
    if (rc != GCP_OK) break;
    GCP_TRY(d2h_copy(ctx, (char*)out + off * out_b, dout, m * out_b, st));
    GCP_TRY(d2h_copy(ctx, status + off, dst, m, st));
  }
  cudaError_t e0 = cudaStreamSynchronize(ctx->stream[0]);
  cudaError_t e1 = cudaStreamSynchronize(ctx->stream[1]);
  if (rc != GCP_OK) return rc;
  if (e0 != cudaSuccess) return ctx->cuda_fail(e0, "stream sync");
  if (e1 != cudaSuccess) return ctx->cuda_fail(e1, "stream sync");
  return GCP_OK;
}

int gcp_elgamal_fixed_base_mul(gcp_ctx* ctx, const void* scalars, size_t n, void* out_points, uint8_t* status, int fmt) {
  return elgamal_host(ctx, 0, nullptr, 0, scalars, nullptr, n, out_points, status, fmt);
}
int gcp_elgamal_encrypt(gcp_ctx* ctx, const void* pub_key, int pk_per_item, const void* k, const void* m, size_t n,
                        void* out_ct, uint8_t* status, int fmt) {
  return elgamal_host(ctx, 1, pub_key, pk_per_item, k, m, n, out_ct, status, fmt);
}
int gcp_elgamal_add(gcp_ctx* ctx, const void* a, const void* b, size_t n, void* out, uint8_t* status, int fmt) {
  return elgamal_host(ctx, 2, nullptr, 0, a, b, n, out, status, fmt);
}
int gcp_elgamal_neg(gcp_ctx* ctx, const void* a, size_t n, void* out, uint8_t* status, int fmt) {
  return elgamal_host(ctx, 3, nullptr, 0, a, nullptr, n, out, status, fmt);
}

// Ciphertext.IsEqual / Select (elgamal/ciphertext.go:79-96): element-wise, HBM/PCIe-bound; chunked on the two streams.
//   kind 0: is_equal(a, b) -> flags;  kind 1: select(sel, i1, i2) -> ciphertexts
static int ct_elementwise_host(gcp_ctx* ctx, int kind, const uint8_t* sel, const void* a, const void* b, size_t n, void* out,
                               uint8_t* status) {
  if (!ctx) return GCP_ERR_BAD_ARG;
  std::lock_guard<std::recursive_mutex> lk(ctx->mu);
  DeviceGuard device_guard(ctx->device);
  if (!device_guard.ok) return ctx->fail(GCP_ERR_CUDA, "cudaSetDevice failed");
  StreamGuard guard{ctx};
  if (n == 0) return GCP_OK;
  if (!a || !b || !out || !status || (kind == 1 && !sel)) return ctx->fail(GCP_ERR_BAD_ARG, "null buffer");
  const size_t out_b = kind == 0 ? 1 : 128;
  const size_t chunk = std::min<size_t>(n, (size_t)1 << 20);
  size_t k = 0;
  for (size_t off = 0; off < n; off += chunk, k++) {
    const size_t m = std::min(chunk, n - off);
    const int s = (int)(k & 1);
    cudaStream_t st = ctx->stream[s];
    const int bs = 48 + s * 8;
    void* da = ctx->buf(bs + 0, m * 128);
    void* db = ctx->buf(bs + 1, m * 128);
    uint8_t* dsel = kind == 1 ? (uint8_t*)ctx->buf(bs + 2, m) : nullptr;
    void* dout = ctx->buf(bs + 3, m * out_b);
    uint8_t* dst = (uint8_t*)ctx->buf(bs + 4, m);
    if (!da || !db || !dout || !dst || (kind == 1 && !dsel)) return ctx->fail(GCP_ERR_ALLOC, "device allocation failed");
    GCP_TRY(h2d_copy(ctx, da, (const char*)a + off * 128, m * 128, st));
    GCP_TRY(h2d_copy(ctx, db, (const char*)b + off * 128, m * 128, st));
    if (kind == 1) {
      GCP_TRY(h2d_copy(ctx, dsel, sel + off, m, st));
      CU(launch_ct_select(dsel, (const u32*)da, (const u32*)db, m, (u32*)dout, dst, st), "select kernel");
    } else {
      CU(launch_ct_is_equal((const u32*)da, (const u32*)db, m, (uint8_t*)dout, dst, st), "is-equal kernel");
    }
    ctx->launches++;
    GCP_TRY(d2h_copy(ctx, (char*)out + off * out_b, dout, m * out_b, st));
    GCP_TRY(d2h_copy(ctx, status + off, dst, m, st));
  }
  CU(cudaStreamSynchronize(ctx->stream[0]), "stream sync");
  CU(cudaStreamSynchronize(ctx->stream[1]), "stream sync");
  return GCP_OK;
}

int gcp_elgamal_is_equal(gcp_ctx* ctx, const void* a, const void* b, size_t n, uint8_t* out_flags, uint8_t* status) {
  return ct_elementwise_host(ctx, 0, nullptr, a, b, n, out_flags, status);
}
int gcp_elgamal_select(gcp_ctx* ctx, const uint8_t* sel, const void* i1, const void* i2, size_t n, void* out,
                       uint8_t* status) {
  return ct_elementwise_host(ctx, 1, sel, i1, i2, n, out, status);
}

// Tally over host-resident ciphertexts: chunks are reduced on the device as they arrive; the per-chunk partial
// ciphertexts are themselves tallied at the end (addition is associative, so the result does not depend on chunking).
static int tally_host(gcp_ctx* ctx, const void* ct, size_t n_ballots, int n_fields, void* out, uint8_t* status, int fmt,
                      bool out_on_device) {
  if (!ctx) return GCP_ERR_BAD_ARG;
  std::lock_guard<std::recursive_mutex> lk(ctx->mu);
  DeviceGuard device_guard(ctx->device);
  if (!device_guard.ok) return ctx->fail(GCP_ERR_CUDA, "cudaSetDevice failed");
  StreamGuard guard{ctx};
  int rc = check_fmt_points(ctx, fmt);
  if (rc != GCP_OK) return rc;
  if (n_fields < 1 || n_fields > 64) return ctx->fail(GCP_ERR_BAD_ARG, "n_fields must be in [1, 64]");
  if (!out || !status || (n_ballots && !ct)) return ctx->fail(GCP_ERR_BAD_ARG, "null buffer");
  const size_t ballot_b = (size_t)n_fields * 128;
  // the kernel outruns PCIe 5:1, so the call is one long copy: 64 MB chunks keep the tail after the last copy short
  size_t chunk = std::max<size_t>(1, std::min<size_t>(std::max<size_t>(n_ballots, 1), ((size_t)64 << 20) / ballot_b));
  size_t n_chunks = n_ballots ? (n_ballots + chunk - 1) / chunk : 1;
  u32* d_parts = (u32*)ctx->buf(64, n_chunks * ballot_b);
  uint8_t* d_part_status = (uint8_t*)ctx->buf(65, n_chunks * n_fields);
  if (!d_parts || !d_part_status) return ctx->fail(GCP_ERR_ALLOC, "device allocation failed");
  for (size_t c = 0; c < n_chunks; c++) {
    size_t off = c * chunk, m = n_ballots ? std::min(chunk, n_ballots - off) : 0;
    int s = (int)(c & 1);
    cudaStream_t st = ctx->stream[s];
    void* d_ct = ctx->buf(48 + s * 8, std::max<size_t>(std::min(chunk, n_ballots), 1) * ballot_b);
    if (!d_ct) return ctx->fail(GCP_ERR_ALLOC, "device allocation failed");
    if (m) GCP_TRY(h2d_copy(ctx, d_ct, (const char*)ct + off * ballot_b, m * ballot_b, st));
    rc = tally_dev_locked(ctx, d_ct, m, n_fields, (char*)d_parts + c * ballot_b, d_part_status + c * n_fields, fmt, st,
                          48 + s * 8 + 1);
    if (rc != GCP_OK) return rc;
  }
  CU(cudaStreamSynchronize(ctx->stream[0]), "stream sync");
  CU(cudaStreamSynchronize(ctx->stream[1]), "stream sync");
  return finish_partials(ctx, d_parts, d_part_status, n_chunks, n_fields, fmt, nullptr, out, status, out_on_device);
}

int gcp_elgamal_tally(gcp_ctx* ctx, const void* ct, size_t n_ballots, int n_fields, void* out, uint8_t* status, int fmt) {
  return tally_host(ctx, ct, n_ballots, n_fields, out, status, fmt, false);
}

// ---------------------------------------------------------------------------------------------------
// End-to-end ballot batch (BASELINE config 5): census inclusion proof + ballot encryption + aggregation
// ---------------------------------------------------------------------------------------------------
// Per voter: smt.InclusionVerifier on the census proof, then Encrypt of the voter's n_fields values; the ciphertexts
// of voters whose proof verified (flag 1, status 0) are folded with Ciphertext.Add.  The fold lives in the
// reference's caller (davinci-node); here it is the masked fused encrypt-tally kernel, so no ciphertext is stored.
int gcp_ballot_batch_dev(gcp_ctx* ctx, int n_levels, size_t n_voters, const void* d_roots, int shared_root,
                         const void* d_siblings, const void* d_keys, const void* d_values, const void* d_pub_key,
                         const void* d_k, const void* d_m, int n_fields, uint8_t* d_flags, uint8_t* d_status,
                         void* d_tally, uint8_t* d_tally_status, int fmt, void* stream) {
  if (!ctx) return GCP_ERR_BAD_ARG;
  std::lock_guard<std::recursive_mutex> lk(ctx->mu);
  DeviceGuard device_guard(ctx->device);
  if (!device_guard.ok) return ctx->fail(GCP_ERR_CUDA, "cudaSetDevice failed");
  cudaStream_t st = (cudaStream_t)stream;
  int rc = encrypt_tally_check(ctx, d_pub_key, d_k, d_m, n_voters, n_fields, d_tally, d_tally_status, fmt);
  if (rc != GCP_OK) return rc;
  rc = ensure_pk_table(ctx, d_pub_key, true, fmt, st, (uint64_t)n_voters * n_fields);
  if (rc != GCP_OK) return rc;
  rc = smt_verify_dev_locked(ctx, n_levels, n_voters, d_roots, shared_root, d_siblings, nullptr, nullptr, nullptr, d_keys,
                             d_values, nullptr, nullptr, d_flags, d_status, nullptr, elem_fmt(fmt), st, 2);
  if (rc != GCP_OK) return rc;
  // flags are 0 wherever status != 0 (smt_path_kernel), so the flag array is the admission mask
  rc = encrypt_tally_dev_locked(ctx, d_k, d_m, d_flags, n_voters, n_fields, d_tally, d_tally_status, fmt, st, 44);
  if (rc != GCP_OK) return rc;
  CU(launch_tally_status_merge(nullptr, 0, n_fields, 1, ctx->d_flagPK, (u32*)d_tally, d_tally_status, st), "tally status kernel");
  ctx->launches++;
  return GCP_OK;
}

// Host-buffer form of the ballot batch: voters are streamed in chunks on the two streams (proofs dense, or arbo packed
// when siblings == NULL), each chunk runs verifier -> masked encrypt-tally on the device, the per-chunk partial
// ciphertexts are folded at the end.  out_flags / out_status: per voter; out_tally / out_tally_status: per field.
static int ballot_batch_host(gcp_ctx* ctx, int n_levels, size_t n_voters, const void* roots, int shared_root,
                             const void* siblings, const uint8_t* packed, const uint64_t* offsets, const void* keys,
                             const void* values, const void* pub_key, const void* k, const void* m, int n_fields,
                             uint8_t* out_flags, uint8_t* out_status, void* out_tally, uint8_t* out_tally_status, int fmt,
                             bool tally_on_device) {
  if (!ctx) return GCP_ERR_BAD_ARG;
  std::lock_guard<std::recursive_mutex> lk(ctx->mu);
  DeviceGuard device_guard(ctx->device);
  if (!device_guard.ok) return ctx->fail(GCP_ERR_CUDA, "cudaSetDevice failed");
  StreamGuard guard{ctx};
  const bool is_packed = siblings == nullptr;
  const size_t n = n_voters;
  int rc = encrypt_tally_check(ctx, pub_key, k, m, n, n_fields, out_tally, out_tally_status, fmt);
  if (rc != GCP_OK) return rc;
  if (n_levels < 2 || n_levels > 253) return ctx->fail(GCP_ERR_BAD_ARG, "n_levels must be in [2, 253]");
  if (n && (!roots || !keys || !values || !out_flags || !out_status || (is_packed && (!packed || !offsets))))
    return ctx->fail(GCP_ERR_BAD_ARG, "null buffer");
  rc = ensure_pk_table(ctx, pub_key, false, fmt, ctx->stream[0], (uint64_t)n * n_fields);
  if (rc != GCP_OK) return rc;
  const size_t sib_bytes = (size_t)n_levels * 32, ballot_in = (size_t)n_fields * 32, ballot_ct = (size_t)n_fields * 128;
  const size_t ballot_m = (size_t)n_fields * msg_bytes(fmt);
  ChunkPlan plan(n, smt_path_wave_items(ctx->sm_count), ((size_t)1 << 30) / (sib_bytes + ballot_in + ballot_m), true);
  std::vector<size_t> sizes;
  {
    ChunkPlan walk = plan;
    for (size_t c = walk.next(); c != 0; c = walk.next()) sizes.push_back(c);
    if (sizes.empty()) sizes.push_back(0);
  }
  const size_t n_chunks = sizes.size(), cap = std::max<size_t>(plan.largest(), 1);
  u32* d_parts = (u32*)ctx->buf(64, n_chunks * ballot_ct);
  uint8_t* d_part_status = (uint8_t*)ctx->buf(65, n_chunks * n_fields);
  void* d_shared_root = nullptr;
  if (!d_parts || !d_part_status) return ctx->fail(GCP_ERR_ALLOC, "device allocation failed");
  if (shared_root && n) {
    d_shared_root = ctx->buf(3, 32);
    if (!d_shared_root) return ctx->fail(GCP_ERR_ALLOC, "device allocation failed");
    CU(cudaMemcpy(d_shared_root, roots, 32, cudaMemcpyHostToDevice), "H2D root");
  }
  size_t off = 0;
  for (size_t c = 0; c < n_chunks; off += sizes[c], c++) {
    const size_t cnt = sizes[c];
    const int s = (int)(c & 1);
    cudaStream_t st = ctx->stream[s];
    const int b = 10 + s * 14;
    // every slot is sized for the largest chunk of the plan: growing one later would cudaFree under running work
    void* dk = ctx->buf(48 + s * 8, cap * ballot_in);
    void* dm = ctx->buf(48 + s * 8 + 4, cap * ballot_m);
    uint8_t* d_flags = (uint8_t*)ctx->buf(b + 9, cap);
    if (!dk || !dm || !d_flags) return ctx->fail(GCP_ERR_ALLOC, "device allocation failed");
    if (cnt) {
      void* d_sib = ctx->buf(b + 0, cap * sib_bytes);
      void* d_roots = shared_root ? d_shared_root : ctx->buf(b + 1, cap * 32);
      void* d_keys = ctx->buf(b + 2, cap * 32);
      void* d_vals = ctx->buf(b + 3, cap * 32);
      uint8_t* d_status = (uint8_t*)ctx->buf(b + 10, cap);
      if (!d_sib || !d_roots || !d_keys || !d_vals || !d_status || !ctx->buf(b + 12, smt_scratch_bytes(cap)))
        return ctx->fail(GCP_ERR_ALLOC, "device allocation failed");
      uint8_t* d_bad = nullptr;
      if (is_packed) {
        const uint64_t pbeg = offsets[off], pend = offsets[off + cnt];
        if (pend < pbeg) return ctx->fail(GCP_ERR_BAD_ARG, "packed offsets must be non-decreasing");
        const size_t pbytes = (size_t)(pend - pbeg);
        uint8_t* d_packed = (uint8_t*)ctx->buf(80 + s * 3 + 0, pbytes + 4);
        uint64_t* d_off = (uint64_t*)ctx->buf(80 + s * 3 + 1, (cap + 1) * 8);
        d_bad = (uint8_t*)ctx->buf(80 + s * 3 + 2, cap);
        if (!d_packed || !d_off || !d_bad) return ctx->fail(GCP_ERR_ALLOC, "device allocation failed");
        if (pbytes) GCP_TRY(h2d_copy(ctx, d_packed, packed + pbeg, pbytes, st));
        GCP_TRY(h2d_copy(ctx, d_off, offsets + off, (cnt + 1) * 8, st));
        CU(launch_smt_unpack(d_packed, d_off, pbeg, pbytes, cnt, n_levels, (u32*)d_sib, d_bad, elem_fmt(fmt), st), "smt unpack kernel");
        ctx->launches++;
      } else {
        GCP_TRY(h2d_copy(ctx, d_sib, (const char*)siblings + off * sib_bytes, cnt * sib_bytes, st));
      }
      if (!shared_root) GCP_TRY(h2d_copy(ctx, d_roots, (const char*)roots + off * 32, cnt * 32, st));
      GCP_TRY(h2d_copy(ctx, d_keys, (const char*)keys + off * 32, cnt * 32, st));
      GCP_TRY(h2d_copy(ctx, d_vals, (const char*)values + off * 32, cnt * 32, st));
      GCP_TRY(h2d_copy(ctx, dk, (const char*)k + off * ballot_in, cnt * ballot_in, st));
      GCP_TRY(h2d_copy(ctx, dm, (const char*)m + off * ballot_m, cnt * ballot_m, st));
      rc = smt_verify_dev_locked(ctx, n_levels, cnt, d_roots, shared_root, d_sib, nullptr, nullptr, nullptr, d_keys, d_vals,
                                 nullptr, nullptr, d_flags, d_status, nullptr, elem_fmt(fmt), st, b + 12);
      if (rc != GCP_OK) return rc;
      if (is_packed) {
        CU(launch_smt_apply_bad(d_bad, cnt, d_flags, d_status, nullptr, st), "smt apply-bad kernel");
        ctx->launches++;
      }
      GCP_TRY(d2h_copy(ctx, out_flags + off, d_flags, cnt, st));
      GCP_TRY(d2h_copy(ctx, out_status + off, d_status, cnt, st));
    }
    // flags are 0 wherever status != 0, so the flag array is the admission mask of the fold
    rc = encrypt_tally_dev_locked(ctx, dk, dm, d_flags, cnt, n_fields, (char*)d_parts + c * ballot_ct,
                                  d_part_status + c * n_fields, fmt, st, 48 + s * 8 + 1);
    if (rc != GCP_OK) return rc;
  }
  CU(cudaStreamSynchronize(ctx->stream[0]), "stream sync");
  CU(cudaStreamSynchronize(ctx->stream[1]), "stream sync");
  return finish_partials(ctx, d_parts, d_part_status, n_chunks, n_fields, fmt, ctx->d_flagPK, out_tally, out_tally_status,
                         tally_on_device);
}

int gcp_ballot_batch(gcp_ctx* ctx, int n_levels, size_t n_voters, const void* roots, int shared_root, const void* siblings,
                     const uint8_t* packed, const uint64_t* offsets, const void* keys, const void* values,
                     const void* pub_key, const void* k, const void* m, int n_fields, uint8_t* out_flags,
                     uint8_t* out_status, void* out_tally, uint8_t* out_tally_status, int fmt) {
  return ballot_batch_host(ctx, n_levels, n_voters, roots, shared_root, siblings, packed, offsets, keys, values, pub_key, k,
                           m, n_fields, out_flags, out_status, out_tally, out_tally_status, fmt, false);
}

// Folds that leave their result on the device, for the single-process group (group.cu): the partial tally of a device's
// slice goes straight into the all-gather's send buffer.  Declared in internal.h, not part of the public ABI.
int gcp_internal_tally_to_dev(gcp_ctx* ctx, const void* ct, size_t n_ballots, int n_fields, void* d_out, uint8_t* d_status,
                              int fmt) {
  return tally_host(ctx, ct, n_ballots, n_fields, d_out, d_status, fmt, true);
}
int gcp_internal_merge_status_dev(gcp_ctx* ctx, const uint8_t* d_part_status, int n_parts, int n_fields, void* d_ct,
                                  uint8_t* d_status, void* stream) {
  if (!ctx) return GCP_ERR_BAD_ARG;
  std::lock_guard<std::recursive_mutex> lk(ctx->mu);
  DeviceGuard device_guard(ctx->device);
  if (!device_guard.ok) return ctx->fail(GCP_ERR_CUDA, "cudaSetDevice failed");
  CU(launch_tally_status_merge(d_part_status, n_parts, n_fields, 1, nullptr, (u32*)d_ct, d_status, (cudaStream_t)stream),
     "tally status kernel");
  ctx->launches++;
  return GCP_OK;
}
int gcp_internal_encrypt_tally_to_dev(gcp_ctx* ctx, const void* pub_key, const void* k, const void* m, size_t n_ballots,
                                      int n_fields, void* d_out, uint8_t* d_status, int fmt) {
  return encrypt_tally_host(ctx, pub_key, k, m, n_ballots, n_fields, d_out, d_status, fmt, true);
}
int gcp_internal_ballot_batch_to_dev(gcp_ctx* ctx, int n_levels, size_t n_voters, const void* roots, int shared_root,
                                     const void* siblings, const uint8_t* packed, const uint64_t* offsets, const void* keys,
                                     const void* values, const void* pub_key, const void* k, const void* m, int n_fields,
                                     uint8_t* out_flags, uint8_t* out_status, void* d_tally, uint8_t* d_tally_status, int fmt) {
  return ballot_batch_host(ctx, n_levels, n_voters, roots, shared_root, siblings, packed, offsets, keys, values, pub_key, k,
                           m, n_fields, out_flags, out_status, d_tally, d_tally_status, fmt, true);
}

// ---------------------------------------------------------------------------------------------------
// Decryption checks and coordinate conversion: host-buffer forms
// ---------------------------------------------------------------------------------------------------
namespace {
struct Upload {
  const void* host;
  size_t bytes_per_item;
};
}  // namespace

// uploads `ins` (slot 70.. on stream 0), returns device pointers
static int upload_all(gcp_ctx* ctx, const Upload* ins, int n_ins, size_t n, void** d_ptrs) {
  for (int i = 0; i < n_ins; i++) {
    d_ptrs[i] = ctx->buf(70 + i, n * ins[i].bytes_per_item);
    if (!d_ptrs[i]) return ctx->fail(GCP_ERR_ALLOC, "device allocation failed");
    GCP_TRY(h2d_copy(ctx, d_ptrs[i], ins[i].host, n * ins[i].bytes_per_item, ctx->stream[0]));
  }
  return GCP_OK;
}

// Flag-producing per-item kernels from host buffers: chunks of 2^18 items alternate between the two streams, so the
// copies of one chunk overlap the kernel of the other (large pageable sources go through h2d_copy's staging ring).
extern "C++" {
template <typename Launch>
static int per_item_pipeline(gcp_ctx* ctx, const Upload* ins, int n_ins, size_t n, uint8_t* out_flags, uint8_t* status,
                             const char* what, Launch launch) {
  // chunks of whole resident waves of the window kernel (the first one a single wave: its copy is the only exposed one)
  ChunkPlan plan(n, varbase_wave_items(ctx->sm_count), (size_t)1 << 18);
  const size_t cap = plan.largest();
  size_t k = 0, off = 0;
  for (size_t m = plan.next(); m != 0; off += m, m = plan.next(), k++) {
    const int s = (int)(k & 1);
    cudaStream_t st = ctx->stream[s];
    const int base = s ? 86 : 70;
    void* d[8];
    for (int i = 0; i < n_ins; i++) {
      d[i] = ctx->buf(base + i, cap * ins[i].bytes_per_item);
      if (!d[i]) return ctx->fail(GCP_ERR_ALLOC, "device allocation failed");
      GCP_TRY(h2d_copy(ctx, d[i], (const char*)ins[i].host + off * ins[i].bytes_per_item, m * ins[i].bytes_per_item, st));
    }
    uint8_t* d_flags = (uint8_t*)ctx->buf(s ? 92 : 77, cap);
    uint8_t* d_status = (uint8_t*)ctx->buf(s ? 93 : 78, cap);
    if (!d_flags || !d_status) return ctx->fail(GCP_ERR_ALLOC, "device allocation failed");
    CU(launch(d, m, cap, d_flags, d_status, s, st), what);
    GCP_TRY(d2h_copy(ctx, out_flags + off, d_flags, m, st));
    GCP_TRY(d2h_copy(ctx, status + off, d_status, m, st));
  }
  CU(cudaStreamSynchronize(ctx->stream[0]), "stream sync");
  CU(cudaStreamSynchronize(ctx->stream[1]), "stream sync");
  return GCP_OK;
}
}  // extern "C++"

// curve.ScalarMul (gnark twistededwards; call sites elgamal/encrypt.go:55, ciphertext.go:58,147-160): out[i] = [s[i]]P[i],
// or with a second base [s[i]]P[i] + [s2[i]]P2[i] in one pass that shares the doublings.
static int scalar_mul_dev_locked(gcp_ctx* ctx, const void* d_points, const void* d_scalars, const void* d_points2,
                                 const void* d_scalars2, size_t n, void* d_out, uint8_t* d_status, int fmt, cudaStream_t st,
                                 int slot) {
  int rc = check_fmt_points(ctx, fmt);
  if (rc != GCP_OK || n == 0) return rc;
  if (!d_points || !d_scalars || !d_out || !d_status || ((d_points2 == nullptr) != (d_scalars2 == nullptr)))
    return ctx->fail(GCP_ERR_BAD_ARG, "null buffer");
  const int nb = d_points2 ? 2 : 1;
  u32* scratch = (u32*)ctx->buf(slot, scalar_mul_scratch_bytes(n, nb) + n * 128);
  if (!scratch) return ctx->fail(GCP_ERR_ALLOC, "device allocation failed");
  u32* ext = scratch + scalar_mul_scratch_bytes(n, nb) / 4;
  int nl = 0;
  CU(launch_scalar_mul((const u32*)d_points, (const u32*)d_scalars, (const u32*)d_points2, (const u32*)d_scalars2, n, ext,
                       d_status, elem_fmt(fmt), scratch, &nl, st, coords_te(fmt)),
     "scalar-mul kernels");
  CU(launch_normalize(ext, n, (u32*)d_out, d_status, 1, elem_fmt(fmt), st, 32, coords_te(fmt)), "normalize kernel");
  ctx->launches += nl + 1;
  return GCP_OK;
}

int gcp_elgamal_scalar_mul_dev(gcp_ctx* ctx, const void* d_points, const void* d_scalars, const void* d_points2,
                               const void* d_scalars2, size_t n, void* d_out_points, uint8_t* d_status, int fmt,
                               void* stream) {
  if (!ctx) return GCP_ERR_BAD_ARG;
  std::lock_guard<std::recursive_mutex> lk(ctx->mu);
  DeviceGuard device_guard(ctx->device);
  if (!device_guard.ok) return ctx->fail(GCP_ERR_CUDA, "cudaSetDevice failed");
  return scalar_mul_dev_locked(ctx, d_points, d_scalars, d_points2, d_scalars2, n, d_out_points, d_status, fmt,
                               (cudaStream_t)stream, 98);
}

int gcp_elgamal_scalar_mul(gcp_ctx* ctx, const void* points, const void* scalars, const void* points2, const void* scalars2,
                           size_t n, void* out_points, uint8_t* status, int fmt) {
  if (!ctx) return GCP_ERR_BAD_ARG;
  std::lock_guard<std::recursive_mutex> lk(ctx->mu);
  DeviceGuard device_guard(ctx->device);
  if (!device_guard.ok) return ctx->fail(GCP_ERR_CUDA, "cudaSetDevice failed");
  StreamGuard guard{ctx};
  int rc = check_fmt_points(ctx, fmt);
  if (rc != GCP_OK || n == 0) return rc;
  if (!points || !scalars || !out_points || !status || ((points2 == nullptr) != (scalars2 == nullptr)))
    return ctx->fail(GCP_ERR_BAD_ARG, "null buffer");
  const int n_ins = points2 ? 4 : 2;
  const Upload ins[4] = {{points, 64}, {scalars, 32}, {points2, 64}, {scalars2, 32}};
  const size_t chunk = std::min<size_t>(n, (size_t)1 << 18);
  size_t k = 0;
  for (size_t off = 0; off < n; off += chunk, k++) {
    const size_t m = std::min(chunk, n - off);
    const int s = (int)(k & 1);
    cudaStream_t st = ctx->stream[s];
    void* d[4] = {nullptr, nullptr, nullptr, nullptr};
    for (int i = 0; i < n_ins; i++) {
      d[i] = ctx->buf((s ? 86 : 70) + i, m * ins[i].bytes_per_item);
      if (!d[i]) return ctx->fail(GCP_ERR_ALLOC, "device allocation failed");
      GCP_TRY(h2d_copy(ctx, d[i], (const char*)ins[i].host + off * ins[i].bytes_per_item, m * ins[i].bytes_per_item, st));
    }
    void* d_out = ctx->buf(s ? 91 : 76, m * 64);
    uint8_t* d_status = (uint8_t*)ctx->buf(s ? 93 : 78, m);
    if (!d_out || !d_status) return ctx->fail(GCP_ERR_ALLOC, "device allocation failed");
    GCP_TRY(scalar_mul_dev_locked(ctx, d[0], d[1], d[2], d[3], m, d_out, d_status, fmt, st, 96 + s));
    GCP_TRY(d2h_copy(ctx, (char*)out_points + off * 64, d_out, m * 64, st));
    GCP_TRY(d2h_copy(ctx, status + off, d_status, m, st));
  }
  CU(cudaStreamSynchronize(ctx->stream[0]), "stream sync");
  CU(cudaStreamSynchronize(ctx->stream[1]), "stream sync");
  return GCP_OK;
}

int gcp_elgamal_assert_decrypt(gcp_ctx* ctx, const void* ct, const void* priv_keys, const void* msgs, size_t n,
                               uint8_t* out_flags, uint8_t* status, int fmt) {
  if (!ctx) return GCP_ERR_BAD_ARG;
  std::lock_guard<std::recursive_mutex> lk(ctx->mu);
  DeviceGuard device_guard(ctx->device);
  if (!device_guard.ok) return ctx->fail(GCP_ERR_CUDA, "cudaSetDevice failed");
  StreamGuard guard{ctx};
  int rc = check_fmt_points(ctx, fmt);
  if (rc != GCP_OK || n == 0) return rc;
  if (!ct || !priv_keys || !msgs || !out_flags || !status) return ctx->fail(GCP_ERR_BAD_ARG, "null buffer");
  Upload ins[3] = {{ct, 128}, {priv_keys, 32}, {msgs, 32}};
  return per_item_pipeline(ctx, ins, 3, n, out_flags, status, "assert-decrypt kernels",
                           [&](void** d, size_t m, size_t cap, uint8_t* d_flags, uint8_t* d_status, int s, cudaStream_t st) {
                             u32* scratch = (u32*)ctx->buf(96 + s, varbase_scratch_bytes(1, cap));
                             if (!scratch) return cudaErrorMemoryAllocation;
                             int nl = 0;
                             cudaError_t e = launch_assert_decrypt(ctx->d_tabG, (const u32*)d[0], (const u32*)d[1],
                                                                   (const u32*)d[2], m, d_flags, d_status, elem_fmt(fmt), scratch, &nl,
                                                                   st, coords_te(fmt));
                             ctx->launches += nl;
                             return e;
                           });
}

int gcp_elgamal_verify_decryption_proof(gcp_ctx* ctx, const void* pub_keys, const void* ct, const void* msgs,
                                        const void* a1, const void* a2, const void* z, size_t n, uint8_t* out_flags,
                                        uint8_t* status, int fmt) {
  if (!ctx) return GCP_ERR_BAD_ARG;
  std::lock_guard<std::recursive_mutex> lk(ctx->mu);
  DeviceGuard device_guard(ctx->device);
  if (!device_guard.ok) return ctx->fail(GCP_ERR_CUDA, "cudaSetDevice failed");
  StreamGuard guard{ctx};
  int rc = check_fmt_points(ctx, fmt);
  if (rc != GCP_OK || n == 0) return rc;
  if (!pub_keys || !ct || !msgs || !a1 || !a2 || !z || !out_flags || !status) return ctx->fail(GCP_ERR_BAD_ARG, "null buffer");
  Upload ins[6] = {{pub_keys, 64}, {ct, 128}, {msgs, 32}, {a1, 64}, {a2, 64}, {z, 32}};
  return per_item_pipeline(ctx, ins, 6, n, out_flags, status, "decryption-proof kernels",
                           [&](void** d, size_t m, size_t cap, uint8_t* d_flags, uint8_t* d_status, int s, cudaStream_t st) {
                             u32* scratch = (u32*)ctx->buf(96 + s, varbase_scratch_bytes(2, cap));
                             if (!scratch) return cudaErrorMemoryAllocation;
                             int nl = 0;
                             cudaError_t e = launch_decryption_proof(ctx->d_tabG, ctx->tab[13], (const u32*)d[0], (const u32*)d[1],
                                                                     (const u32*)d[2], (const u32*)d[3], (const u32*)d[4],
                                                                     (const u32*)d[5], m, d_flags, d_status, elem_fmt(fmt), scratch,
                                                                     &nl, st, coords_te(fmt));
                             ctx->launches += nl;
                             return e;
                           });
}

int gcp_eddsa_verify(gcp_ctx* ctx, const void* pub_keys_te, const void* sig_r_te, const void* sig_s, const void* msgs,
                     size_t n, uint8_t* out_flags, uint8_t* status, int fmt) {
  if (!ctx) return GCP_ERR_BAD_ARG;
  std::lock_guard<std::recursive_mutex> lk(ctx->mu);
  DeviceGuard device_guard(ctx->device);
  if (!device_guard.ok) return ctx->fail(GCP_ERR_CUDA, "cudaSetDevice failed");
  StreamGuard guard{ctx};
  int rc = check_fmt(ctx, fmt);
  if (rc != GCP_OK || n == 0) return rc;
  if (!pub_keys_te || !sig_r_te || !sig_s || !msgs || !out_flags || !status) return ctx->fail(GCP_ERR_BAD_ARG, "null buffer");
  Upload ins[4] = {{pub_keys_te, 64}, {sig_r_te, 64}, {sig_s, 32}, {msgs, 32}};
  return per_item_pipeline(ctx, ins, 4, n, out_flags, status, "eddsa kernels",
                           [&](void** d, size_t m, size_t cap, uint8_t* d_flags, uint8_t* d_status, int s, cudaStream_t st) {
                             u32* scratch = (u32*)ctx->buf(96 + s, varbase_scratch_bytes(3, cap));
                             if (!scratch) return cudaErrorMemoryAllocation;
                             int nl = 0;
                             cudaError_t e = launch_eddsa_verify(ctx->d_tabG, ctx->tab[6], (const u32*)d[0], (const u32*)d[1],
                                                                 (const u32*)d[2], (const u32*)d[3], m, d_flags, d_status, fmt,
                                                                 scratch, &nl, st);
                             ctx->launches += nl;
                             return e;
                           });
}

static int te_rte_host(gcp_ctx* ctx, const void* in, size_t n_points, void* out, uint8_t* status, int to_rte) {
  if (!ctx) return GCP_ERR_BAD_ARG;
  std::lock_guard<std::recursive_mutex> lk(ctx->mu);
  DeviceGuard device_guard(ctx->device);
  if (!device_guard.ok) return ctx->fail(GCP_ERR_CUDA, "cudaSetDevice failed");
  StreamGuard guard{ctx};
  if (n_points == 0) return GCP_OK;
  if (!in || !out || !status) return ctx->fail(GCP_ERR_BAD_ARG, "null buffer");
  Upload ins[1] = {{in, 64}};
  void* d[1];
  int rc = upload_all(ctx, ins, 1, n_points, d);
  if (rc != GCP_OK) return rc;
  void* d_out = ctx->buf(76, n_points * 64);
  uint8_t* d_status = (uint8_t*)ctx->buf(78, n_points);
  if (!d_out || !d_status) return ctx->fail(GCP_ERR_ALLOC, "device allocation failed");
  cudaStream_t st = ctx->stream[0];
  CU(launch_te_rte((const u32*)d[0], n_points, (u32*)d_out, d_status, to_rte, st), "te/rte kernel");
  ctx->launches++;
  GCP_TRY(d2h_copy(ctx, out, d_out, n_points * 64, st));
  GCP_TRY(d2h_copy(ctx, status, d_status, n_points, st));
  CU(cudaStreamSynchronize(st), "stream sync");
  return GCP_OK;
}

int gcp_te_to_rte(gcp_ctx* ctx, const void* points, size_t n_points, void* out, uint8_t* status) {
  return te_rte_host(ctx, points, n_points, out, status, 1);
}
int gcp_rte_to_te(gcp_ctx* ctx, const void* points, size_t n_points, void* out, uint8_t* status) {
  return te_rte_host(ctx, points, n_points, out, status, 0);
}

// ---------------------------------------------------------------------------------------------------
// MiMC7 (hash/native/bn254/mimc7)
// ---------------------------------------------------------------------------------------------------
int gcp_mimc7_hash_dev(gcp_ctx* ctx, const void* d_in, int len, size_t n, void* d_out, uint8_t* d_status, int fmt,
                       void* stream) {
  if (!ctx) return GCP_ERR_BAD_ARG;
  std::lock_guard<std::recursive_mutex> lk(ctx->mu);
  DeviceGuard device_guard(ctx->device);
  if (!device_guard.ok) return ctx->fail(GCP_ERR_CUDA, "cudaSetDevice failed");
  int rc = check_fmt(ctx, fmt);
  if (rc != GCP_OK) return rc;
  if (!ctx->have_mimc7) return ctx->fail(GCP_ERR_CONSTANTS, "MiMC7 constants (data/mimc7_bn254.bin) were not found");
  if (len < 1 || len > 62) return ctx->fail(GCP_ERR_BAD_ARG, "MiMC7 takes 1..62 inputs");  // mimc.go:9,33-38
  if (n == 0) return GCP_OK;
  if (!d_in || !d_out || !d_status) return ctx->fail(GCP_ERR_BAD_ARG, "null buffer");
  CU(launch_mimc7((const u32*)d_in, len, n, (u32*)d_out, d_status, fmt, (cudaStream_t)stream), "mimc7 kernel");
  ctx->launches++;
  return GCP_OK;
}

int gcp_mimc7_hash(gcp_ctx* ctx, const void* in, int len, size_t n, void* out, uint8_t* status, int fmt) {
  if (!ctx) return GCP_ERR_BAD_ARG;
  std::lock_guard<std::recursive_mutex> lk(ctx->mu);  // held over upload, kernel and read-back: the slots are this call's
  DeviceGuard device_guard(ctx->device);
  if (!device_guard.ok) return ctx->fail(GCP_ERR_CUDA, "cudaSetDevice failed");
  StreamGuard guard{ctx};
  if (len < 1 || len > 62) return ctx->fail(GCP_ERR_BAD_ARG, "MiMC7 takes 1..62 inputs");
  if (n == 0) return GCP_OK;
  if (!in || !out || !status) return ctx->fail(GCP_ERR_BAD_ARG, "null buffer");
  void* d[1];
  Upload ins[1] = {{in, (size_t)len * 32}};
  int rc = upload_all(ctx, ins, 1, n, d);
  if (rc != GCP_OK) return rc;
  void* d_out = ctx->buf(76, n * 32);
  uint8_t* d_status = (uint8_t*)ctx->buf(78, n);
  if (!d_out || !d_status) return ctx->fail(GCP_ERR_ALLOC, "device allocation failed");
  rc = gcp_mimc7_hash_dev(ctx, d[0], len, n, d_out, d_status, fmt, ctx->stream[0]);
  if (rc != GCP_OK) return rc;
  GCP_TRY(d2h_copy(ctx, out, d_out, n * 32, ctx->stream[0]));
  GCP_TRY(d2h_copy(ctx, status, d_status, n, ctx->stream[0]));
  CU(cudaStreamSynchronize(ctx->stream[0]), "stream sync");
  return GCP_OK;
}

// ---------------------------------------------------------------------------------------------------
// Poseidon2, width 2 (hash/native/bn254/poseidon2)
// ---------------------------------------------------------------------------------------------------
int gcp_poseidon2_set_round_keys(gcp_ctx* ctx, const void* keys, size_t n_keys, int fmt) {
  if (!ctx) return GCP_ERR_BAD_ARG;
  std::lock_guard<std::recursive_mutex> lk(ctx->mu);
  DeviceGuard device_guard(ctx->device);
  if (!device_guard.ok) return ctx->fail(GCP_ERR_CUDA, "cudaSetDevice failed");
  int rc = check_fmt(ctx, fmt);
  if (rc != GCP_OK) return rc;
  if (!keys) return ctx->fail(GCP_ERR_BAD_ARG, "null buffer");
  if (n_keys != 62) return ctx->fail(GCP_ERR_BAD_ARG, "Poseidon2 (t=2, rF=6, rP=50) takes 62 round keys");
  // every key must be < r in either format (fr.Element's invariant / SetBytes reduces)
  static const uint64_t P[4] = {0x43e1f593f0000001ull, 0x2833e84879b97091ull, 0xb85045b68181585dull, 0x30644e72e131a029ull};
  for (size_t i = 0; i < n_keys; i++) {
    uint64_t w[4];
    memcpy(w, (const char*)keys + i * 32, 32);
    bool lt = false;
    for (int l = 3; l >= 0; l--) {
      if (w[l] != P[l]) {
        lt = w[l] < P[l];
        break;
      }
    }
    if (!lt) return ctx->fail(GCP_ERR_BAD_ARG, "Poseidon2 round key >= r");
  }
  cudaStream_t st = ctx->stream[0];
  CU(cudaStreamSynchronize(ctx->stream[1]), "stream sync");
  GCP_TRY(h2d_copy(ctx, ctx->d_p2_keys, keys, 62 * 32, st));
  if (fmt == GCP_FMT_CANONICAL) {
    CU(launch_to_mont(ctx->d_p2_keys, 62, st), "to_mont kernel");
    ctx->launches++;
  }
  CU(cudaStreamSynchronize(st), "stream sync");
  ctx->have_p2_keys = true;
  return GCP_OK;
}

static int p2_check(gcp_ctx* ctx, int fmt) {
  int rc = check_fmt(ctx, fmt);
  if (rc != GCP_OK) return rc;
  if (!ctx->have_p2_keys)
    return ctx->fail(GCP_ERR_CONSTANTS,
                     "Poseidon2 round keys missing (data/poseidon2_bn254_t2.bin not found and gcp_poseidon2_set_round_keys not called)");
  return GCP_OK;
}

int gcp_poseidon2_hash_dev(gcp_ctx* ctx, const void* d_in, int len, size_t n, void* d_out, uint8_t* d_status, int fmt,
                           void* stream) {
  if (!ctx) return GCP_ERR_BAD_ARG;
  std::lock_guard<std::recursive_mutex> lk(ctx->mu);
  DeviceGuard device_guard(ctx->device);
  if (!device_guard.ok) return ctx->fail(GCP_ERR_CUDA, "cudaSetDevice failed");
  int rc = p2_check(ctx, fmt);
  if (rc != GCP_OK) return rc;
  if (len != 2 && len != 3) return ctx->fail(GCP_ERR_BAD_ARG, "poseidon2: need 2 or 3 limbs");  // native.go:31-33
  if (n == 0) return GCP_OK;
  if (!d_in || !d_out || !d_status) return ctx->fail(GCP_ERR_BAD_ARG, "null buffer");
  CU(launch_poseidon2_hash(ctx->d_p2_keys, (const u32*)d_in, len, n, (u32*)d_out, d_status, fmt, (cudaStream_t)stream),
     "poseidon2 hash kernel");
  ctx->launches++;
  return GCP_OK;
}

int gcp_poseidon2_permutation_dev(gcp_ctx* ctx, const void* d_in, size_t n, void* d_out, uint8_t* d_status, int fmt,
                                  void* stream) {
  if (!ctx) return GCP_ERR_BAD_ARG;
  std::lock_guard<std::recursive_mutex> lk(ctx->mu);
  DeviceGuard device_guard(ctx->device);
  if (!device_guard.ok) return ctx->fail(GCP_ERR_CUDA, "cudaSetDevice failed");
  int rc = p2_check(ctx, fmt);
  if (rc != GCP_OK) return rc;
  if (n == 0) return GCP_OK;
  if (!d_in || !d_out || !d_status) return ctx->fail(GCP_ERR_BAD_ARG, "null buffer");
  CU(launch_poseidon2_permutation(ctx->d_p2_keys, (const u32*)d_in, n, (u32*)d_out, d_status, fmt, (cudaStream_t)stream),
     "poseidon2 permutation kernel");
  ctx->launches++;
  return GCP_OK;
}

// host buffers; out_elems = elements written per item (1 for the hash, 2 for the permutation), len = 0 selects the permutation
static int p2_host(gcp_ctx* ctx, const void* in, int len, size_t n, void* out, uint8_t* status, int fmt) {
  if (!ctx) return GCP_ERR_BAD_ARG;
  const bool perm = len == 0;
  const size_t in_bytes = perm ? 64 : (size_t)len * 32, out_bytes = perm ? 64 : 32;
  std::lock_guard<std::recursive_mutex> lk(ctx->mu);  // held for the whole call: the scratch slots are this call's
  DeviceGuard device_guard(ctx->device);
  if (!device_guard.ok) return ctx->fail(GCP_ERR_CUDA, "cudaSetDevice failed");
  StreamGuard guard{ctx};
  {
    int rc = p2_check(ctx, fmt);
    if (rc != GCP_OK) return rc;
    if (!perm && len != 2 && len != 3) return ctx->fail(GCP_ERR_BAD_ARG, "poseidon2: need 2 or 3 limbs");
    if (n == 0) return GCP_OK;
    if (!in || !out || !status) return ctx->fail(GCP_ERR_BAD_ARG, "null buffer");
  }
  const size_t chunk = (size_t)1 << 20;
  size_t k = 0;
  for (size_t off = 0; off < n; off += chunk, k++) {
    const size_t m = std::min(chunk, n - off);
    const int s = (int)(k & 1);
    void *d_in, *d_out;
    uint8_t* d_status;
    {
      d_in = ctx->buf(s ? 86 : 70, m * in_bytes);
      d_out = ctx->buf(s ? 91 : 76, m * out_bytes);
      d_status = (uint8_t*)ctx->buf(s ? 93 : 78, m);
      if (!d_in || !d_out || !d_status) return ctx->fail(GCP_ERR_ALLOC, "device allocation failed");
      GCP_TRY(h2d_copy(ctx, d_in, (const char*)in + off * in_bytes, m * in_bytes, ctx->stream[s]));
    }
    int rc = perm ? gcp_poseidon2_permutation_dev(ctx, d_in, m, d_out, d_status, fmt, ctx->stream[s])
                  : gcp_poseidon2_hash_dev(ctx, d_in, len, m, d_out, d_status, fmt, ctx->stream[s]);
    if (rc != GCP_OK) return rc;
    GCP_TRY(d2h_copy(ctx, (char*)out + off * out_bytes, d_out, m * out_bytes, ctx->stream[s]));
    GCP_TRY(d2h_copy(ctx, status + off, d_status, m, ctx->stream[s]));
  }
  CU(cudaStreamSynchronize(ctx->stream[0]), "stream sync");
  CU(cudaStreamSynchronize(ctx->stream[1]), "stream sync");
  return GCP_OK;
}

int gcp_poseidon2_hash(gcp_ctx* ctx, const void* in, int len, size_t n, void* out, uint8_t* status, int fmt) {
  if (ctx && len == 0) return ctx->fail(GCP_ERR_BAD_ARG, "poseidon2: need 2 or 3 limbs");
  return p2_host(ctx, in, len, n, out, status, fmt);
}

int gcp_poseidon2_permutation(gcp_ctx* ctx, const void* in, size_t n, void* out, uint8_t* status, int fmt) {
  return p2_host(ctx, in, 0, n, out, status, fmt);
}

// ---------------------------------------------------------------------------------------------------
// Keccak address derivation
// ---------------------------------------------------------------------------------------------------
int gcp_keccak_address_dev(gcp_ctx* ctx, const void* d_pub_xy_be, size_t n, void* d_out_addr, void* stream) {
  if (!ctx) return GCP_ERR_BAD_ARG;
  std::lock_guard<std::recursive_mutex> lk(ctx->mu);
  DeviceGuard device_guard(ctx->device);
  if (!device_guard.ok) return ctx->fail(GCP_ERR_CUDA, "cudaSetDevice failed");
  if (n == 0) return GCP_OK;
  if (!d_pub_xy_be || !d_out_addr) return ctx->fail(GCP_ERR_BAD_ARG, "null buffer");
  CU(launch_keccak_address((const u8*)d_pub_xy_be, n, (u8*)d_out_addr, (cudaStream_t)stream), "keccak kernel");
  ctx->launches++;
  return GCP_OK;
}

int gcp_keccak_address(gcp_ctx* ctx, const void* pub_xy_be, size_t n, void* out_addr) {
  if (!ctx) return GCP_ERR_BAD_ARG;
  std::lock_guard<std::recursive_mutex> lk(ctx->mu);
  DeviceGuard device_guard(ctx->device);
  if (!device_guard.ok) return ctx->fail(GCP_ERR_CUDA, "cudaSetDevice failed");
  StreamGuard guard{ctx};
  if (n == 0) return GCP_OK;
  if (!pub_xy_be || !out_addr) return ctx->fail(GCP_ERR_BAD_ARG, "null buffer");
  const size_t chunk = (size_t)1 << 22;
  size_t k = 0;
  for (size_t off = 0; off < n; off += chunk, k++) {
    size_t m = std::min(chunk, n - off);
    int s = (int)(k & 1);
    cudaStream_t st = ctx->stream[s];
    void* d_in = ctx->buf(68 + s * 2, m * 64);
    void* d_out = ctx->buf(69 + s * 2, m * 20);
    if (!d_in || !d_out) return ctx->fail(GCP_ERR_ALLOC, "device allocation failed");
    GCP_TRY(h2d_copy(ctx, d_in, (const char*)pub_xy_be + off * 64, m * 64, st));
    CU(launch_keccak_address((const u8*)d_in, m, (u8*)d_out, st), "keccak kernel");
    ctx->launches++;
    GCP_TRY(d2h_copy(ctx, (char*)out_addr + off * 20, d_out, m * 20, st));
  }
  CU(cudaStreamSynchronize(ctx->stream[0]), "stream sync");
  CU(cudaStreamSynchronize(ctx->stream[1]), "stream sync");
  return GCP_OK;
}

}  // extern "C"

