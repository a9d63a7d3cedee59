// kmg_api.cu — host side of libkmeans_gpu.so: context, per-call workspaces, k-means job
// orchestration and the extern "C" entry points declared in include/kmeans_gpu.h.
//
// There is no CPU compute path in this file: every operation that the reference runs on the GPU
// (core/shaders/*.wgsl) is a kernel launch here.  The only host arithmetic is what the reference
// also does on the host (the `palette` crate conversions, kmg_host.cpp).
#include <algorithm>
#include <cmath>
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include <dlfcn.h>

#include "../../include/kmeans_gpu.h"
#include "kmg_kernels.cuh"
#include "kmg_small.cuh"
#include "kmg_lloyd_ring.cuh"
#include "kmg_audit.cuh"
#include "kmg_init_lazy.cuh"

#if __has_include(<nccl.h>)
#include <nccl.h>
#define KMG_HAVE_NCCL_HEADER 1
#else
#define KMG_HAVE_NCCL_HEADER 0
#endif

using namespace kmg;

// Lloyd-pass variants: <table length (0 = runtime), accumulator capacity, threads, pixels/thread,
// thread-private accumulators, min blocks/SM>
// Geometry variants of the k <= 8 pass (KMG_LLOYD8_VARIANT selects one; tools/sweep_lloyd8.py times them)
// Variants of the thread-private Lloyd pass (k <= 32).  The first entry of a class is its default;
// KMG_LLOYD8_VARIANT / KMG_LLOYD16_VARIANT / KMG_LLOYD32_VARIANT pick another one
// (tools/sweep_lloyd.py times them all and checks that they produce identical sums).
typedef void (*lloyd_fn)(JobPtrs, const float4*, unsigned long long, int, int, PeerXchg, int, unsigned int);
struct LloydVariant {
  lloyd_fn fn;
  int kcap, threads, px, const_tab;
  size_t smem;
  const char* name;
};
static constexpr size_t LLOYD32_SMEM = (32 / 8) * CHUNK_BYTES + 32 * 128 * 16;
#define LV8(P, MINB, CT, ATOM, T) k_lloyd<8, 8, T, P, true, MINB, CT, ATOM>, 8, T, P, CT, (size_t)8 * T * 16
#define LV16(P, MINB, CT, ATOM) k_lloyd<16, 16, 256, P, true, MINB, CT, ATOM>, 16, 256, P, CT, (size_t)16 * 256 * 16
// warp-specialised TMA ring (kmg_lloyd_ring.cuh): 8 consumer warps + 1 producer warp, P px per lane and stage, D stages
#define LVR(P, D, MINB) k_lloyd_ring<8, 8, P, D, MINB>, 8, 288, P, true, (size_t)RingLayout<8, 8, P, D>::BYTES
#define LVRF(P, D, MINB, F) k_lloyd_ring<8, 8, P, D, MINB, F>, 8, 288, P, true, (size_t)RingLayout<8, 8, P, D>::BYTES
static const LloydVariant LLOYD_VARIANTS[] = {
    // k <= 8
    {LVR(4, 4, 2), "TMA ring 4 x 16 KiB, table resident in uniform registers, 8+1 warps x 4 px, 2 blocks/SM"},
    {LV8(4, 2, true, true, 256), "const table, atomic slots, 256 thr x 4 px, 2 blocks/SM"},
    {LV8(4, 2, false, true, 256), "smem table, atomic slots, 256 thr x 4 px, 2 blocks/SM"},
    {LV8(4, 2, false, false, 256), "smem table, 128-bit RMW slots, 256 thr x 4 px, 2 blocks/SM"},
    {LV8(4, 3, true, true, 256), "const table, atomic slots, 256 thr x 4 px, 3 blocks/SM"},
    {LV8(4, 2, true, false, 256), "const table, RMW slots, 256 thr x 4 px, 2 blocks/SM"},
    {LV8(2, 3, false, false, 256), "smem table, RMW slots, 256 thr x 2 px, 3 blocks/SM"},
    {LVR(2, 4, 2), "TMA ring 4 x 8 KiB, 2 blocks/SM"},
    {LVR(2, 8, 2), "TMA ring 8 x 8 KiB, 2 px per lane and stage, 2 blocks/SM"},
    {LVR(2, 4, 3), "TMA ring 4 x 8 KiB, 3 blocks/SM"},
    {LVR(8, 2, 2), "TMA ring 2 x 32 KiB, 2 blocks/SM"},
    {LVR(4, 2, 3), "TMA ring 2 x 16 KiB, 3 blocks/SM"},
    {LVR(4, 2, 2), "TMA ring 2 x 16 KiB, 2 blocks/SM"},
    {LVRF(2, 8, 2, 8), "TMA ring 8 x 8 KiB, memory side alone (timing experiment, sums dropped)"},
    // k <= 16
    {LV16(2, 2, true, true), "const table, atomic slots, 256 thr x 2 px, 2 blocks/SM"},
    {LV16(2, 2, false, true), "smem table, atomic slots, 256 thr x 2 px, 2 blocks/SM"},
    {LV16(2, 2, false, false), "smem table, RMW slots, 256 thr x 2 px, 2 blocks/SM"},
    {LV16(2, 2, true, false), "const table, RMW slots, 256 thr x 2 px, 2 blocks/SM"},
    // k <= 32 (chunked search, table in dynamic shared memory)
    {k_lloyd<0, 32, 128, 4, true, 3, false, true>, 32, 128, 4, 0, LLOYD32_SMEM, "atomic slots, 128 thr x 4 px, 3 blocks/SM"},
    {k_lloyd<0, 32, 128, 4, true, 3, false, false>, 32, 128, 4, 0, LLOYD32_SMEM, "RMW slots, 128 thr x 4 px, 3 blocks/SM"},
};
static constexpr int N_LLOYD_VARIANTS = (int)(sizeof(LLOYD_VARIANTS) / sizeof(LLOYD_VARIANTS[0]));
// index in LLOYD_VARIANTS of variant v of class kcap (-1: none)
static int lloyd_variant_index(int kcap, int v) {
  for (int i = 0; i < N_LLOYD_VARIANTS; ++i)
    if (LLOYD_VARIANTS[i].kcap == kcap && v-- == 0) return i;
  return -1;
}
#define LLOYDG k_lloyd<0, 0, 256, 4, false, 2>
#define LLOYDGS k_lloyd<0, 0, 256, 4, false, 2, false, true>  // block accumulators in shared memory
#define LLOYDGB k_lloyd<0, 0, 256, 4, false, 2, true, true>   // same, multipliers of the chunk loop from c_big (uniform registers)
static constexpr uint32_t LLOYDGS_MAX_K = 2048;  // table 52 KiB + accumulators 56 KiB per block
// Whole-k-means-in-one-launch variants (kmg_small.cuh): <table/accumulator capacity, threads>
#define SMALL8 k_kmeans_small<8, 512>
#define SMALL16 k_kmeans_small<16, 512>
#define SMALL32 k_kmeans_small<32, 256>

// ------------------------------------------------------------------------------------------------
// errors

static thread_local std::string g_last_error = "";

static int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}

#define CU(expr)                                                                                  \
  do {                                                                                            \
    cudaError_t e__ = (expr);                                                                     \
    if (e__ != cudaSuccess)                                                                       \
      return fail(e__ == cudaErrorMemoryAllocation ? KMG_ERR_OOM : KMG_ERR_CUDA, "%s failed: %s", \
                  #expr, cudaGetErrorString(e__));                                                \
  } while (0)
#define TRY(expr)            \
  do {                       \
    int r__ = (expr);        \
    if (r__ != KMG_OK) return r__; \
  } while (0)

// ------------------------------------------------------------------------------------------------
// NCCL through dlopen, so single-GPU users need no NCCL at link or load time.

#if KMG_HAVE_NCCL_HEADER
struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
};
static NcclApi& nccl_api() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    // If the process already loaded an NCCL (e.g. torch's bundled one) the SONAME lookup reuses it.
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
      api.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
      if (api.handle) break;
    }
    if (!api.handle) return;
    api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(api.handle, "ncclGetUniqueId");
    api.CommInitRank = (decltype(api.CommInitRank))dlsym(api.handle, "ncclCommInitRank");
    api.CommDestroy = (decltype(api.CommDestroy))dlsym(api.handle, "ncclCommDestroy");
    api.AllReduce = (decltype(api.AllReduce))dlsym(api.handle, "ncclAllReduce");
    api.GetErrorString = (decltype(api.GetErrorString))dlsym(api.handle, "ncclGetErrorString");
    api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllReduce && api.GetErrorString;
  });
  return api;
}
#define NC(expr)                                                                                     \
  do {                                                                                               \
    ncclResult_t r__ = (expr);                                                                       \
    if (r__ != ncclSuccess) return fail(KMG_ERR_NCCL, "%s failed: %s", #expr, nccl_api().GetErrorString(r__)); \
  } while (0)
#endif

// ------------------------------------------------------------------------------------------------
// context / workspace

struct Buf {
  void* p = nullptr;
  size_t cap = 0;
  int ensure(size_t bytes) {
    if (bytes <= cap) return KMG_OK;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) {
      cudaGetLastError();
      e = cudaMalloc(&p, bytes);
      want = bytes;
    }
    if (e != cudaSuccess) {
      cudaGetLastError();
      return fail(KMG_ERR_OOM, "cudaMalloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
    }
    cap = want;
    return KMG_OK;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
};

// Page-locked host staging (small results that must not turn an async read-back into a blocking one)
struct HostBuf {
  void* p = nullptr;
  size_t cap = 0;
  int ensure(size_t bytes) {
    if (bytes <= cap) return KMG_OK;
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
    if (cudaMallocHost(&p, bytes) != cudaSuccess) {
      cudaGetLastError();
      p = nullptr;
      return fail(KMG_ERR_OOM, "cudaMallocHost(%zu bytes) failed", bytes);
    }
    cap = bytes;
    return KMG_OK;
  }
  void release() {
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
  }
};

struct Workspace {
  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr;  // created on first use: uploads that overlap kernels on `stream`
  cudaStream_t out_stream = nullptr;   // ... and read-backs that do (pipelined reduce)
  std::vector<cudaEvent_t> band_done;  // one per band in flight (banded upload + convert)
  std::vector<cudaEvent_t> band_out;   // remap of a band finished (pipelined reduce)
  cudaEvent_t palette_ready = nullptr;
  Buf in, out, small, work, dmin, blob;
  HostBuf stage;                // per-frame palettes and pass counts of a batch on their way to the caller
  JobState* h_state = nullptr;  // pinned
  void release() {
    stage.release();
    for (cudaEvent_t e : band_done) cudaEventDestroy(e);
    band_done.clear();
    for (cudaEvent_t e : band_out) cudaEventDestroy(e);
    band_out.clear();
    if (palette_ready) cudaEventDestroy(palette_ready);
    palette_ready = nullptr;
    if (copy_stream) cudaStreamDestroy(copy_stream);
    copy_stream = nullptr;
    if (out_stream) cudaStreamDestroy(out_stream);
    out_stream = nullptr;
    in.release();
    out.release();
    small.release();
    work.release();
    dmin.release();
    blob.release();
    if (h_state) cudaFreeHost(h_state);
    if (stream) cudaStreamDestroy(stream);
  }
};

struct kmg_ctx {
  int device = 0;
  int sms = 0;
  float* d_lut = nullptr;
  cudaStream_t stream = nullptr;  // context stream for set-up work (always synchronised before returning)
  std::mutex mu;
  std::vector<Workspace*> pool;
  std::atomic<uint64_t> launches{0};
  int lloyd_sel[3] = {0, 0, 0};    // LLOYD_VARIANTS index in use for k <= 8 / 16 / 32
  int lloyd_nocst[3] = {0, 0, 0};  // ... and the one taken when no constant-bank slot is free
  int occ_lloyd[32] = {0};
  // fused small-image k-means: dynamic shared memory available per variant (0 = not launchable)
  // and whether clusters of 8 / 16 CTAs can be scheduled
  size_t small_dyn[3] = {0, 0, 0};
  bool small_cluster_ok[3][2] = {{false, false}, {false, false}, {false, false}};
  int small_occ1[3] = {0, 0, 0};  // resident CTAs per SM in throughput mode (no planes in shared memory)
  // multi-GPU
  int n_ranks = 1, rank = 0;
#if KMG_HAVE_NCCL_HEADER
  ncclComm_t comm = nullptr;
#endif
  // peer mailboxes for the in-kernel exchange of the per-pass sums (PeerXchg); p2p == false: NCCL all-reduce
  bool p2p = false;
  void* mbox_own = nullptr;                // this GPU's mailbox (cudaMalloc, exported through CUDA IPC)
  void* mbox_peer[MAX_PEERS] = {nullptr};  // every rank's mailbox as mapped here (own entry = mbox_own)
  uint32_t xchg_seq = 0x1234567u;          // flag sequence base handed to the next sharded job
  std::atomic<int> live_sharded{0};        // sharded jobs alive on this context (at most one)
  // constant-bank table slots (kmg_kernels.cuh: c_tab), handed to jobs with k <= 8
  void* c_tab_dev = nullptr;
  bool big_block_acc = true;   // KMG_LLOYDG_BLOCKACC=0: accumulate through L2 atomics instead of shared memory
  void* c_big_dev = nullptr;   // c_big addressed as global memory
  bool big_const = true;       // KMG_LLOYD_BIGCONST=0: chunk loop of the k > 32 pass from shared memory only
  bool init_eager = false;     // KMG_INIT_EAGER=1: one full sweep per init round instead of the lazy cooperative launch
  int init_eager_rounds = 8;   // full sweeps before the lazy launch takes over (KMG_INIT_EAGER_ROUNDS)
  uint32_t init_lazy_min_k = 32;  // lazy rounds only for k above this (KMG_INIT_LAZY_MIN_K; tests lower it)
  int block_flush_log2 = 19;   // KMG_BLOCKACC_FLUSH_LOG2 (10..19): drain interval of the block accumulators (tests)
};

// Ownership of one slot of the constant-bank tables (c_tab).  Copies of a job (the
// per-chunk copies of a batched remap) do not own the slot: the copy constructor leaves it empty.
struct CSlot {
  int slot = -1, device = 0;
  CSlot() = default;
  CSlot(const CSlot&) {}
  CSlot& operator=(const CSlot&) { return *this; }
  ~CSlot();
};

// Ownership of c_big (the large constant-bank table of the k > 32 pass): one job per device; copies of
// a job do not own it.
struct CBig {
  int device = -1;  // >= 0: held
  CBig() = default;
  CBig(const CBig&) {}
  CBig& operator=(const CBig&) { return *this; }
  ~CBig();
};

struct kmg_job {
  kmg_ctx* ctx = nullptr;
  JobPtrs P{};
  void* blob = nullptr;
  bool owns_blob = false;
  float* dmin = nullptr;
  bool owns_dmin = false;
  JobState* h_state = nullptr;  // pinned
  bool owns_h_state = false;
  const float4* work = nullptr;
  uint32_t w = 0, h = 0, k = 0;
  int color_space = 0;
  kmg_opts opts{};
  // shard of a distributed image
  const uint8_t* rgba_src = nullptr;  // set: the work plane is still to be written by the first init round (fused conversion)
  bool sharded = false;
  uint32_t global_w = 0, global_h = 0, row_offset = 0;
  float* d_xfer = nullptr;  // 4 floats, colour broadcast during distributed init
  uint32_t xchg_base = 0;   // PeerXchg::seq_base of this job
  // slot of the constant-bank table (small-k passes), taken at the first pass, returned with the job
  CSlot cs_owner;
  CBig cbig_owner;
  bool cbig_tried = false;
  int cslot = -1;  // == cs_owner.slot once acquired; plain copy so that copies of the job can launch with it
  bool cslot_tried = false;
};

// The constant bank belongs to the device (one copy of the module per device), not to a context:
// the free list is per device and process-wide.
static std::mutex g_cslot_mu;
static bool g_cbig_taken[64];  // per device: some job holds c_big (guarded by g_cslot_mu)
static std::vector<int> g_cslots_free[64];  // per device
static bool g_cslots_ready[64];
static int cslot_acquire(kmg_job* j) {
  kmg_ctx* ctx = j->ctx;
  if (!ctx->c_tab_dev || ctx->device >= 64) return -1;
  std::lock_guard<std::mutex> g(g_cslot_mu);
  std::vector<int>& fl = g_cslots_free[ctx->device];
  if (!g_cslots_ready[ctx->device]) {
    g_cslots_ready[ctx->device] = true;
    for (int i = CTAB_SLOTS - 1; i >= 0; --i) fl.push_back(i);
  }
  if (fl.empty()) return -1;
  int s = fl.back();
  fl.pop_back();
  j->cs_owner.slot = s;
  j->cs_owner.device = ctx->device;
  return s;
}
// Jobs are only dropped after the work they enqueued has been waited for (every blocking entry
// point synchronises; kmg_job_destroy frees device memory first, which synchronises the device).
CBig::~CBig() {
  if (device >= 0 && device < 64) {
    std::lock_guard<std::mutex> g(g_cslot_mu);
    g_cbig_taken[device] = false;
  }
}
CSlot::~CSlot() {
  if (slot >= 0) {
    std::lock_guard<std::mutex> g(g_cslot_mu);
    g_cslots_free[device].push_back(slot);
  }
}

#if KMG_HAVE_NCCL_HEADER
static int p2p_setup(kmg_ctx* ctx);
static void p2p_teardown(kmg_ctx* ctx);
#endif

// Accumulator copies: the thread-private kernels (k <= 32) flush rarely, 8 copies are plenty; the
// global-reduction kernel adds 4 values per pixel, so every resident block gets its own copy
// (bounded to 8 MiB) and same-address serialisation in L2 disappears.
static uint32_t job_acc_copies(uint32_t k) {
  if (k <= 2048) return 8;  // private slots / block accumulators: one drain per block, folded by the last block
  size_t cap = ((size_t)8 << 20) / ((size_t)k * 32);
  return (uint32_t)std::max<size_t>(8, std::min<size_t>(304, cap));
}
static size_t job_blob_bytes(uint32_t k) {
  size_t kp = pad32(k);
  return 256 + (size_t)k * 16 + kp * sizeof(CentRec) + (size_t)(job_acc_copies(k) + 1) * k * 4 * 8 + (size_t)k * 8 + (size_t)k * 4 + 64;
}
static void job_carve(kmg_job* j, void* blob) {
  unsigned char* b = (unsigned char*)blob;
  size_t kp = pad32(j->k);
  j->P.st = (JobState*)b;
  b += 256;
  j->P.cent = (float4*)b;
  b += (size_t)j->k * 16;
  j->P.tab = (CentRec*)b;
  b += kp * sizeof(CentRec);
  j->P.acc = (long long*)b;
  j->P.acc_copies = job_acc_copies(j->k);
  b += (size_t)j->P.acc_copies * j->k * 4 * 8;
  j->P.last = (long long*)b;
  b += (size_t)j->k * 4 * 8;
  j->P.keys = (unsigned long long*)b;
  b += (size_t)j->k * 8;
  j->P.pal = (uint32_t*)b;
  b += (size_t)j->k * 4;
  j->d_xfer = (float*)(((uintptr_t)b + 15) & ~(uintptr_t)15);
}

static Workspace* ws_acquire(kmg_ctx* ctx) {
  {
    std::lock_guard<std::mutex> g(ctx->mu);
    if (!ctx->pool.empty()) {
      Workspace* w = ctx->pool.back();
      ctx->pool.pop_back();
      return w;
    }
  }
  Workspace* w = new Workspace();
  if (cudaStreamCreateWithFlags(&w->stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaMallocHost((void**)&w->h_state, sizeof(JobState)) != cudaSuccess) {
    cudaGetLastError();
    w->release();
    delete w;
    return nullptr;
  }
  return w;
}
static void ws_release(kmg_ctx* ctx, Workspace* w) {
  std::lock_guard<std::mutex> g(ctx->mu);
  ctx->pool.push_back(w);
}
// Hands the workspace back when the call returns — on every path.  An early error return can leave
// copies to or from the CALLER's buffers in flight on any of the three streams; they are waited for
// here, before the caller gets control back (and may free those buffers) and before another call can
// take the workspace.  On the success paths the streams are idle already and the three calls cost
// nothing measurable.
struct WsGuard {
  kmg_ctx* ctx;
  Workspace* ws;
  ~WsGuard() {
    if (!ws) return;
    if (ws->copy_stream) cudaStreamSynchronize(ws->copy_stream);
    if (ws->out_stream) cudaStreamSynchronize(ws->out_stream);
    if (ws->stream) cudaStreamSynchronize(ws->stream);
    cudaGetLastError();
    ws_release(ctx, ws);
  }
};

static inline int grid_for(kmg_ctx* ctx, unsigned long long items, int per_block, int blocks_per_sm) {
  unsigned long long blocks = (items + per_block - 1) / per_block;
  unsigned long long cap = (unsigned long long)ctx->sms * blocks_per_sm;
  if (blocks < 1) blocks = 1;
  return (int)std::min(blocks, cap);
}

#define LAUNCHED(ctx) (ctx)->launches.fetch_add(1, std::memory_order_relaxed)
#define CHECK_LAUNCH() CU(cudaGetLastError())


// ------------------------------------------------------------------------------------------------
// fused small-image k-means (kmg_small.cuh): capability probe, plan, launch

struct SmallPlan {
  int variant = -1;          // 0: <8,512>, 1: <16,512>, 2: <32,256>
  unsigned int kcap = 0, threads = 0, csize = 0, ppc = 0;
  size_t smem = 0;
  // throughput mode (large batches): csize 1, `grid` persistent CTAs, planes in `scratch` bytes of HBM/L2
  bool throughput = false;
  unsigned int grid = 0;
  size_t scratch = 0;
};
static const void* small_fn(int variant) {
  return variant == 0 ? (const void*)SMALL8 : variant == 1 ? (const void*)SMALL16 : (const void*)SMALL32;
}
static const unsigned int SMALL_KCAP[3] = {8, 16, 32};
static const unsigned int SMALL_THREADS[3] = {512, 512, 256};

static void small_probe(kmg_ctx* ctx, const cudaDeviceProp& prop) {
  for (int v = 0; v < 3; ++v) {
    const void* fn = small_fn(v);
    cudaFuncAttributes fa;
    if (cudaFuncGetAttributes(&fa, fn) != cudaSuccess) {
      cudaGetLastError();
      continue;
    }
    if ((size_t)prop.sharedMemPerBlockOptin <= fa.sharedSizeBytes + 1024) continue;
    const size_t dyn = (size_t)prop.sharedMemPerBlockOptin - fa.sharedSizeBytes;
    if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn) != cudaSuccess ||
        cudaFuncSetAttribute(fn, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) {
      cudaGetLastError();
      continue;
    }
    ctx->small_dyn[v] = dyn;
    {
      int occ = 0;
      const size_t smem1 = small_smem_bytes(0, SMALL_KCAP[v], SMALL_THREADS[v], 1);
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, (int)SMALL_THREADS[v], smem1) == cudaSuccess)
        ctx->small_occ1[v] = occ;
      else
        cudaGetLastError();
    }
    for (int c = 0; c < 2; ++c) {
      const unsigned int csize = c == 0 ? 8u : 16u;
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = dim3(csize);
      cfg.blockDim = dim3(SMALL_THREADS[v]);
      cfg.dynamicSmemBytes = dyn;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = csize;
      at[0].val.clusterDim.y = 1;
      at[0].val.clusterDim.z = 1;
      cfg.attrs = at;
      cfg.numAttrs = 1;
      int n = 0;
      if (cudaOccupancyMaxActiveClusters(&n, fn, &cfg) == cudaSuccess && n > 0)
        ctx->small_cluster_ok[v][c] = true;
      else
        cudaGetLastError();
    }
  }
}

// Can the image (n clustered pixels, k centroids) run in the fused kernel?  A single image takes
// the largest cluster (lowest latency); batches take the smallest one that fits (throughput).
static bool small_plan(kmg_ctx* ctx, unsigned long long n, uint32_t k, uint32_t n_frames, SmallPlan* plan) {
  if (k > 32 || n == 0 || n > (1ull << 20)) return false;
  const int v = k <= 8 ? 0 : (k <= 16 ? 1 : 2);
  if (ctx->small_dyn[v] == 0) return false;
  // A cluster finishes one image in ~0.15-0.25 ms but leaves its 8-16 SMs idle during barriers; a
  // lone CTA needs ~1 ms per image and keeps its SM busy.  From ~40 frames on the lone CTAs win.
  if (n_frames >= 40 && ctx->small_occ1[v] > 0 && n <= 65536) {
    plan->variant = v;
    plan->kcap = SMALL_KCAP[v];
    plan->threads = SMALL_THREADS[v];
    plan->csize = 1;
    plan->ppc = (unsigned int)((n + 3) & ~3ull);
    plan->smem = small_smem_bytes(0, SMALL_KCAP[v], SMALL_THREADS[v], 1);
    plan->throughput = true;
    plan->grid = (unsigned int)std::min<unsigned long long>(n_frames, (unsigned long long)ctx->sms * ctx->small_occ1[v]);
    plan->scratch = (size_t)plan->grid * plan->ppc * 20;
    return true;
  }
  const int order_single[2] = {1, 0}, order_batch[2] = {0, 1};
  const int* order = n_frames > 1 ? order_batch : order_single;
  for (int o = 0; o < 2; ++o) {
    const int c = order[o];
    if (!ctx->small_cluster_ok[v][c]) continue;
    const unsigned int csize = c == 0 ? 8u : 16u;
    const unsigned int ppc = (unsigned int)(((n + csize - 1) / csize + 3) & ~3ull);
    const size_t smem = small_smem_bytes(ppc, SMALL_KCAP[v], SMALL_THREADS[v], csize);
    if (smem > ctx->small_dyn[v]) continue;
    plan->variant = v;
    plan->kcap = SMALL_KCAP[v];
    plan->threads = SMALL_THREADS[v];
    plan->csize = csize;
    plan->ppc = ppc;
    plan->smem = smem;
    return true;
  }
  return false;
}

static int launch_small(kmg_ctx* ctx, const SmallPlan& plan, const SmallParams& prm, const JobPtrs& J0, uint32_t n_frames,
                        cudaStream_t s) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(plan.throughput ? plan.grid : plan.csize * n_frames);
  cfg.blockDim = dim3(plan.threads);
  cfg.dynamicSmemBytes = plan.smem;
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = plan.csize;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  if (plan.variant == 0)
    CU(cudaLaunchKernelEx(&cfg, SMALL8, prm, J0));
  else if (plan.variant == 1)
    CU(cudaLaunchKernelEx(&cfg, SMALL16, prm, J0));
  else
    CU(cudaLaunchKernelEx(&cfg, SMALL32, prm, J0));
  LAUNCHED(ctx);
  return KMG_OK;
}

// ------------------------------------------------------------------------------------------------
// lifetime

extern "C" int kmg_abi_version(void) { return KMG_ABI_VERSION; }
extern "C" const char* kmg_last_error(void) { return g_last_error.c_str(); }

extern "C" void kmg_default_opts(kmg_opts* o) {
  if (!o) return;
  o->struct_size = sizeof(kmg_opts);
  o->max_dim = 256;
  o->max_iter = 128;
  o->check_every = 8;
  o->convergence = -1.0f;
  o->seed_x_frac = 0.5625f;
  o->seed_y_frac = 0.93359375f;
  o->seed_x = -1;
  o->seed_y = -1;
  o->flags = 0;
}

static kmg_opts resolve_opts(const kmg_opts* in) {
  kmg_opts o;
  kmg_default_opts(&o);
  if (in) {
    size_t n = std::min<size_t>(in->struct_size ? in->struct_size : sizeof(kmg_opts), sizeof(kmg_opts));
    memcpy(&o, in, n);
    o.struct_size = sizeof(kmg_opts);
  }
  return o;
}

static int ctx_setup(kmg_ctx* ctx, const cudaDeviceProp& prop);

extern "C" int kmg_create(int device, kmg_ctx** out) {
  if (!out) return fail(KMG_ERR_BAD_ARG, "kmg_create: out is NULL");
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    cudaGetLastError();
    return fail(KMG_ERR_CUDA, "kmg_create: no CUDA device available (%s); this library has no CPU fallback",
                e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
  }
  if (device < 0 || device >= count) return fail(KMG_ERR_BAD_ARG, "kmg_create: device %d out of range [0,%d)", device, count);
  CU(cudaSetDevice(device));
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10)
    return fail(KMG_ERR_UNSUPPORTED, "kmg_create: device %d is sm_%d%d; this library is built for sm_100a (B200) only",
                device, prop.major, prop.minor);
  kmg_ctx* ctx = new kmg_ctx();
  ctx->device = device;
  ctx->sms = prop.multiProcessorCount;
  const int r = ctx_setup(ctx, prop);
  if (r != KMG_OK) {
    kmg_destroy(ctx);  // releases whatever ctx_setup had created (stream, table, workspaces)
    return r;
  }
  *out = ctx;
  return KMG_OK;
}

static int ctx_setup(kmg_ctx* ctx, const cudaDeviceProp& prop) {
  CU(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
  CU(cudaMalloc((void**)&ctx->d_lut, 256 * sizeof(float)));
  k_build_srgb_table<<<1, 256, 0, ctx->stream>>>(ctx->d_lut);
  LAUNCHED(ctx);
  CHECK_LAUNCH();
  // opt in to > 48 KiB dynamic shared memory
  const int big = (int)(tab_smem_bytes(MAX_K) + MAX_K * 4);
  CU(cudaFuncSetAttribute(LLOYDG, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
  CU(cudaFuncSetAttribute(k_assign<256, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
  CU(cudaFuncSetAttribute(k_remap<0, 0, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
  CU(cudaFuncSetAttribute(k_remap<1, 0, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
  CU(cudaFuncSetAttribute(k_remap_meld, cudaFuncAttributeMaxDynamicSharedMemorySize, MAX_K * 16));
  small_probe(ctx, prop);
  CU(cudaGetSymbolAddress(&ctx->c_tab_dev, c_tab));
  CU(cudaGetSymbolAddress(&ctx->c_big_dev, c_big));
  if (const char* e = getenv("KMG_LLOYD_BIGCONST")) ctx->big_const = atoi(e) != 0;
  CU(cudaFuncSetAttribute(LLOYDGB, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
  if (const char* e = getenv("KMG_LLOYDG_BLOCKACC")) ctx->big_block_acc = atoi(e) != 0;
  if (const char* e = getenv("KMG_INIT_EAGER")) ctx->init_eager = atoi(e) != 0;
  if (const char* e = getenv("KMG_INIT_EAGER_ROUNDS")) ctx->init_eager_rounds = std::max(1, atoi(e));
  if (const char* e = getenv("KMG_INIT_LAZY_MIN_K")) ctx->init_lazy_min_k = (uint32_t)std::max(0, atoi(e));
  if (const char* e = getenv("KMG_BLOCKACC_FLUSH_LOG2")) ctx->block_flush_log2 = std::min(19, std::max(10, atoi(e)));
  CU(cudaFuncSetAttribute(LLOYDGS, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
  for (int v = 0; v < N_LLOYD_VARIANTS; ++v) {
    const LloydVariant& V = LLOYD_VARIANTS[v];
    CU(cudaFuncSetAttribute(V.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)V.smem));
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->occ_lloyd[v], V.fn, V.threads, V.smem));
  }
  {
    static const int kcaps[3] = {8, 16, 32};
    static const char* envs[3] = {"KMG_LLOYD8_VARIANT", "KMG_LLOYD16_VARIANT", "KMG_LLOYD32_VARIANT"};
    for (int c = 0; c < 3; ++c) {
      int sel = lloyd_variant_index(kcaps[c], 0);
      if (const char* e = getenv(envs[c])) {
        int i = lloyd_variant_index(kcaps[c], atoi(e));
        if (i >= 0) sel = i;
      }
      ctx->lloyd_sel[c] = sel;
      int alt = sel;  // first variant of the class that does not need the constant bank
      for (int v = 0; LLOYD_VARIANTS[alt].const_tab; ++v) alt = lloyd_variant_index(kcaps[c], v);
      ctx->lloyd_nocst[c] = alt;
    }
  }
  CU(cudaStreamSynchronize(ctx->stream));
  return KMG_OK;
}

extern "C" void kmg_destroy(kmg_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaDeviceSynchronize();
#if KMG_HAVE_NCCL_HEADER
  if (ctx->comm) {
    p2p_teardown(ctx);
    nccl_api().CommDestroy(ctx->comm);
  }
#endif
  for (Workspace* w : ctx->pool) {
    w->release();
    delete w;
  }
  if (ctx->d_lut) cudaFree(ctx->d_lut);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

extern "C" void* kmg_alloc_pinned(size_t bytes) {
  void* p = nullptr;
  if (bytes == 0 || cudaMallocHost(&p, bytes) != cudaSuccess) {
    cudaGetLastError();
    fail(KMG_ERR_OOM, "kmg_alloc_pinned(%zu) failed", bytes);
    return nullptr;
  }
  return p;
}
extern "C" void kmg_free_pinned(void* p) {
  if (p) cudaFreeHost(p);
}

extern "C" uint64_t kmg_launch_count(kmg_ctx* ctx) { return ctx ? ctx->launches.load() : 0; }

extern "C" void kmg_resized_dims(uint32_t w, uint32_t h, uint32_t max_size, uint32_t* ow, uint32_t* oh) {
  // core/src/structures.rs:79-89 (strict `width > height`; f32 arithmetic, truncation, max(1))
  uint32_t nw, nh;
  if (w > h) {
    nw = max_size;
    nh = std::max<uint32_t>((uint32_t)((float)h * (float)max_size / (float)w), 1u);
  } else {
    nw = std::max<uint32_t>((uint32_t)((float)w * (float)max_size / (float)h), 1u);
    nh = max_size;
  }
  if (ow) *ow = nw;
  if (oh) *oh = nh;
}

// ------------------------------------------------------------------------------------------------
// stage launchers (device pointers)

// NULL is the CUDA (legacy) default stream, as everywhere in the runtime API — it orders with the
// caller's default-stream work (torch's default stream is handle 0).
static cudaStream_t pick_stream(kmg_ctx*, void* stream) { return (cudaStream_t)stream; }

static int launch_convert(kmg_ctx* ctx, const uint8_t* d_rgba, uint64_t n, int cs, float* d_work, cudaStream_t s) {
  int grid = grid_for(ctx, n, 256, 8);
  k_convert<<<grid, 256, 0, s>>>((const uint32_t*)d_rgba, n, cs, ctx->d_lut, (float4*)d_work);
  LAUNCHED(ctx);
  CHECK_LAUNCH();
  return KMG_OK;
}
static int launch_resize(kmg_ctx* ctx, const uint8_t* src, uint32_t sw, uint32_t sh, uint8_t* dst, uint32_t dw,
                         uint32_t dh, cudaStream_t s) {
  int grid = grid_for(ctx, (unsigned long long)dw * dh, 256, 8);
  k_resize<<<grid, 256, 0, s>>>((const uint32_t*)src, sw, sh, (uint32_t*)dst, dw, dh);
  LAUNCHED(ctx);
  CHECK_LAUNCH();
  return KMG_OK;
}
static int launch_prepare(kmg_job* j, bool palette, cudaStream_t s) {
  k_prepare<<<1, 256, 0, s>>>(j->P, j->color_space, palette ? 1 : 0);
  LAUNCHED(j->ctx);
  CHECK_LAUNCH();
  return KMG_OK;
}

static constexpr unsigned int MBOX_XCAP = MAX_K * 8;  // 8-byte words per (parity, rank): k x 4 values, two words each
static constexpr size_t MBOX_DATA_BYTES = (size_t)2 * MAX_PEERS * MBOX_XCAP * 8;
static constexpr size_t MBOX_BYTES = MBOX_DATA_BYTES + (size_t)2 * MAX_PEERS * 4;

static PeerXchg peer_xchg(const kmg_ctx* ctx, const kmg_job* j, bool fused) {
  PeerXchg X;
  memset(&X, 0, sizeof(X));
  if (fused) {
    X.n_ranks = (unsigned int)ctx->n_ranks;
    X.rank = (unsigned int)ctx->rank;
    X.xcap = MBOX_XCAP;
    X.seq_base = j->xchg_base;
    for (int r = 0; r < ctx->n_ranks; ++r) {
      X.mbox[r] = (long long*)ctx->mbox_peer[r];
      X.flags[r] = (unsigned int*)((unsigned char*)ctx->mbox_peer[r] + MBOX_DATA_BYTES);
    }
  }
  return X;
}

static int launch_lloyd(kmg_job* j, cudaStream_t s) {
  kmg_ctx* ctx = j->ctx;
  const unsigned long long n = (unsigned long long)j->w * j->h;
  const bool dist = j->sharded && ctx->n_ranks > 1;
  const bool fused = dist && ctx->p2p;
  const int partial = dist ? (fused ? 2 : 1) : 0;
  const PeerXchg X = peer_xchg(ctx, j, fused);
  if (j->k <= 32) {
    const int cls = j->k <= 8 ? 0 : j->k <= 16 ? 1 : 2;
    int v = ctx->lloyd_sel[cls];
    if (LLOYD_VARIANTS[v].const_tab) {
      if (j->cslot < 0 && !j->cslot_tried) {
        j->cslot_tried = true;
        j->cslot = cslot_acquire(j);
      }
      if (j->cslot < 0) v = ctx->lloyd_nocst[cls];  // no free slot in the constant bank: shared-memory table
    }
    const LloydVariant& V = LLOYD_VARIANTS[v];
    if (V.const_tab && !j->P.ctab) {
      // first pass with a slot: the current table goes to the job's slot of the constant bank once; from
      // then on build_table (k_prepare, the last block of every pass) keeps the slot up to date itself
      float* slot = (float*)((char*)ctx->c_tab_dev + (size_t)j->cslot * CTAB_FLOATS * 4);
      CU(cudaMemcpyAsync(slot, j->P.tab, (size_t)V.kcap * sizeof(CentRec), cudaMemcpyDeviceToDevice, s));
      j->P.ctab = slot;
    }
    int grid = grid_for(ctx, n, (V.threads == 288 ? 256 : V.threads) * V.px, ctx->occ_lloyd[v]);
    V.fn<<<grid, V.threads, V.smem, s>>>(j->P, j->work, n, j->color_space, partial, X, V.const_tab ? j->cslot : 0, j->k);
  } else {
    size_t smem = tab_smem_bytes(pad32(j->k));
    int grid = grid_for(ctx, n, 256 * 4, smem > 100 * 1024 ? 1 : 2);
    if (j->k <= LLOYDGS_MAX_K && ctx->big_block_acc) {
      smem += (size_t)7 * pad32(j->k) * 4;
      grid = grid_for(ctx, n, 256 * 4, smem > 100 * 1024 ? 1 : 2);
      // c_big: one job per device at a time; the first pass that gets it copies the current table over
      // (rearranged by k_big_table), later passes find it refreshed by build_table
      if (ctx->big_const && pad32(j->k) <= CBIG_MAX_K && !j->P.cbig && !j->cbig_tried && ctx->device < 64) {
        j->cbig_tried = true;
        std::lock_guard<std::mutex> g(g_cslot_mu);
        if (!g_cbig_taken[ctx->device]) {
          g_cbig_taken[ctx->device] = true;
          j->cbig_owner.device = ctx->device;
          j->P.cbig = (float*)ctx->c_big_dev;
        }
        if (j->P.cbig) {
          k_big_table<<<(pad32(j->k) + 255) / 256, 256, 0, s>>>(j->P, pad32(j->k));
          LAUNCHED(ctx);
        }
      }
      if (j->P.cbig)
        LLOYDGB<<<grid, 256, smem, s>>>(j->P, j->work, n, j->color_space, partial, X, 0, j->k);
      else
      LLOYDGS<<<grid, 256, smem, s>>>(j->P, j->work, n, j->color_space, partial, X, ctx->block_flush_log2, j->k);
    } else {
      LLOYDG<<<grid, 256, smem, s>>>(j->P, j->work, n, j->color_space, partial, X, 0, j->k);
    }
  }
  LAUNCHED(ctx);
  CHECK_LAUNCH();
#if KMG_HAVE_NCCL_HEADER
  if (partial == 1) {
    NC(nccl_api().AllReduce(j->P.acc, j->P.acc, (size_t)j->k * 4, ncclInt64, ncclSum, ctx->comm, s));
    k_finalize<<<1, 256, 0, s>>>(j->P, j->color_space);
    LAUNCHED(ctx);
    CHECK_LAUNCH();
  }
#endif
  return KMG_OK;
}

// Grid of a remap launch: x = persistent blocks per frame, y = frames.  A single image gets
// sms x blocks/SM blocks; a batch shares that budget between its frames (at least 16 per frame so
// every block still amortises its table load over many pixel groups).
static dim3 remap_grid(kmg_ctx* ctx, unsigned long long groups, int blocks_per_sm, uint32_t n_frames) {
  unsigned long long need = std::max<unsigned long long>(1, (groups + 255) / 256);
  unsigned long long cap = (unsigned long long)ctx->sms * blocks_per_sm;
  if (n_frames > 1) cap = std::max<unsigned long long>(16, (2 * cap + n_frames - 1) / n_frames);
  return dim3((unsigned int)std::min(need, cap), n_frames, 1);
}

template <int MODE>
static int launch_remap_mode(kmg_job* j, const uint8_t* d_rgba, uint32_t w, unsigned long long n, uint8_t* d_out,
                             uint32_t n_frames, size_t blob_stride, cudaStream_t s) {
  kmg_ctx* ctx = j->ctx;
  const unsigned long long groups = (n + 3) / 4;
  const uint32_t* in = (const uint32_t*)d_rgba;
  uint32_t* out = (uint32_t*)d_out;
  if (j->k <= 8) {
    dim3 grid = remap_grid(ctx, groups, 4, n_frames);
    k_remap<MODE, 8, 256><<<grid, 256, tab_smem_bytes(8) + 8 * 4, s>>>(j->P, in, w, n, j->color_space, ctx->d_lut, out, blob_stride);
  } else if (j->k <= 16) {
    dim3 grid = remap_grid(ctx, groups, 4, n_frames);
    k_remap<MODE, 16, 256><<<grid, 256, tab_smem_bytes(16) + 16 * 4, s>>>(j->P, in, w, n, j->color_space, ctx->d_lut, out, blob_stride);
  } else {
    size_t smem = tab_smem_bytes(pad32(j->k)) + (size_t)pad32(j->k) * 4;
    dim3 grid = remap_grid(ctx, groups, smem > 100 * 1024 ? 1 : (smem > 60 * 1024 ? 2 : 3), n_frames);
    k_remap<MODE, 0, 256><<<grid, 256, smem, s>>>(j->P, in, w, n, j->color_space, ctx->d_lut, out, blob_stride);
  }
  LAUNCHED(ctx);
  CHECK_LAUNCH();
  return KMG_OK;
}

// n_frames > 1: frames of w x h pixels back to back in d_rgba / d_out, their job blobs blob_stride
// bytes apart starting at j's (all with j's k and colour space).  prepared: table, dither
// threshold and RGBA8 palette are already in the blob(s) (written by k_kmeans_small).
static int launch_remap(kmg_job* j, const uint8_t* d_rgba, uint32_t w, uint32_t h, int mode, uint8_t* d_out,
                        cudaStream_t s, bool prepared = false, uint32_t n_frames = 1, size_t blob_stride = 0) {
  const unsigned long long n = (unsigned long long)w * h;
  if (!prepared) {
    if (n_frames != 1) return fail(KMG_ERR_BAD_ARG, "batched remap needs prepared job blobs");
    TRY(launch_prepare(j, true, s));
  }
  for (uint32_t f0 = 0; f0 < n_frames; f0 += 32768) {  // gridDim.y <= 65535
    const uint32_t nf = std::min<uint32_t>(32768, n_frames - f0);
    kmg_job jf = *j;
    if (f0) {
      unsigned char* b = (unsigned char*)j->blob + (size_t)f0 * blob_stride;
      jf.blob = b;
      job_carve(&jf, b);
    }
    const uint8_t* in = d_rgba + (size_t)f0 * n * 4;
    uint8_t* out = d_out + (size_t)f0 * n * 4;
    if (mode == KMG_REPLACE) {
      TRY(launch_remap_mode<0>(&jf, in, w, n, out, nf, blob_stride, s));
    } else if (mode == KMG_DITHER) {
      TRY(launch_remap_mode<1>(&jf, in, w, n, out, nf, blob_stride, s));
    } else if (mode == KMG_MELD) {
      dim3 grid = remap_grid(j->ctx, n, 8, nf);
      k_remap_meld<<<grid, 256, (size_t)j->k * 16, s>>>(jf.P, (const uint32_t*)in, n, j->color_space, j->ctx->d_lut,
                                                       (uint32_t*)out, blob_stride);
      LAUNCHED(j->ctx);
      CHECK_LAUNCH();
    } else {
      return fail(KMG_ERR_BAD_ARG, "unknown reduce mode %d", mode);
    }
  }
  return KMG_OK;
}

// ------------------------------------------------------------------------------------------------
// jobs

// Distance plane of the initialisation: f32 running minimum + the u16 upper bounds and u16 fold
// counts of the lazy rounds (kmg_init_lazy.cuh), each part 16-byte aligned.
static size_t dmin_part(size_t bytes) { return (bytes + 15) & ~(size_t)15; }
static size_t dmin_bytes(unsigned long long n) { return dmin_part(n * 4) + 2 * dmin_part(n * 2); }

static int validate_dims(uint32_t w, uint32_t h, uint32_t k) {
  if (w == 0 || h == 0) return fail(KMG_ERR_BAD_ARG, "image is empty (%ux%u)", w, h);
  if ((unsigned long long)w * h >= (1ull << 32)) return fail(KMG_ERR_BAD_ARG, "image has >= 2^32 pixels (%ux%u)", w, h);
  if (k == 0) return fail(KMG_ERR_BAD_ARG, "k must be at least 1");
  if (k > (uint32_t)MAX_K) return fail(KMG_ERR_UNSUPPORTED, "k = %u exceeds the supported maximum of %d", k, MAX_K);
  return KMG_OK;
}

// Initialise a job over caller-provided storage (blob of job_blob_bytes(k), optional dmin plane).
static int job_setup(kmg_job* j, kmg_ctx* ctx, const float* d_work, uint32_t w, uint32_t h, uint32_t k, int cs,
                     const kmg_opts& o, void* blob, float* dmin, JobState* h_state, cudaStream_t s) {
  j->ctx = ctx;
  j->work = (const float4*)d_work;
  j->w = w;
  j->h = h;
  j->k = k;
  j->color_space = cs;
  j->opts = o;
  j->blob = blob;
  j->dmin = dmin;
  j->h_state = h_state;
  job_carve(j, blob);
  CU(cudaMemsetAsync(blob, 0, job_blob_bytes(k), s));
  JobState st;
  memset(&st, 0, sizeof(st));
  st.k = k;
  st.max_iter = o.max_iter ? o.max_iter : 1;
  st.check_every = o.check_every;
  st.conv_threshold = o.convergence >= 0.0f ? o.convergence : (cs == KMG_LAB ? 1.0f : 0.01f);  // lib.rs:189-194
  *h_state = st;
  CU(cudaMemcpyAsync(j->P.st, h_state, sizeof(JobState), cudaMemcpyHostToDevice, s));
  return KMG_OK;
}

static int job_read_state(kmg_job* j, cudaStream_t s) {
  CU(cudaMemcpyAsync(j->h_state, j->P.st, sizeof(JobState), cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));
  if (j->h_state->conv == PASS_FAULT)
    return fail(KMG_ERR_NCCL, "a peer GPU did not deliver its partial sums within 4 s, or had given up before (pass %u); "
                              "every rank of the job fails together and the communicator stays unusable until it is re-created",
                j->h_state->passes);
  return KMG_OK;
}

#if KMG_HAVE_NCCL_HEADER
// Distributed init: share the colour of a globally chosen pixel (exactly one rank contributes
// non-zero values, so the float sum is exact).
static int share_colour(kmg_job* j, unsigned int slot, cudaStream_t s) {
  kmg_ctx* ctx = j->ctx;
  NC(nccl_api().AllReduce(j->d_xfer, j->d_xfer, 4, ncclFloat32, ncclSum, ctx->comm, s));
  CU(cudaMemcpyAsync(j->P.cent + slot, j->d_xfer, 16, cudaMemcpyDeviceToDevice, s));
  return KMG_OK;
}
#endif

static int resolve_seed(const kmg_opts& o, uint32_t gw, uint32_t gh, unsigned long long* seed);

static int job_init_impl(kmg_job* j, uint32_t* pick_index, float* pick_dist, cudaStream_t s) {
  kmg_ctx* ctx = j->ctx;
  const unsigned long long n = (unsigned long long)j->w * j->h;
  const bool dist = j->sharded && ctx->n_ranks > 1;
  const uint32_t gw = j->sharded ? j->global_w : j->w;
  const uint32_t gh = j->sharded ? j->global_h : j->h;
  const unsigned long long offset = j->sharded ? (unsigned long long)j->row_offset * gw : 0ull;
  // plus_plus_init.wgsl:161-167 — seed pixel on the clustered image
  unsigned long long seed = 0;
  TRY(resolve_seed(j->opts, gw, gh, &seed));
  const bool seed_local = seed >= offset && seed - offset < n;
  if (!j->dmin) return fail(KMG_ERR_BAD_ARG, "job has no distance plane");
#if KMG_HAVE_NCCL_HEADER
  if (dist) CU(cudaMemsetAsync(j->d_xfer, 0, 16, s));
#endif
  const uint32_t* rgba_src = (const uint32_t*)j->rgba_src;
  if (rgba_src && (dist || j->k < 2)) {  // no fused first round on these paths: convert now
    TRY(launch_convert(ctx, j->rgba_src, n, j->color_space, (float*)j->work, s));
    rgba_src = nullptr;
  }
  k_init_seed<<<1, 32, 0, s>>>(j->P, j->work, seed_local ? seed - offset : 0ull, seed_local ? 1 : 0, rgba_src, ctx->d_lut,
                               j->color_space);
  LAUNCHED(ctx);
  CHECK_LAUNCH();
#if KMG_HAVE_NCCL_HEADER
  if (dist) {
    if (seed_local) CU(cudaMemcpyAsync(j->d_xfer, j->P.cent, 16, cudaMemcpyDeviceToDevice, s));
    TRY(share_colour(j, 0, s));
  }
#endif
  const int grid = grid_for(ctx, n, 256 * 4, 8);
  // single GPU: the round's last block resolves the winner (PICK 1); sharded with peer mailboxes:
  // arg-max and colour travel inside the round's launch (PICK 2); sharded over NCCL: PICK 0 + k_init_pick
  const bool fused = dist && ctx->p2p;
  const PeerXchg X = peer_xchg(ctx, j, fused);
  // All rounds in one cooperative launch that only refreshes the pixels that can still win
  // (kmg_init_lazy.cuh); KMG_INIT_EAGER=1 keeps the one-sweep-per-round kernels (also the NCCL path).
  // (measured at 8192^2: k = 16 takes 5.0 ms lazily against 3.7 ms with full sweeps, k = 64 9.7 against
  // 15.7, k = 256 27.5 against 63 — the lazy launch pays off once there are many rounds)
  if ((!dist || fused) && j->k > std::max<uint32_t>(ctx->init_lazy_min_k, (uint32_t)ctx->init_eager_rounds + 1) && !ctx->init_eager) {
    unsigned short* ub = (unsigned short*)((unsigned char*)j->dmin + dmin_part(n * 4));
    unsigned short* fold = (unsigned short*)((unsigned char*)ub + dmin_part(n * 2));
    const size_t smem = (size_t)j->k * 16 + (size_t)8 * LAZY_QCAP * 4;
    const void* fn = fused ? (const void*)k_init_lazy<2> : (const void*)k_init_lazy<1>;
    int per_sm = 0;
    CU(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, 256, smem));
    if (per_sm < 1) return fail(KMG_ERR_CUDA, "k_init_lazy does not fit on an SM (k = %u)", j->k);
    const unsigned long long steps = (n + 255) / 256;
    const int lazy_grid = (int)std::min<unsigned long long>((unsigned long long)ctx->sms * per_sm, std::max<unsigned long long>(1, (steps + 7) / 8));
    JobPtrs P = j->P;
    const float4* work = j->work;
    float* dm = j->dmin;
    unsigned long long nn = n, off = offset;
    PeerXchg Xc = X;
    // the first rounds as full sweeps, the bounds kept up to date; then the lazy launch
    const uint32_t j0 = std::min<uint32_t>(j->k, (uint32_t)ctx->init_eager_rounds + 1);
    for (uint32_t c = 1; c < j0; ++c) {
      if (fused && c == 1)
        k_init_round<true, 2><<<grid, 256, 0, s>>>(j->P, j->work, j->dmin, n, offset, c, X, ub);
      else if (fused)
        k_init_round<false, 2><<<grid, 256, 0, s>>>(j->P, j->work, j->dmin, n, offset, c, X, ub);
      else if (c == 1 && rgba_src)
        k_init_round<true, 1, true><<<grid, 256, 0, s>>>(j->P, j->work, j->dmin, n, offset, c, X, ub, rgba_src,
                                                         (float4*)j->work, ctx->d_lut, j->color_space);
      else if (c == 1)
        k_init_round<true, 1><<<grid, 256, 0, s>>>(j->P, j->work, j->dmin, n, offset, c, X, ub);
      else
        k_init_round<false, 1><<<grid, 256, 0, s>>>(j->P, j->work, j->dmin, n, offset, c, X, ub);
      LAUNCHED(ctx);
      CHECK_LAUNCH();
    }
    if (j0 < j->k) {
      CU(cudaMemsetAsync(fold, 0, n * 2, s));
      unsigned int j0v = j0;
      void* args[] = {&P, &work, &dm, &ub, &fold, &nn, &off, &Xc, &j0v};
      CU(cudaLaunchCooperativeKernel(fn, dim3(lazy_grid), dim3(256), args, smem, s));
      LAUNCHED(ctx);
    }
  } else
  for (uint32_t c = 1; c < j->k; ++c) {
    if (!dist) {
      if (c == 1 && rgba_src)
        k_init_round<true, 1, true><<<grid, 256, 0, s>>>(j->P, j->work, j->dmin, n, offset, c, X, nullptr, rgba_src,
                                                         (float4*)j->work, ctx->d_lut, j->color_space);
      else if (c == 1)
        k_init_round<true, 1><<<grid, 256, 0, s>>>(j->P, j->work, j->dmin, n, offset, c, X);
      else
        k_init_round<false, 1><<<grid, 256, 0, s>>>(j->P, j->work, j->dmin, n, offset, c, X);
      LAUNCHED(ctx);
      CHECK_LAUNCH();
      continue;
    }
    if (fused) {
      if (c == 1)
        k_init_round<true, 2><<<grid, 256, 0, s>>>(j->P, j->work, j->dmin, n, offset, c, X);
      else
        k_init_round<false, 2><<<grid, 256, 0, s>>>(j->P, j->work, j->dmin, n, offset, c, X);
      LAUNCHED(ctx);
      CHECK_LAUNCH();
      continue;
    }
    if (c == 1)
      k_init_round<true, 0><<<grid, 256, 0, s>>>(j->P, j->work, j->dmin, n, offset, c, X);
    else
      k_init_round<false, 0><<<grid, 256, 0, s>>>(j->P, j->work, j->dmin, n, offset, c, X);
    LAUNCHED(ctx);
    CHECK_LAUNCH();
#if KMG_HAVE_NCCL_HEADER
    NC(nccl_api().AllReduce(j->P.keys + c, j->P.keys + c, 1, ncclUint64, ncclMax, ctx->comm, s));
    CU(cudaMemsetAsync(j->d_xfer, 0, 16, s));
    CU(cudaMemsetAsync(j->P.cent + c, 0, 16, s));
#endif
    k_init_pick<<<1, 32, 0, s>>>(j->P, j->work, n, offset, c);
    LAUNCHED(ctx);
    CHECK_LAUNCH();
#if KMG_HAVE_NCCL_HEADER
    CU(cudaMemcpyAsync(j->d_xfer, j->P.cent + c, 16, cudaMemcpyDeviceToDevice, s));
    TRY(share_colour(j, c, s));
#endif
  }
  TRY(launch_prepare(j, false, s));
  if (pick_index || pick_dist) {
    std::vector<unsigned long long> keys(j->k);
    CU(cudaMemcpyAsync(keys.data(), j->P.keys, (size_t)j->k * 8, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    for (uint32_t c = 0; c < j->k; ++c) {
      unsigned long long key = keys[c];
      uint32_t bits = (uint32_t)(key >> 32);
      float d;
      memcpy(&d, &bits, 4);
      if (pick_index) pick_index[c] = c == 0 ? (uint32_t)seed : (uint32_t)((key >> 32) == 0 ? 0 : ((key & 0xffffffffull) ^ 15ull));
      if (pick_dist) pick_dist[c] = c == 0 ? 0.0f : d;
    }
  }
  return KMG_OK;
}

// core/src/modules.rs:763-840.  Passes are enqueued up to the next iteration at which the
// reference tests convergence; only then is the 64-byte state read back.
static int job_run_impl(kmg_job* j, uint32_t* passes_out, cudaStream_t s) {
  const uint32_t max_iter = j->opts.max_iter ? j->opts.max_iter : 1;
  const uint32_t every = j->opts.check_every;
  uint32_t it = 0;
  while (it < max_iter) {
    uint32_t stop = max_iter - 1;  // last iteration index of this chunk
    if (every != 0) {
      uint32_t next_check = it == 0 ? every : ((it + every - 1) / every) * every;
      if (next_check == 0) next_check = every;
      stop = std::min(stop, next_check);
    }
    for (; it <= stop; ++it) TRY(launch_lloyd(j, s));
    TRY(job_read_state(j, s));
    if (j->h_state->done) break;
  }
  if (passes_out) *passes_out = j->h_state->passes;
  return KMG_OK;
}

extern "C" int kmg_job_create(kmg_ctx* ctx, const float* d_work, uint32_t w, uint32_t h, uint32_t k, int cs,
                              const kmg_opts* opts, kmg_job** out) {
  if (!ctx || !d_work || !out) return fail(KMG_ERR_BAD_ARG, "kmg_job_create: NULL argument");
  TRY(validate_dims(w, h, k));
  if (cs != KMG_LAB && cs != KMG_RGB) return fail(KMG_ERR_BAD_ARG, "unknown colour space %d", cs);
  CU(cudaSetDevice(ctx->device));
  kmg_job* j = new kmg_job();
  void* blob = nullptr;
  float* dmin = nullptr;
  JobState* hs = nullptr;
  if (cudaMalloc(&blob, job_blob_bytes(k)) != cudaSuccess || cudaMalloc((void**)&dmin, dmin_bytes((unsigned long long)w * h)) != cudaSuccess ||
      cudaMallocHost((void**)&hs, sizeof(JobState)) != cudaSuccess) {
    cudaGetLastError();
    if (blob) cudaFree(blob);
    if (dmin) cudaFree(dmin);
    delete j;
    return fail(KMG_ERR_OOM, "kmg_job_create: allocation failed");
  }
  j->owns_blob = j->owns_dmin = j->owns_h_state = true;
  int r = job_setup(j, ctx, d_work, w, h, k, cs, resolve_opts(opts), blob, dmin, hs, ctx->stream);
  if (r == KMG_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess) r = fail(KMG_ERR_CUDA, "job setup failed");
  if (r != KMG_OK) {
    kmg_job_destroy(j);
    return r;
  }
  *out = j;
  return KMG_OK;
}

extern "C" void kmg_job_destroy(kmg_job* j) {
  if (!j) return;
  cudaSetDevice(j->ctx->device);
  if (j->sharded) j->ctx->live_sharded.store(0);
  if (j->owns_blob && j->blob) cudaFree(j->blob);
  if (j->owns_dmin && j->dmin) cudaFree(j->dmin);
  if (j->owns_h_state && j->h_state) cudaFreeHost(j->h_state);
  delete j;
}

extern "C" int kmg_job_set_shard(kmg_job* j, uint32_t gw, uint32_t gh, uint32_t row_offset) {
  if (!j) return fail(KMG_ERR_BAD_ARG, "kmg_job_set_shard: NULL job");
  if (gw != j->w || (unsigned long long)row_offset + j->h > gh)
    return fail(KMG_ERR_BAD_ARG, "shard rows [%u,%u) of width %u do not fit the %ux%u image", row_offset,
                row_offset + j->h, j->w, gw, gh);
  // One live sharded job per context: the mailboxes of the in-kernel exchange are indexed by
  // (pass parity, rank) only, so two sharded jobs stepping on different streams of one context would
  // overwrite each other's partial sums.  Destroy the previous job (or use a second kmg_ctx).
  if (!j->sharded) {
    int expected = 0;
    if (!j->ctx->live_sharded.compare_exchange_strong(expected, 1))
      return fail(KMG_ERR_BAD_ARG, "kmg_job_set_shard: this context already has a live sharded job (one at a time; "
                                   "destroy it first or use another kmg_ctx)");
  }
  j->sharded = true;
  // every rank creates its sharded jobs in the same order, so the bases agree across ranks
  j->xchg_base = j->ctx->xchg_seq;
  j->ctx->xchg_seq += 0x9E3779B1u;
  j->global_w = gw;
  j->global_h = gh;
  j->row_offset = row_offset;
  return KMG_OK;
}

extern "C" int kmg_job_init(kmg_job* j, uint32_t* pick_index, float* pick_dist, void* stream) {
  if (!j) return fail(KMG_ERR_BAD_ARG, "kmg_job_init: NULL job");
  CU(cudaSetDevice(j->ctx->device));
  return job_init_impl(j, pick_index, pick_dist, pick_stream(j->ctx, stream));
}

extern "C" int kmg_job_set_centroids(kmg_job* j, const float* c, void* stream) {
  if (!j || !c) return fail(KMG_ERR_BAD_ARG, "kmg_job_set_centroids: NULL argument");
  CU(cudaSetDevice(j->ctx->device));
  cudaStream_t s = pick_stream(j->ctx, stream);
  CU(cudaMemcpyAsync(j->P.cent, c, (size_t)j->k * 16, cudaMemcpyHostToDevice, s));
  TRY(launch_prepare(j, false, s));
  CU(cudaStreamSynchronize(s));
  return KMG_OK;
}

extern "C" int kmg_job_get_centroids(kmg_job* j, float* c, void* stream) {
  if (!j || !c) return fail(KMG_ERR_BAD_ARG, "kmg_job_get_centroids: NULL argument");
  CU(cudaSetDevice(j->ctx->device));
  cudaStream_t s = pick_stream(j->ctx, stream);
  CU(cudaMemcpyAsync(c, j->P.cent, (size_t)j->k * 16, cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));
  return KMG_OK;
}

extern "C" int kmg_job_step(kmg_job* j, uint32_t count, void* stream) {
  if (!j) return fail(KMG_ERR_BAD_ARG, "kmg_job_step: NULL job");
  CU(cudaSetDevice(j->ctx->device));
  cudaStream_t s = pick_stream(j->ctx, stream);
  for (uint32_t i = 0; i < count; ++i) TRY(launch_lloyd(j, s));
  return KMG_OK;
}

extern "C" int kmg_job_run(kmg_job* j, uint32_t* passes_out, void* stream) {
  if (!j) return fail(KMG_ERR_BAD_ARG, "kmg_job_run: NULL job");
  CU(cudaSetDevice(j->ctx->device));
  return job_run_impl(j, passes_out, pick_stream(j->ctx, stream));
}

extern "C" int kmg_job_stats(kmg_job* j, uint32_t* conv, uint32_t* passes, uint64_t* slow, void* stream) {
  if (!j) return fail(KMG_ERR_BAD_ARG, "kmg_job_stats: NULL job");
  CU(cudaSetDevice(j->ctx->device));
  TRY(job_read_state(j, pick_stream(j->ctx, stream)));
  if (conv) *conv = j->h_state->conv;
  if (passes) *passes = j->h_state->passes;
  if (slow) *slow = j->h_state->slow_pixels;
  return KMG_OK;
}

extern "C" int kmg_job_init_stats(kmg_job* j, uint32_t* sweeps, uint64_t* refreshed, uint64_t* folds, uint64_t* exact,
                                  void* stream) {
  if (!j) return fail(KMG_ERR_BAD_ARG, "kmg_job_init_stats: NULL job");
  CU(cudaSetDevice(j->ctx->device));
  TRY(job_read_state(j, pick_stream(j->ctx, stream)));
  if (sweeps) *sweeps = j->h_state->init_attempts;
  if (refreshed) *refreshed = j->h_state->init_refreshed;
  if (folds) *folds = j->h_state->init_folds;
  if (exact) *exact = j->h_state->init_exact;
  return KMG_OK;
}

extern "C" int kmg_job_get_sums(kmg_job* j, int64_t* sums_out, void* stream) {
  if (!j || !sums_out) return fail(KMG_ERR_BAD_ARG, "kmg_job_get_sums: NULL argument");
  CU(cudaSetDevice(j->ctx->device));
  cudaStream_t s = pick_stream(j->ctx, stream);
  CU(cudaMemcpyAsync(sums_out, j->P.last, (size_t)j->k * 32, cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));
  return KMG_OK;
}

// ------------------------------------------------------------------------------------------------
// device-resident stage entry points

extern "C" int kmg_dev_convert(kmg_ctx* ctx, const uint8_t* d_rgba, uint64_t n, int cs, float* d_work, void* stream) {
  if (!ctx || !d_rgba || !d_work) return fail(KMG_ERR_BAD_ARG, "kmg_dev_convert: NULL argument");
  if (n == 0) return fail(KMG_ERR_BAD_ARG, "kmg_dev_convert: empty input");
  CU(cudaSetDevice(ctx->device));
  return launch_convert(ctx, d_rgba, n, cs, d_work, pick_stream(ctx, stream));
}

extern "C" int kmg_dev_resize(kmg_ctx* ctx, const uint8_t* d_src, uint32_t sw, uint32_t sh, uint8_t* d_dst,
                              uint32_t dw, uint32_t dh, void* stream) {
  if (!ctx || !d_src || !d_dst) return fail(KMG_ERR_BAD_ARG, "kmg_dev_resize: NULL argument");
  if (!sw || !sh || !dw || !dh) return fail(KMG_ERR_BAD_ARG, "kmg_dev_resize: empty image");
  CU(cudaSetDevice(ctx->device));
  return launch_resize(ctx, d_src, sw, sh, d_dst, dw, dh, pick_stream(ctx, stream));
}

// A throw-away job over fixed centroids (remap / assign with host-provided centroids).
struct TempJob {
  kmg_job job;
  void* blob = nullptr;
  JobState* hs = nullptr;
  ~TempJob() {
    if (blob) cudaFree(blob);
    if (hs) cudaFreeHost(hs);
  }
  int setup(kmg_ctx* ctx, uint32_t w, uint32_t h, const float* cent_host, uint32_t k, int cs, cudaStream_t s) {
    CU(cudaMalloc(&blob, job_blob_bytes(k)));
    CU(cudaMallocHost((void**)&hs, sizeof(JobState)));
    kmg_opts o;
    kmg_default_opts(&o);
    TRY(job_setup(&job, ctx, nullptr, w, h, k, cs, o, blob, nullptr, hs, s));
    // fixed centroids: force w = 1.0 like CentroidsBuffer::fixed_centroids (structures.rs:536,540)
    std::vector<float> c(cent_host, cent_host + (size_t)k * 4);
    CU(cudaMemcpyAsync(job.P.cent, c.data(), (size_t)k * 16, cudaMemcpyHostToDevice, s));
    CU(cudaStreamSynchronize(s));  // c goes out of scope
    return KMG_OK;
  }
};

extern "C" int kmg_dev_assign(kmg_ctx* ctx, const float* d_work, uint64_t n, const float* cent, uint32_t k,
                              uint32_t* d_labels, void* stream) {
  if (!ctx || !d_work || !cent || !d_labels) return fail(KMG_ERR_BAD_ARG, "kmg_dev_assign: NULL argument");
  if (n == 0 || n >= (1ull << 32)) return fail(KMG_ERR_BAD_ARG, "kmg_dev_assign: bad pixel count");
  TRY(validate_dims(1, 1, k));
  CU(cudaSetDevice(ctx->device));
  cudaStream_t s = pick_stream(ctx, stream);
  TempJob t;
  TRY(t.setup(ctx, (uint32_t)n, 1, cent, k, KMG_LAB, s));
  TRY(launch_prepare(&t.job, false, s));
  size_t smem = tab_smem_bytes(pad32(k));
  int grid = grid_for(ctx, n, 256 * 4, smem > 100 * 1024 ? 1 : 2);
  k_assign<256, 4><<<grid, 256, smem, s>>>(t.job.P, (const float4*)d_work, n, d_labels);
  LAUNCHED(ctx);
  CHECK_LAUNCH();
  CU(cudaStreamSynchronize(s));
  return KMG_OK;
}

template <int SEARCH>
static int launch_audit(kmg_job* j, const float4* work, const uint32_t* rgba, uint32_t w, unsigned long long n, int mode,
                        unsigned long long* d_counters, cudaStream_t s) {
  kmg_ctx* ctx = j->ctx;
  const unsigned int kp = (SEARCH == 0 || SEARCH == 3) ? 8u : (SEARCH == 1 ? 16u : pad32(j->k));
  const size_t smem = tab_smem_bytes(kp);
  const int grid = ctx->sms * (smem > 100 * 1024 ? 1 : (smem > 48 * 1024 ? 2 : 4));
  if (mode == 0) {
    CU(cudaFuncSetAttribute(k_audit<SEARCH, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_audit<SEARCH, 0><<<grid, 256, smem, s>>>(j->P, work, rgba, w, n, j->color_space, ctx->d_lut, d_counters);
  } else if (mode == 1) {
    CU(cudaFuncSetAttribute(k_audit<SEARCH, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_audit<SEARCH, 1><<<grid, 256, smem, s>>>(j->P, work, rgba, w, n, j->color_space, ctx->d_lut, d_counters);
  } else {
    CU(cudaFuncSetAttribute(k_audit<SEARCH, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_audit<SEARCH, 2><<<grid, 256, smem, s>>>(j->P, work, rgba, w, n, j->color_space, ctx->d_lut, d_counters);
  }
  LAUNCHED(ctx);
  CHECK_LAUNCH();
  return KMG_OK;
}

extern "C" int kmg_dev_audit(kmg_ctx* ctx, const float* d_work, const uint8_t* d_rgba, uint32_t w, uint32_t h,
                             const float* cent, uint32_t k, int cs, int search, int mode, uint64_t* wrong_out,
                             uint64_t* uncertified_out, void* stream) {
  if (!ctx || !cent || !wrong_out) return fail(KMG_ERR_BAD_ARG, "kmg_dev_audit: NULL argument");
  if (mode < 0 || mode > 2 || search < 0 || search > 4) return fail(KMG_ERR_BAD_ARG, "kmg_dev_audit: unknown search %d / mode %d", search, mode);
  if (mode == 0 ? !d_work : !d_rgba) return fail(KMG_ERR_BAD_ARG, "kmg_dev_audit: mode %d needs the %s", mode, mode == 0 ? "work plane" : "RGBA8 image");
  TRY(validate_dims(w, h, k));
  if (((search == 0 || search == 3) && k > 8) || (search == 1 && k > 16))
    return fail(KMG_ERR_BAD_ARG, "kmg_dev_audit: search %d holds at most %d centroids", search, search == 1 ? 16 : 8);
  if (cs != KMG_LAB && cs != KMG_RGB) return fail(KMG_ERR_BAD_ARG, "unknown colour space %d", cs);
  CU(cudaSetDevice(ctx->device));
  cudaStream_t s = pick_stream(ctx, stream);
  TempJob t;
  TRY(t.setup(ctx, w, h, cent, k, cs, s));
  TRY(launch_prepare(&t.job, true, s));  // table, dither threshold, palette
  struct Scratch : Buf {
    ~Scratch() { release(); }
  } counters;
  TRY(counters.ensure(16));
  CU(cudaMemsetAsync(counters.p, 0, 16, s));
  const unsigned long long n = (unsigned long long)w * h;
  unsigned long long* dc = (unsigned long long*)counters.p;
  const float4* work = (const float4*)d_work;
  const uint32_t* rgba = (const uint32_t*)d_rgba;
  if (search == 4) {  // margin probe
    const size_t smem = tab_smem_bytes(pad32(k));
    const int grid = ctx->sms * (smem > 100 * 1024 ? 1 : (smem > 48 * 1024 ? 2 : 4));
    if (mode == 0) {
      CU(cudaFuncSetAttribute(k_audit_margin<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      k_audit_margin<0><<<grid, 256, smem, s>>>(t.job.P, work, rgba, w, n, cs, ctx->d_lut, dc);
    } else if (mode == 1) {
      CU(cudaFuncSetAttribute(k_audit_margin<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      k_audit_margin<1><<<grid, 256, smem, s>>>(t.job.P, work, rgba, w, n, cs, ctx->d_lut, dc);
    } else {
      CU(cudaFuncSetAttribute(k_audit_margin<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      k_audit_margin<2><<<grid, 256, smem, s>>>(t.job.P, work, rgba, w, n, cs, ctx->d_lut, dc);
    }
    LAUNCHED(ctx);
    CHECK_LAUNCH();
  }
  int r = search == 4   ? KMG_OK
          : search == 0 ? launch_audit<0>(&t.job, work, rgba, w, n, mode, dc, s)
          : search == 1 ? launch_audit<1>(&t.job, work, rgba, w, n, mode, dc, s)
          : search == 2 ? launch_audit<2>(&t.job, work, rgba, w, n, mode, dc, s)
                        : launch_audit<3>(&t.job, work, rgba, w, n, mode, dc, s);
  if (r != KMG_OK) {
    cudaStreamSynchronize(s);
    return r;
  }
  unsigned long long h_c[2] = {0, 0};
  cudaError_t e = cudaMemcpyAsync(h_c, dc, 16, cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  if (e != cudaSuccess) return fail(KMG_ERR_CUDA, "kmg_dev_audit: %s", cudaGetErrorString(e));
  *wrong_out = h_c[0];
  if (uncertified_out) *uncertified_out = h_c[1];
  return KMG_OK;
}

extern "C" int kmg_dev_remap(kmg_ctx* ctx, const uint8_t* d_rgba, uint32_t w, uint32_t h, const float* cent,
                             uint32_t k, int cs, int mode, uint8_t* d_out, void* stream) {
  if (!ctx || !d_rgba || !cent || !d_out) return fail(KMG_ERR_BAD_ARG, "kmg_dev_remap: NULL argument");
  TRY(validate_dims(w, h, k));
  CU(cudaSetDevice(ctx->device));
  cudaStream_t s = pick_stream(ctx, stream);
  TempJob t;
  TRY(t.setup(ctx, w, h, cent, k, cs, s));
  TRY(launch_remap(&t.job, d_rgba, w, h, mode, d_out, s));
  CU(cudaStreamSynchronize(s));
  return KMG_OK;
}

extern "C" int kmg_dev_remap_job(kmg_ctx* ctx, const uint8_t* d_rgba, uint32_t w, uint32_t h, kmg_job* job, int mode,
                                 uint8_t* d_out, void* stream) {
  if (!ctx || !d_rgba || !job || !d_out) return fail(KMG_ERR_BAD_ARG, "kmg_dev_remap_job: NULL argument");
  TRY(validate_dims(w, h, job->k));
  CU(cudaSetDevice(ctx->device));
  return launch_remap(job, d_rgba, w, h, mode, d_out, pick_stream(ctx, stream));
}

extern "C" int kmg_dev_synth(kmg_ctx* ctx, uint8_t* d_rgba, uint64_t first, uint64_t n, uint32_t frame, uint32_t seed,
                             uint32_t blobs, void* stream) {
  if (!ctx || !d_rgba || n == 0) return fail(KMG_ERR_BAD_ARG, "kmg_dev_synth: bad argument");
  CU(cudaSetDevice(ctx->device));
  int grid = grid_for(ctx, n, 256, 8);
  k_synth<<<grid, 256, 0, pick_stream(ctx, stream)>>>((uint32_t*)d_rgba, first, n, frame, seed, blobs);
  LAUNCHED(ctx);
  CHECK_LAUNCH();
  return KMG_OK;
}

extern "C" int kmg_dev_srgb_table(kmg_ctx* ctx, float* table_out) {
  if (!ctx || !table_out) return fail(KMG_ERR_BAD_ARG, "kmg_dev_srgb_table: NULL argument");
  CU(cudaSetDevice(ctx->device));
  CU(cudaMemcpy(table_out, ctx->d_lut, 256 * sizeof(float), cudaMemcpyDeviceToHost));
  return KMG_OK;
}

extern "C" int kmg_dev_fp32_peak(kmg_ctx* ctx, double* fma_per_second_out) {
  if (!ctx || !fma_per_second_out) return fail(KMG_ERR_BAD_ARG, "kmg_dev_fp32_peak: NULL argument");
  CU(cudaSetDevice(ctx->device));
  float* d = nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  CU(cudaMalloc((void**)&d, 4));
  if (cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess) {
    if (e0) cudaEventDestroy(e0);
    cudaFree(d);
    cudaGetLastError();
    return fail(KMG_ERR_CUDA, "kmg_dev_fp32_peak: cudaEventCreate failed");
  }
  const int iters = 4096, grid = ctx->sms * 8;
  float best_ms = 1e30f;
  cudaError_t e = cudaSuccess;
  for (int rep = 0; rep < 4 && e == cudaSuccess; ++rep) {  // the first launch warms the clocks up
    cudaEventRecord(e0, ctx->stream);
    k_fp32_peak<<<grid, 256, 0, ctx->stream>>>(d, iters);
    LAUNCHED(ctx);
    cudaEventRecord(e1, ctx->stream);
    e = cudaEventSynchronize(e1);
    float ms = 0.f;
    if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms < best_ms) best_ms = ms;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d);
  if (e != cudaSuccess) return fail(KMG_ERR_CUDA, "kmg_dev_fp32_peak: %s", cudaGetErrorString(e));
  *fma_per_second_out = (double)grid * 256.0 * iters * 64.0 / ((double)best_ms * 1e-3);
  return KMG_OK;
}

extern "C" int kmg_dev_fast_lab_error(kmg_ctx* ctx, float* max_err_out) {
  if (!ctx || !max_err_out) return fail(KMG_ERR_BAD_ARG, "kmg_dev_fast_lab_error: NULL argument");
  CU(cudaSetDevice(ctx->device));
  float* d = nullptr;
  CU(cudaMalloc((void**)&d, 4));
  CU(cudaMemsetAsync(d, 0, 4, ctx->stream));
  k_fast_lab_error<<<ctx->sms * 8, 256, 0, ctx->stream>>>(ctx->d_lut, d);
  LAUNCHED(ctx);
  cudaError_t e = cudaMemcpyAsync(max_err_out, d, 4, cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  cudaFree(d);
  if (e != cudaSuccess) return fail(KMG_ERR_CUDA, "kmg_dev_fast_lab_error: %s", cudaGetErrorString(e));
  return KMG_OK;
}

// ------------------------------------------------------------------------------------------------
// k-means on a device-resident RGBA8 image using workspace storage.
// operations::extract_palette_kmeans (core/src/operations.rs:15-88).

// Seed pixel on the clustered image (plus_plus_init.wgsl:161-167).
static int resolve_seed(const kmg_opts& o, uint32_t gw, uint32_t gh, unsigned long long* seed) {
  int64_t sx = o.seed_x >= 0 ? o.seed_x : (int64_t)(int32_t)((float)gw * o.seed_x_frac);
  int64_t sy = o.seed_y >= 0 ? o.seed_y : (int64_t)(int32_t)((float)gh * o.seed_y_frac);
  if (sx < 0 || sy < 0 || sx >= (int64_t)gw || sy >= (int64_t)gh)
    return fail(KMG_ERR_BAD_ARG, "seed pixel (%lld,%lld) outside the %ux%u image", (long long)sx, (long long)sy, gw, gh);
  *seed = (unsigned long long)sy * gw + (unsigned long long)sx;
  return KMG_OK;
}

// The fused path: n_frames images of w x h (back to back in d_rgba), one thread-block cluster
// each, job blobs blob_stride apart in `blob`.  Leaves centroids, state, table, dither threshold
// and RGBA8 palette in every blob.  Asynchronous on s.
static int kmeans_small_on_device(kmg_ctx* ctx, const SmallPlan& plan, const uint8_t* d_rgba, uint32_t n_frames,
                                  uint32_t w, uint32_t h, uint32_t iw, uint32_t ih, uint32_t k, int cs, const kmg_opts& o,
                                  void* blob, size_t blob_stride, int tail, void* scratch, kmg_job* job, cudaStream_t s) {
  unsigned long long seed = 0;
  TRY(resolve_seed(o, iw, ih, &seed));
  job->ctx = ctx;
  job->work = nullptr;
  job->w = iw;
  job->h = ih;
  job->k = k;
  job->color_space = cs;
  job->opts = o;
  job->blob = blob;
  job_carve(job, blob);
  SmallParams prm;
  prm.src = (const uint32_t*)d_rgba;
  prm.frame_px = (unsigned long long)w * h;
  prm.sw = w;
  prm.sh = h;
  prm.dw = iw;
  prm.dh = ih;
  prm.shrink = (iw != w || ih != h) ? 1 : 0;
  prm.ppc = plan.ppc;
  prm.seed = (unsigned int)seed;
  prm.k = k;
  prm.max_iter = o.max_iter ? o.max_iter : 1;
  prm.check_every = o.check_every;
  prm.conv_threshold = o.convergence >= 0.0f ? o.convergence : (cs == KMG_LAB ? 1.0f : 0.01f);  // lib.rs:189-194
  prm.color_space = cs;
  prm.tail = tail;
  prm.blob_stride = blob_stride;
  prm.lut = ctx->d_lut;
  prm.gscratch = plan.throughput ? (unsigned char*)scratch : nullptr;
  prm.n_frames = n_frames;
  if (plan.throughput && !scratch) return fail(KMG_ERR_BAD_ARG, "throughput plan without a scratch buffer");
  return launch_small(ctx, plan, prm, job->P, n_frames, s);
}

// Returns with the job's final state on its way into ws->h_state (read it after the caller's
// next stream synchronisation).  *prepared: the blob already holds table + palette for the remap.
// What k_kmeans_small leaves in the blob besides the centroids (SmallParams::tail) for a remap mode.
static int tail_for_mode(int mode) { return mode == KMG_DITHER ? 2 : (mode == KMG_REPLACE ? 1 : 0); }

// Upload of a host image into ws->in.  Large images that will be clustered at full size by the
// staged launches (no shrink, too large for the fused kernel) are uploaded in 16 MiB bands on a
// second stream, and every band is converted to the work plane as soon as it has arrived, so the
// conversion hides behind the upload (*plane_ready: ws->work already holds the plane).
static int upload_image(kmg_ctx* ctx, Workspace* ws, const uint8_t* rgba, uint32_t w, uint32_t h, uint32_t k, int cs,
                        const kmg_opts& o, bool* plane_ready) {
  const size_t n = (size_t)w * h, bytes = n * 4;
  *plane_ready = false;
  TRY(ws->in.ensure(bytes));
  const bool shrink = o.max_dim != 0 && (w > o.max_dim || h > o.max_dim);
  SmallPlan plan;
  const bool fused = !(o.flags & KMG_OPT_NO_FUSED_KMEANS) && small_plan(ctx, n, k, 1, &plan);
  if (shrink || fused || bytes < ((size_t)32 << 20)) {
    CU(cudaMemcpyAsync(ws->in.p, rgba, bytes, cudaMemcpyHostToDevice, ws->stream));
    return KMG_OK;
  }
  if (!ws->copy_stream) CU(cudaStreamCreateWithFlags(&ws->copy_stream, cudaStreamNonBlocking));
  TRY(ws->work.ensure(n * 16));
  const size_t band_px = ((size_t)16 << 20) / 4;
  const size_t n_bands = (n + band_px - 1) / band_px;
  while (ws->band_done.size() < n_bands) {
    cudaEvent_t e;
    CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    ws->band_done.push_back(e);
  }
  for (size_t b = 0; b < n_bands; ++b) {
    const size_t p0 = b * band_px, np = std::min(band_px, n - p0);
    CU(cudaMemcpyAsync((uint8_t*)ws->in.p + p0 * 4, rgba + p0 * 4, np * 4, cudaMemcpyHostToDevice, ws->copy_stream));
    CU(cudaEventRecord(ws->band_done[b], ws->copy_stream));
    CU(cudaStreamWaitEvent(ws->stream, ws->band_done[b], 0));
    TRY(launch_convert(ctx, (const uint8_t*)ws->in.p + p0 * 4, np, cs, (float*)ws->work.p + p0 * 4, ws->stream));
  }
  *plane_ready = true;
  return KMG_OK;
}

static int kmeans_on_device(kmg_ctx* ctx, Workspace* ws, const uint8_t* d_rgba, uint32_t w, uint32_t h, uint32_t k,
                            int cs, const kmg_opts& o, kmg_job* job, bool* prepared, int tail = 0,
                            bool plane_ready = false) {
  cudaStream_t s = ws->stream;
  uint32_t iw = w, ih = h;
  const bool shrink = o.max_dim != 0 && (w > o.max_dim || h > o.max_dim);  // structures.rs:67-74
  if (shrink) kmg_resized_dims(w, h, o.max_dim, &iw, &ih);
  const size_t n = (size_t)iw * ih;
  TRY(ws->blob.ensure(job_blob_bytes(k)));
  SmallPlan plan;
  if (!(o.flags & KMG_OPT_NO_FUSED_KMEANS) && small_plan(ctx, n, k, 1, &plan)) {
    job->h_state = ws->h_state;
    TRY(kmeans_small_on_device(ctx, plan, d_rgba, 1, w, h, iw, ih, k, cs, o, ws->blob.p, 0, tail, nullptr, job, s));
    CU(cudaMemcpyAsync(ws->h_state, job->P.st, sizeof(JobState), cudaMemcpyDeviceToHost, s));
    if (prepared) *prepared = true;
    return KMG_OK;
  }
  const uint8_t* img = d_rgba;
  if (shrink) {
    TRY(ws->small.ensure((size_t)iw * ih * 4));
    TRY(launch_resize(ctx, d_rgba, w, h, (uint8_t*)ws->small.p, iw, ih, s));
    img = (const uint8_t*)ws->small.p;
  }
  TRY(ws->work.ensure(n * 16));
  TRY(ws->dmin.ensure(dmin_bytes(n)));
  // the first consumer of the plane is the first init round: the conversion is fused into it (k >= 2)
  const bool fuse_convert = !plane_ready && k >= 2;
  if (!plane_ready && !fuse_convert) TRY(launch_convert(ctx, img, n, cs, (float*)ws->work.p, s));
  TRY(job_setup(job, ctx, (const float*)ws->work.p, iw, ih, k, cs, o, ws->blob.p, (float*)ws->dmin.p, ws->h_state, s));
  job->rgba_src = fuse_convert ? img : nullptr;
  TRY(job_init_impl(job, nullptr, nullptr, s));
  TRY(job_run_impl(job, nullptr, s));
  if (prepared) *prepared = false;
  return KMG_OK;
}

static int check_image_args(const char* fn, kmg_ctx* ctx, const uint8_t* rgba, uint32_t w, uint32_t h, uint32_t k,
                            int cs) {
  if (!ctx) return fail(KMG_ERR_BAD_ARG, "%s: ctx is NULL", fn);
  if (!rgba) return fail(KMG_ERR_BAD_ARG, "%s: image pointer is NULL", fn);
  if (cs != KMG_LAB && cs != KMG_RGB) return fail(KMG_ERR_BAD_ARG, "%s: unknown colour space %d", fn, cs);
  return validate_dims(w, h, k);
}

extern "C" int kmg_kmeans_palette(kmg_ctx* ctx, const uint8_t* rgba, uint32_t w, uint32_t h, uint32_t k, int cs,
                                  const kmg_opts* opts, float* centroids_out, uint32_t* passes_out) {
  TRY(check_image_args("kmg_kmeans_palette", ctx, rgba, w, h, k, cs));
  if (!centroids_out) return fail(KMG_ERR_BAD_ARG, "kmg_kmeans_palette: centroids_out is NULL");
  CU(cudaSetDevice(ctx->device));
  Workspace* ws = ws_acquire(ctx);
  if (!ws) return fail(KMG_ERR_CUDA, "could not create a workspace");
  WsGuard guard{ctx, ws};
  const kmg_opts o = resolve_opts(opts);
  bool plane_ready = false;
  TRY(upload_image(ctx, ws, rgba, w, h, k, cs, o, &plane_ready));
  kmg_job job;
  TRY(kmeans_on_device(ctx, ws, (const uint8_t*)ws->in.p, w, h, k, cs, o, &job, nullptr, 0, plane_ready));
  CU(cudaMemcpyAsync(centroids_out, job.P.cent, (size_t)k * 16, cudaMemcpyDeviceToHost, ws->stream));
  CU(cudaStreamSynchronize(ws->stream));
  if (passes_out) *passes_out = ws->h_state->passes;
  return KMG_OK;
}

// Remap of one large host image as a pipeline of row bands over three workspaces (stream + buffers
// each): the upload of band i+1 and the read-back of band i-1 overlap the kernel of band i, so the
// call costs about one direction of PCIe traffic instead of upload + kernel + read-back in series.
// Bands start on rows that are multiples of 4, which keeps the 4x4 dither matrix aligned
// (mix_colors.wgsl:21-27 indexes it with x % 4 + 4 (y % 4)); every pixel is independent otherwise.
// Overlap needs page-locked caller buffers; pageable ones still work, the copies then serialise.
static int remap_banded(kmg_ctx* ctx, const uint8_t* rgba, uint32_t w, uint32_t h, const float* centroids, uint32_t k,
                        int cs, int mode, uint8_t* out_rgba) {
  const size_t row_bytes = (size_t)w * 4;
  const size_t total = row_bytes * h;
  size_t band_bytes = std::min<size_t>((size_t)16 << 20, std::max<size_t>((size_t)4 << 20, total / 8));
  uint32_t band_rows = (uint32_t)std::max<size_t>(4, (band_bytes / row_bytes) & ~(size_t)3);
  const uint32_t n_bands = (h + band_rows - 1) / band_rows;
  Workspace* wss[3] = {nullptr, nullptr, nullptr};
  cudaEvent_t prepared = nullptr;
  int rc = KMG_OK;
  for (int i = 0; i < 3 && rc == KMG_OK; ++i) {
    wss[i] = ws_acquire(ctx);
    if (!wss[i]) rc = fail(KMG_ERR_CUDA, "could not create a workspace");
  }
  auto run = [&]() -> int {
    CU(cudaEventCreateWithFlags(&prepared, cudaEventDisableTiming));
    TRY(wss[0]->blob.ensure(job_blob_bytes(k)));
    kmg_job job;
    kmg_opts o;
    kmg_default_opts(&o);
    TRY(job_setup(&job, ctx, nullptr, w, h, k, cs, o, wss[0]->blob.p, nullptr, wss[0]->h_state, wss[0]->stream));
    CU(cudaMemcpyAsync(job.P.cent, centroids, (size_t)k * 16, cudaMemcpyHostToDevice, wss[0]->stream));
    TRY(launch_prepare(&job, true, wss[0]->stream));
    CU(cudaEventRecord(prepared, wss[0]->stream));
    CU(cudaStreamWaitEvent(wss[1]->stream, prepared, 0));
    CU(cudaStreamWaitEvent(wss[2]->stream, prepared, 0));
    for (uint32_t b = 0; b < n_bands; ++b) {
      Workspace* ws = wss[b % 3];
      const uint32_t r0 = b * band_rows, rows = std::min(band_rows, h - r0);
      const size_t bytes = row_bytes * rows, off = row_bytes * r0;
      TRY(ws->in.ensure(row_bytes * band_rows));
      TRY(ws->out.ensure(row_bytes * band_rows));
      CU(cudaMemcpyAsync(ws->in.p, rgba + off, bytes, cudaMemcpyHostToDevice, ws->stream));
      TRY(launch_remap(&job, (const uint8_t*)ws->in.p, w, rows, mode, (uint8_t*)ws->out.p, ws->stream, true));
      CU(cudaMemcpyAsync(out_rgba + off, ws->out.p, bytes, cudaMemcpyDeviceToHost, ws->stream));
    }
    for (int i = 0; i < 3; ++i) CU(cudaStreamSynchronize(wss[i]->stream));
    return KMG_OK;
  };
  if (rc == KMG_OK) rc = run();
  for (int i = 0; i < 3; ++i)
    if (wss[i]) {
      if (rc != KMG_OK) cudaStreamSynchronize(wss[i]->stream);
      ws_release(ctx, wss[i]);
    }
  if (prepared) cudaEventDestroy(prepared);
  return rc;
}

extern "C" int kmg_remap(kmg_ctx* ctx, const uint8_t* rgba, uint32_t w, uint32_t h, const float* centroids, uint32_t k,
                         int cs, int mode, uint8_t* out_rgba) {
  TRY(check_image_args("kmg_remap", ctx, rgba, w, h, k, cs));
  if (!centroids || !out_rgba) return fail(KMG_ERR_BAD_ARG, "kmg_remap: NULL argument");
  if (mode < KMG_REPLACE || mode > KMG_MELD) return fail(KMG_ERR_BAD_ARG, "kmg_remap: unknown mode %d", mode);
  CU(cudaSetDevice(ctx->device));
  if ((size_t)w * h * 4 >= ((size_t)12 << 20) && h >= 12) return remap_banded(ctx, rgba, w, h, centroids, k, cs, mode, out_rgba);
  Workspace* ws = ws_acquire(ctx);
  if (!ws) return fail(KMG_ERR_CUDA, "could not create a workspace");
  WsGuard guard{ctx, ws};
  cudaStream_t s = ws->stream;
  const size_t bytes = (size_t)w * h * 4;
  TRY(ws->in.ensure(bytes));
  TRY(ws->out.ensure(bytes));
  TRY(ws->blob.ensure(job_blob_bytes(k)));
  CU(cudaMemcpyAsync(ws->in.p, rgba, bytes, cudaMemcpyHostToDevice, s));
  kmg_job job;
  kmg_opts o;
  kmg_default_opts(&o);
  TRY(job_setup(&job, ctx, nullptr, w, h, k, cs, o, ws->blob.p, nullptr, ws->h_state, s));
  CU(cudaMemcpyAsync(job.P.cent, centroids, (size_t)k * 16, cudaMemcpyHostToDevice, s));
  TRY(launch_remap(&job, (const uint8_t*)ws->in.p, w, h, mode, (uint8_t*)ws->out.p, s));
  CU(cudaMemcpyAsync(out_rgba, ws->out.p, bytes, cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));
  return KMG_OK;
}

// reduce() of a large host image that is shrunk before clustering (the reference's default,
// core/src/structures.rs:67-74), as one pipeline.  The k-means only reads the source rows the
// bilinear taps touch (resize.wgsl:5-19: at most 2 per clustered row, a few per cent of a large
// image), so those rows are uploaded first and clustered; only then does the whole image stream
// through: upload of band i+1 (copy stream), remap of band i (main stream) and read-back of band
// i-1 (output stream) overlap, and the call costs about one direction of PCIe traffic instead of
// upload + kernels + read-back in series.  The full upload starts after the k-means so that no
// copy ever writes rows a running kernel reads.
static int reduce_pipelined(kmg_ctx* ctx, Workspace* ws, const uint8_t* rgba, uint32_t w, uint32_t h, uint32_t k, int cs,
                            int mode, const kmg_opts& o, const std::vector<uint8_t>& tap_rows, uint8_t* out_rgba,
                            float* centroids_out, uint32_t* passes_out) {
  cudaStream_t s = ws->stream;
  const size_t row_bytes = (size_t)w * 4, bytes = row_bytes * h;
  TRY(ws->in.ensure(bytes));
  TRY(ws->out.ensure(bytes));
  if (!ws->copy_stream) CU(cudaStreamCreateWithFlags(&ws->copy_stream, cudaStreamNonBlocking));
  if (!ws->out_stream) CU(cudaStreamCreateWithFlags(&ws->out_stream, cudaStreamNonBlocking));
  if (!ws->palette_ready) CU(cudaEventCreateWithFlags(&ws->palette_ready, cudaEventDisableTiming));
  // 1. the rows the shrink reads, as runs of consecutive rows
  for (uint32_t r = 0; r < h;) {
    if (!tap_rows[r]) {
      ++r;
      continue;
    }
    uint32_t e = r;
    while (e < h && tap_rows[e]) ++e;
    CU(cudaMemcpyAsync((uint8_t*)ws->in.p + row_bytes * r, rgba + row_bytes * r, row_bytes * (e - r), cudaMemcpyHostToDevice, s));
    r = e;
  }
  // 2. k-means on the shrunk image (table, palette and dither threshold end up in the job blob)
  kmg_job job;
  bool prepared = false;
  TRY(kmeans_on_device(ctx, ws, (const uint8_t*)ws->in.p, w, h, k, cs, o, &job, &prepared, tail_for_mode(mode)));
  if (!prepared) TRY(launch_prepare(&job, true, s));
  CU(cudaEventRecord(ws->palette_ready, s));
  CU(cudaStreamWaitEvent(ws->copy_stream, ws->palette_ready, 0));
  // 3. the image, band by band
  const uint32_t band_rows = (uint32_t)std::max<size_t>(4, (((size_t)16 << 20) / row_bytes) & ~(size_t)3);
  const uint32_t n_bands = (h + band_rows - 1) / band_rows;
  while (ws->band_done.size() < n_bands || ws->band_out.size() < n_bands) {
    cudaEvent_t e;
    CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    (ws->band_done.size() < n_bands ? ws->band_done : ws->band_out).push_back(e);
  }
  for (uint32_t b = 0; b < n_bands; ++b) {
    const uint32_t r0 = b * band_rows, rows = std::min(band_rows, h - r0);
    const size_t off = row_bytes * r0, sz = row_bytes * rows;
    CU(cudaMemcpyAsync((uint8_t*)ws->in.p + off, rgba + off, sz, cudaMemcpyHostToDevice, ws->copy_stream));
    CU(cudaEventRecord(ws->band_done[b], ws->copy_stream));
    CU(cudaStreamWaitEvent(s, ws->band_done[b], 0));
    TRY(launch_remap(&job, (const uint8_t*)ws->in.p + off, w, rows, mode, (uint8_t*)ws->out.p + off, s, true));
    CU(cudaEventRecord(ws->band_out[b], s));
    CU(cudaStreamWaitEvent(ws->out_stream, ws->band_out[b], 0));
    CU(cudaMemcpyAsync(out_rgba + off, (const uint8_t*)ws->out.p + off, sz, cudaMemcpyDeviceToHost, ws->out_stream));
  }
  if (centroids_out) CU(cudaMemcpyAsync(centroids_out, job.P.cent, (size_t)k * 16, cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));
  CU(cudaStreamSynchronize(ws->out_stream));
  CU(cudaStreamSynchronize(ws->copy_stream));
  if (passes_out) *passes_out = ws->h_state->passes;
  return KMG_OK;
}

// Source rows the bilinear shrink of a w x h image to iw x ih reads (resize_taps in
// kmg_kernels.cuh, same f32 arithmetic), with one row of margin on either side.
static size_t shrink_tap_rows(uint32_t h, uint32_t ih, std::vector<uint8_t>* rows) {
  rows->assign(h, 0);
  for (uint32_t gy = 0; gy < ih; ++gy) {
    const float py = ((float)gy / (float)ih) * (float)h - 0.5f;
    const long long y0 = (long long)std::floor(py);
    for (long long r = y0 - 1; r <= y0 + 2; ++r) (*rows)[(size_t)std::min<long long>(std::max<long long>(r, 0), (long long)h - 1)] = 1;
  }
  size_t count = 0;
  for (uint8_t v : *rows) count += v;
  return count;
}

extern "C" int kmg_reduce(kmg_ctx* ctx, const uint8_t* rgba, uint32_t w, uint32_t h, uint32_t k, int cs, int mode,
                          const kmg_opts* opts, uint8_t* out_rgba, float* centroids_out, uint32_t* passes_out) {
  TRY(check_image_args("kmg_reduce", ctx, rgba, w, h, k, cs));
  if (!out_rgba) return fail(KMG_ERR_BAD_ARG, "kmg_reduce: out_rgba is NULL");
  if (mode < KMG_REPLACE || mode > KMG_MELD) return fail(KMG_ERR_BAD_ARG, "kmg_reduce: unknown mode %d", mode);
  CU(cudaSetDevice(ctx->device));
  Workspace* ws = ws_acquire(ctx);
  if (!ws) return fail(KMG_ERR_CUDA, "could not create a workspace");
  WsGuard guard{ctx, ws};
  cudaStream_t s = ws->stream;
  const size_t bytes = (size_t)w * h * 4;
  {
    const kmg_opts po = resolve_opts(opts);
    if (po.max_dim != 0 && (w > po.max_dim || h > po.max_dim) && bytes >= ((size_t)32 << 20) &&
        !getenv("KMG_NO_REDUCE_PIPELINE")) {  // the switch exists for A/B timing
      uint32_t iw, ih;
      kmg_resized_dims(w, h, po.max_dim, &iw, &ih);
      std::vector<uint8_t> tap_rows;
      if (shrink_tap_rows(h, ih, &tap_rows) * 2 <= h)  // worth it when the taps (plus margin) touch at most half of the rows
        return reduce_pipelined(ctx, ws, rgba, w, h, k, cs, mode, po, tap_rows, out_rgba, centroids_out, passes_out);
    }
  }
  TRY(ws->out.ensure(bytes));
  const kmg_opts o = resolve_opts(opts);
  bool plane_ready = false;
  TRY(upload_image(ctx, ws, rgba, w, h, k, cs, o, &plane_ready));
  kmg_job job;
  bool prepared = false;
  TRY(kmeans_on_device(ctx, ws, (const uint8_t*)ws->in.p, w, h, k, cs, o, &job, &prepared, tail_for_mode(mode), plane_ready));
  TRY(launch_remap(&job, (const uint8_t*)ws->in.p, w, h, mode, (uint8_t*)ws->out.p, s, prepared));
  CU(cudaMemcpyAsync(out_rgba, ws->out.p, bytes, cudaMemcpyDeviceToHost, s));
  if (centroids_out) CU(cudaMemcpyAsync(centroids_out, job.P.cent, (size_t)k * 16, cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));
  if (passes_out) *passes_out = ws->h_state->passes;
  return KMG_OK;
}

extern "C" int kmg_resize(kmg_ctx* ctx, const uint8_t* rgba, uint32_t w, uint32_t h, uint32_t max_size, uint8_t* out) {
  TRY(check_image_args("kmg_resize", ctx, rgba, w, h, 1, KMG_LAB));
  if (!out || max_size == 0) return fail(KMG_ERR_BAD_ARG, "kmg_resize: bad argument");
  CU(cudaSetDevice(ctx->device));
  Workspace* ws = ws_acquire(ctx);
  if (!ws) return fail(KMG_ERR_CUDA, "could not create a workspace");
  WsGuard guard{ctx, ws};
  uint32_t dw, dh;
  kmg_resized_dims(w, h, max_size, &dw, &dh);
  const size_t bytes = (size_t)w * h * 4, obytes = (size_t)dw * dh * 4;
  TRY(ws->in.ensure(bytes));
  TRY(ws->small.ensure(obytes));
  CU(cudaMemcpyAsync(ws->in.p, rgba, bytes, cudaMemcpyHostToDevice, ws->stream));
  TRY(launch_resize(ctx, (const uint8_t*)ws->in.p, w, h, (uint8_t*)ws->small.p, dw, dh, ws->stream));
  CU(cudaMemcpyAsync(out, ws->small.p, obytes, cudaMemcpyDeviceToHost, ws->stream));
  CU(cudaStreamSynchronize(ws->stream));
  return KMG_OK;
}

// ------------------------------------------------------------------------------------------------
// batches of frames (BASELINE config 5)

static size_t batch_blob_stride(uint32_t k) { return (job_blob_bytes(k) + 255) & ~(size_t)255; }

// Clustered size of a w x h frame under opts, and whether the fused kernel can take a batch of them.
static bool batch_plan(kmg_ctx* ctx, uint32_t w, uint32_t h, uint32_t k, uint32_t n_frames, const kmg_opts& o,
                       uint32_t* iw, uint32_t* ih, SmallPlan* plan) {
  *iw = w;
  *ih = h;
  if (o.max_dim != 0 && (w > o.max_dim || h > o.max_dim)) kmg_resized_dims(w, h, o.max_dim, iw, ih);
  if (o.flags & KMG_OPT_NO_FUSED_KMEANS) return false;
  return small_plan(ctx, (unsigned long long)*iw * *ih, k, n_frames, plan);
}

// Fused batch on device buffers: ONE cluster launch for the k-means of all frames, ONE remap launch
// (frames on gridDim.y), strided read-back of centroids / pass counts.  Asynchronous on s.
static int reduce_batch_fused(kmg_ctx* ctx, const SmallPlan& plan, const uint8_t* d_rgba, uint32_t n_frames, uint32_t w,
                              uint32_t h, uint32_t iw, uint32_t ih, uint32_t k, int cs, int mode, const kmg_opts& o,
                              uint8_t* d_out, void* blobs, void* scratch, float* centroids_out, uint32_t* passes_out,
                              cudaStream_t s) {
  const size_t stride = batch_blob_stride(k);
  kmg_job job;
  TRY(kmeans_small_on_device(ctx, plan, d_rgba, n_frames, w, h, iw, ih, k, cs, o, blobs, stride, tail_for_mode(mode), scratch,
                             &job, s));
  TRY(launch_remap(&job, d_rgba, w, h, mode, d_out, s, true, n_frames, stride));
  if (centroids_out)
    CU(cudaMemcpy2DAsync(centroids_out, (size_t)k * 16, job.P.cent, stride, (size_t)k * 16, n_frames, cudaMemcpyDeviceToHost, s));
  if (passes_out)
    CU(cudaMemcpy2DAsync(passes_out, 4, &job.P.st->passes, stride, 4, n_frames, cudaMemcpyDeviceToHost, s));
  return KMG_OK;
}

extern "C" int kmg_dev_reduce_batch(kmg_ctx* ctx, const uint8_t* d_rgba, uint32_t n_frames, uint32_t w, uint32_t h,
                                    uint32_t k, int cs, int mode, const kmg_opts* opts, uint8_t* d_out,
                                    float* centroids_out, uint32_t* passes_out, void* stream) {
  TRY(check_image_args("kmg_dev_reduce_batch", ctx, d_rgba, w, h, k, cs));
  if (!d_out || n_frames == 0) return fail(KMG_ERR_BAD_ARG, "kmg_dev_reduce_batch: bad argument");
  if (mode < KMG_REPLACE || mode > KMG_MELD) return fail(KMG_ERR_BAD_ARG, "kmg_dev_reduce_batch: unknown mode %d", mode);
  CU(cudaSetDevice(ctx->device));
  Workspace* ws = ws_acquire(ctx);
  if (!ws) return fail(KMG_ERR_CUDA, "could not create a workspace");
  WsGuard guard{ctx, ws};
  const kmg_opts o = resolve_opts(opts);
  const size_t frame_bytes = (size_t)w * h * 4;
  uint32_t iw, ih;
  SmallPlan plan;
  if (batch_plan(ctx, w, h, k, n_frames, o, &iw, &ih, &plan)) {
    // the frames were produced on the caller's stream: stay on it
    cudaStream_t s = pick_stream(ctx, stream);
    TRY(ws->blob.ensure((size_t)n_frames * batch_blob_stride(k)));
    if (plan.throughput) TRY(ws->work.ensure(plan.scratch));
    TRY(reduce_batch_fused(ctx, plan, d_rgba, n_frames, w, h, iw, ih, k, cs, mode, o, d_out, ws->blob.p, ws->work.p,
                           centroids_out, passes_out, s));
    CU(cudaStreamSynchronize(s));
    return KMG_OK;
  }
  // frame by frame through the staged path (k > 32, or frames too large for a cluster)
  CU(cudaStreamSynchronize(pick_stream(ctx, stream)));
  for (uint32_t f = 0; f < n_frames; ++f) {
    kmg_job job;
    bool prepared = false;
    const uint8_t* in = d_rgba + (size_t)f * frame_bytes;
    TRY(kmeans_on_device(ctx, ws, in, w, h, k, cs, o, &job, &prepared, tail_for_mode(mode)));
    TRY(launch_remap(&job, in, w, h, mode, d_out + (size_t)f * frame_bytes, ws->stream, prepared));
    if (centroids_out)
      CU(cudaMemcpyAsync(centroids_out + (size_t)f * k * 4, job.P.cent, (size_t)k * 16, cudaMemcpyDeviceToHost, ws->stream));
    // the workspace blob is reused by the next frame: drain before overwriting it
    CU(cudaStreamSynchronize(ws->stream));
    if (passes_out) passes_out[f] = ws->h_state->passes;
  }
  return KMG_OK;
}

extern "C" int kmg_reduce_batch(kmg_ctx* ctx, const uint8_t* rgba, uint32_t n_frames, uint32_t w, uint32_t h,
                                uint32_t k, int cs, int mode, const kmg_opts* opts, uint8_t* out_rgba,
                                float* centroids_out, uint32_t* passes_out) {
  TRY(check_image_args("kmg_reduce_batch", ctx, rgba, w, h, k, cs));
  if (!out_rgba || n_frames == 0) return fail(KMG_ERR_BAD_ARG, "kmg_reduce_batch: bad argument");
  if (mode < KMG_REPLACE || mode > KMG_MELD) return fail(KMG_ERR_BAD_ARG, "kmg_reduce_batch: unknown mode %d", mode);
  CU(cudaSetDevice(ctx->device));
  const kmg_opts o = resolve_opts(opts);
  const size_t frame_bytes = (size_t)w * h * 4;
  uint32_t iw, ih;
  SmallPlan plan;
  if (n_frames == 1 || !batch_plan(ctx, w, h, k, n_frames, o, &iw, &ih, &plan)) {
    for (uint32_t f = 0; f < n_frames; ++f) {
      const size_t off = (size_t)f * frame_bytes;
      TRY(kmg_reduce(ctx, rgba + off, w, h, k, cs, mode, opts, out_rgba + off,
                     centroids_out ? centroids_out + (size_t)f * k * 4 : nullptr, passes_out ? passes_out + f : nullptr));
    }
    return KMG_OK;
  }
  // Chunks of frames flow through up to three workspaces (stream + buffers each): the upload of
  // chunk i+1 and the read-back of chunk i-1 overlap the kernels of chunk i.  Overlap needs the
  // caller's buffers to be page-locked; pageable buffers still work, the copies then serialise.
  // ~64 MiB per chunk: large enough for full-rate DMA, small enough that filling and draining the
  // pipeline costs little; at least 8 frames so the cluster launch still covers the SMs
  const uint32_t chunk = (uint32_t)std::min<size_t>(n_frames, std::max<size_t>(8, ((size_t)64 << 20) / frame_bytes));
  const uint32_t n_chunks = (n_frames + chunk - 1) / chunk;
  const uint32_t n_ws = std::min<uint32_t>(3, n_chunks);
  Workspace* wss[3] = {nullptr, nullptr, nullptr};
  int rc = KMG_OK;
  for (uint32_t i = 0; i < n_ws && rc == KMG_OK; ++i) {
    wss[i] = ws_acquire(ctx);
    if (!wss[i]) rc = fail(KMG_ERR_CUDA, "could not create a workspace");
  }
  auto run = [&]() -> int {
    const size_t stride = batch_blob_stride(k);
    // The per-frame palettes and pass counts are staged in page-locked memory: read straight into
    // the caller's (usually pageable) arrays, each chunk's read-back would block the host until
    // the chunk's kernels had finished, and the next upload could not be queued behind them.
    const size_t cent_bytes = centroids_out ? (size_t)n_frames * k * 16 : 0;
    const size_t pass_bytes = passes_out ? (size_t)n_frames * 4 : 0;
    float* h_cent = nullptr;
    uint32_t* h_pass = nullptr;
    if (cent_bytes + pass_bytes) {
      TRY(wss[0]->stage.ensure(cent_bytes + pass_bytes));
      if (cent_bytes) h_cent = (float*)wss[0]->stage.p;
      if (pass_bytes) h_pass = (uint32_t*)((char*)wss[0]->stage.p + cent_bytes);
    }
    for (uint32_t c = 0; c < n_chunks; ++c) {
      Workspace* ws = wss[c % n_ws];
      const uint32_t f0 = c * chunk, nf = std::min(chunk, n_frames - f0);
      TRY(ws->in.ensure((size_t)nf * frame_bytes));
      TRY(ws->out.ensure((size_t)nf * frame_bytes));
      TRY(ws->blob.ensure((size_t)nf * stride));
      SmallPlan cplan;  // a chunk is a small batch: clusters unless it is large enough for the throughput mode
      if (!small_plan(ctx, (unsigned long long)iw * ih, k, nf, &cplan)) return fail(KMG_ERR_CUDA, "no launch plan for a chunk");
      if (cplan.throughput) TRY(ws->work.ensure(cplan.scratch));
      cudaStream_t s = ws->stream;
      CU(cudaMemcpyAsync(ws->in.p, rgba + (size_t)f0 * frame_bytes, (size_t)nf * frame_bytes, cudaMemcpyHostToDevice, s));
      TRY(reduce_batch_fused(ctx, cplan, (const uint8_t*)ws->in.p, nf, w, h, iw, ih, k, cs, mode, o, (uint8_t*)ws->out.p,
                             ws->blob.p, ws->work.p, h_cent ? h_cent + (size_t)f0 * k * 4 : nullptr,
                             h_pass ? h_pass + f0 : nullptr, s));
      CU(cudaMemcpyAsync(out_rgba + (size_t)f0 * frame_bytes, ws->out.p, (size_t)nf * frame_bytes, cudaMemcpyDeviceToHost, s));
    }
    for (uint32_t i = 0; i < n_ws; ++i) CU(cudaStreamSynchronize(wss[i]->stream));
    if (h_cent) memcpy(centroids_out, h_cent, cent_bytes);
    if (h_pass) memcpy(passes_out, h_pass, pass_bytes);
    return KMG_OK;
  };
  if (rc == KMG_OK) rc = run();
  for (uint32_t i = 0; i < 3; ++i)
    if (wss[i]) {
      if (rc != KMG_OK) cudaStreamSynchronize(wss[i]->stream);
      ws_release(ctx, wss[i]);
    }
  return rc;
}

// ------------------------------------------------------------------------------------------------
// multi-GPU communicator

#if KMG_HAVE_NCCL_HEADER
// Peer mailboxes for the in-kernel exchange: every rank exports its mailbox through CUDA IPC, the
// 64-byte handles travel over the freshly created NCCL communicator, and every rank maps all the
// others.  Any failure (no peer access, IPC unavailable in this container, > 8 ranks) leaves
// p2p == false on ALL ranks and the per-pass NCCL all-reduce stays in use.
static void p2p_teardown(kmg_ctx* ctx) {
  for (unsigned int r = 0; r < MAX_PEERS; ++r) {
    if (ctx->mbox_peer[r] && ctx->mbox_peer[r] != ctx->mbox_own) cudaIpcCloseMemHandle(ctx->mbox_peer[r]);
    ctx->mbox_peer[r] = nullptr;
  }
  if (ctx->mbox_own) cudaFree(ctx->mbox_own);
  ctx->mbox_own = nullptr;
  ctx->p2p = false;
  cudaGetLastError();
}

static int p2p_setup(kmg_ctx* ctx) {
  ctx->p2p = false;
  if (ctx->n_ranks < 2) return KMG_OK;
  cudaStream_t s = ctx->stream;
  const int n = ctx->n_ranks;
  int ok = (n <= (int)MAX_PEERS && !getenv("KMG_NO_P2P")) ? 1 : 0;
  cudaIpcMemHandle_t mine;
  memset(&mine, 0, sizeof(mine));
  if (ok && cudaMalloc(&ctx->mbox_own, MBOX_BYTES) != cudaSuccess) ok = 0;
  if (ok && cudaMemsetAsync(ctx->mbox_own, 0, MBOX_BYTES, s) != cudaSuccess) ok = 0;
  if (ok && cudaIpcGetMemHandle(&mine, ctx->mbox_own) != cudaSuccess) ok = 0;
  cudaGetLastError();
  // all-gather {ok, handle} (device staging buffers; 128 bytes per rank)
  const size_t rec = 128;
  unsigned char *d_send = nullptr, *d_recv = nullptr;
  CU(cudaMalloc((void**)&d_send, rec));
  CU(cudaMalloc((void**)&d_recv, rec * n));
  std::vector<unsigned char> h_send(rec, 0), h_recv(rec * n, 0);
  h_send[0] = (unsigned char)ok;
  static_assert(sizeof(cudaIpcMemHandle_t) <= 64, "IPC handle fits the record");
  memcpy(h_send.data() + 64, &mine, sizeof(mine));
  CU(cudaMemcpyAsync(d_send, h_send.data(), rec, cudaMemcpyHostToDevice, s));
  // one all-reduce-free gather: rank r contributes its record at offset r (sum of zero-padded byte vectors)
  CU(cudaMemsetAsync(d_recv, 0, rec * n, s));
  CU(cudaMemcpyAsync(d_recv + rec * ctx->rank, d_send, rec, cudaMemcpyDeviceToDevice, s));
  NC(nccl_api().AllReduce(d_recv, d_recv, rec * n, ncclUint8, ncclSum, ctx->comm, s));
  CU(cudaMemcpyAsync(h_recv.data(), d_recv, rec * n, cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));
  int all_ok = 1;
  for (int r = 0; r < n; ++r) all_ok &= h_recv[rec * r] ? 1 : 0;
  int mapped = all_ok;
  if (all_ok) {
    for (int r = 0; r < n && mapped; ++r) {
      if (r == ctx->rank) {
        ctx->mbox_peer[r] = ctx->mbox_own;
        continue;
      }
      cudaIpcMemHandle_t h;
      memcpy(&h, h_recv.data() + rec * r + 64, sizeof(h));
      if (cudaIpcOpenMemHandle(&ctx->mbox_peer[r], h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        cudaGetLastError();
        ctx->mbox_peer[r] = nullptr;
        mapped = 0;
      }
    }
  }
  // second round: did every rank map every mailbox?
  h_send.assign(rec, 0);
  h_send[0] = (unsigned char)mapped;
  CU(cudaMemcpyAsync(d_send, h_send.data(), rec, cudaMemcpyHostToDevice, s));
  CU(cudaMemsetAsync(d_recv, 0, rec * n, s));
  CU(cudaMemcpyAsync(d_recv + rec * ctx->rank, d_send, rec, cudaMemcpyDeviceToDevice, s));
  NC(nccl_api().AllReduce(d_recv, d_recv, rec * n, ncclUint8, ncclSum, ctx->comm, s));
  CU(cudaMemcpyAsync(h_recv.data(), d_recv, rec * n, cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));
  cudaFree(d_send);
  cudaFree(d_recv);
  int all_mapped = 1;
  for (int r = 0; r < n; ++r) all_mapped &= h_recv[rec * r] ? 1 : 0;
  if (all_mapped)
    ctx->p2p = true;
  else
    p2p_teardown(ctx);
  return KMG_OK;
}
#endif

extern "C" int kmg_comm_unique_id(kmg_ctx* ctx, uint8_t id_out[128]) {
#if KMG_HAVE_NCCL_HEADER
  if (!ctx || !id_out) return fail(KMG_ERR_BAD_ARG, "kmg_comm_unique_id: NULL argument");
  if (!nccl_api().ok) return fail(KMG_ERR_NCCL, "NCCL could not be loaded (libnccl.so.2)");
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  ncclUniqueId id;
  NC(nccl_api().GetUniqueId(&id));
  memcpy(id_out, &id, 128);
  return KMG_OK;
#else
  (void)ctx;
  (void)id_out;
  return fail(KMG_ERR_UNSUPPORTED, "built without nccl.h");
#endif
}

extern "C" int kmg_comm_init(kmg_ctx* ctx, const uint8_t id_in[128], int n_ranks, int rank) {
#if KMG_HAVE_NCCL_HEADER
  if (!ctx || !id_in) return fail(KMG_ERR_BAD_ARG, "kmg_comm_init: NULL argument");
  if (n_ranks < 1 || rank < 0 || rank >= n_ranks) return fail(KMG_ERR_BAD_ARG, "kmg_comm_init: bad rank %d of %d", rank, n_ranks);
  if (!nccl_api().ok) return fail(KMG_ERR_NCCL, "NCCL could not be loaded (libnccl.so.2)");
  CU(cudaSetDevice(ctx->device));
  if (ctx->comm) return fail(KMG_ERR_BAD_ARG, "kmg_comm_init: communicator already initialised");
  ncclUniqueId id;
  memcpy(&id, id_in, 128);
  NC(nccl_api().CommInitRank(&ctx->comm, n_ranks, id, rank));
  ctx->n_ranks = n_ranks;
  ctx->rank = rank;
  TRY(p2p_setup(ctx));
  return KMG_OK;
#else
  (void)ctx;
  (void)id_in;
  (void)n_ranks;
  (void)rank;
  return fail(KMG_ERR_UNSUPPORTED, "built without nccl.h");
#endif
}

extern "C" int kmg_comm_mode(kmg_ctx* ctx) {
  if (!ctx || ctx->n_ranks < 2) return 0;
  return ctx->p2p ? 2 : 1;
}

extern "C" int kmg_comm_destroy(kmg_ctx* ctx) {
#if KMG_HAVE_NCCL_HEADER
  if (!ctx) return fail(KMG_ERR_BAD_ARG, "kmg_comm_destroy: NULL ctx");
  if (ctx->comm) {
    CU(cudaSetDevice(ctx->device));
    CU(cudaDeviceSynchronize());
    p2p_teardown(ctx);
    NC(nccl_api().CommDestroy(ctx->comm));
    ctx->comm = nullptr;
  }
  ctx->n_ranks = 1;
  ctx->rank = 0;
  return KMG_OK;
#else
  (void)ctx;
  return KMG_OK;
#endif
}

#ifdef KMG_TRACE
// Development aid (make TRACE=1): phase timestamps of the last k_kmeans_small launch, 16 x 512 clock64 values.
extern "C" int kmg_debug_small_trace(uint64_t* out) {
  CU(cudaDeviceSynchronize());
  CU(cudaMemcpyFromSymbol(out, g_small_trace, sizeof(unsigned long long) * SMALL_MAX_CLUSTER * 512));
  return KMG_OK;
}
#endif
