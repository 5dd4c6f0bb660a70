// Host side of libdcol_b200.so: context, shape-table upload, class-pair binning
// (device counting sort), kernel dispatch over streams, host<->device staging.
// Entry points are declared in include/dcol_b200.h.
#include "../../include/dcol_b200.h"
#include "dcol_common.cuh"
#include "dcol_deriv.cuh"
#ifndef DCOL_MAXTHREADS
#define DCOL_MAXTHREADS 256
#endif
#include <cuda_runtime.h>
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace dcol {
// one translation unit per query kind (dcol_kernels_v{0,1,2}.cu, dcol_persistent.cuh)
const void* persistent_kernel_ptr_v0(); const void* persistent_kernel_ptr_v1(); const void* persistent_kernel_ptr_v2();
cudaError_t launch_persistent_v0(int grid, int block, size_t smem, cudaStream_t st, LaunchTable* tt, int* counter, int early);
cudaError_t launch_persistent_v1(int grid, int block, size_t smem, cudaStream_t st, LaunchTable* tt, int* counter, int early);
cudaError_t launch_persistent_v2(int grid, int block, size_t smem, cudaStream_t st, LaunchTable* tt, int* counter, int early);
cudaError_t launch_packed_v0(int block, size_t smem, cudaStream_t st, const QueryParams& P, const PackedGroups& G);
cudaError_t launch_packed_v1(int block, size_t smem, cudaStream_t st, const QueryParams& P, const PackedGroups& G);
cudaError_t launch_packed_v2(int block, size_t smem, cudaStream_t st, const QueryParams& P, const PackedGroups& G);
cudaError_t launch_solve_socp(int64_t n, int nx, int n_ort, int d1, int d2, const double* c, const double* G, const double* h, double tol, int max_iters,
                              double* x, double* s, double* z, int32_t* status, int32_t* iters, double* trace, cudaStream_t st);   // dcol_socp.cu
int socp_limit(int which);
static const void* persistent_kernel_ptr(int var) {
  return var == 2 ? persistent_kernel_ptr_v2() : (var == 1 ? persistent_kernel_ptr_v1() : persistent_kernel_ptr_v0());
}
static cudaError_t launch_persistent(int var, int grid, int block, size_t smem, cudaStream_t st, LaunchTable* tt, int* counter, int early = 0) {
  return var == 2 ? launch_persistent_v2(grid, block, smem, st, tt, counter, early)
       : (var == 1 ? launch_persistent_v1(grid, block, smem, st, tt, counter, early) : launch_persistent_v0(grid, block, smem, st, tt, counter, early));
}
}
using namespace dcol;

namespace {

constexpr int NSIDE = 6;            // side streams: one per shared-memory tier
constexpr int MAX_KEYS = 4096;      // counting-sort key space
constexpr int MAX_SMEM = 227 * 1024;
thread_local std::string g_create_err;

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
  ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_); return -2; } } while (0)

struct DevBuf {
  void* p = nullptr; size_t cap = 0;
  cudaError_t need(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    size_t want = bytes + bytes / 4 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

// Host-side copy engine for calls on PAGEABLE caller memory (plain numpy / Julia arrays).  The driver moves pageable
// memory through its own bounce buffers on one thread (5-10 GB/s, and synchronously); the library instead copies
// between the caller's arrays and its own pinned staging area with a few threads, chunk by chunk, while the GPU works
// on the neighbouring chunks (run_query).  Slices of 1 MiB are handed out through one atomic counter; the calling
// thread takes part.
struct CopyJob { char* dst; const char* src; size_t bytes; };
struct HostCopier {
  // one batch of slices; late-waking workers that still hold an old batch find its counter exhausted and do nothing
  struct Batch { std::vector<CopyJob> jobs; std::atomic<size_t> next{0}, done{0}; };
  std::vector<std::thread> th;
  std::mutex m; std::condition_variable cv;
  std::shared_ptr<Batch> cur;
  std::atomic<uint64_t> gen{0};
  std::atomic<bool> stop{false};
  static void add(std::vector<CopyJob>& J, void* dst, const void* src, size_t bytes) {
    constexpr size_t SL = 512u << 10;
    for (size_t o = 0; o < bytes; o += SL) J.push_back({(char*)dst + o, (const char*)src + o, std::min(SL, bytes - o)});
  }
  static void drain(Batch& B) {
    for (;;) {
      const size_t i = B.next.fetch_add(1, std::memory_order_relaxed);
      if (i >= B.jobs.size()) break;
      memcpy(B.jobs[i].dst, B.jobs[i].src, B.jobs[i].bytes);
      B.done.fetch_add(1, std::memory_order_release);
    }
  }
  static void relax() {
#if defined(__x86_64__) || defined(__i386__)
    __builtin_ia32_pause();
#endif
  }
  // Workers spin for a short while after their last batch before they block: the copy phases of one call (and of calls
  // issued back to back) are fractions of a millisecond apart, and a futex wake-up costs as much as copying a megabyte.
  void loop() {
    uint64_t seen = 0;
    for (;;) {
      const auto t0 = std::chrono::steady_clock::now();
      while (gen.load(std::memory_order_acquire) == seen && !stop.load(std::memory_order_relaxed)) {
        if (std::chrono::steady_clock::now() - t0 > std::chrono::microseconds(1500)) {
          std::unique_lock<std::mutex> lk(m);
          cv.wait(lk, [&] { return stop.load() || gen.load() != seen; });
          break;
        }
        relax();
      }
      if (stop.load()) return;
      std::shared_ptr<Batch> B;
      { std::lock_guard<std::mutex> lk(m); B = cur; seen = gen.load(); }
      if (B) drain(*B);
    }
  }
  void start(int n) { for (int i = 0; i < n; i++) th.emplace_back([this] { loop(); }); }
  void run(std::vector<CopyJob>& J) {      // consumes J
    if (J.empty()) return;
    if (th.empty() || J.size() < 4) { for (const CopyJob& j : J) memcpy(j.dst, j.src, j.bytes); J.clear(); return; }
    std::shared_ptr<Batch> B = std::make_shared<Batch>();
    B->jobs.swap(J);
    { std::lock_guard<std::mutex> lk(m); cur = B; gen.fetch_add(1, std::memory_order_release); }
    cv.notify_all();
    drain(*B);
    while (B->done.load(std::memory_order_acquire) < B->jobs.size()) relax();
  }
  ~HostCopier() {
    { std::lock_guard<std::mutex> lk(m); stop.store(true); }
    cv.notify_all();
    for (std::thread& t : th) t.join();
  }
};

}  // namespace

struct dcol_ctx {
  int device = 0;
  cudaStream_t own = nullptr, user = nullptr;
  cudaStream_t side[NSIDE] = {};
  cudaEvent_t ev_fork = nullptr, ev_join[NSIDE] = {}, ev_k0 = nullptr, ev_k1 = nullptr;
  std::string err;
  // shape table
  int n_shapes = 0;
  std::vector<DevShape> h_shapes;
  std::vector<int> key_cp;            // dense key -> class pair (c1*NCLS+c2)
  int n_keys = 0; bool use_keymap = false;
  int max_rows_cls[NCLS] = {};
  DevBuf d_shapes, d_rows, d_cls, d_keymap;
  // static launch plan derived from the shape table (dcol_set_shapes): one entry per shared-memory tier
  struct Tier { int n_bins = 0, rows_cap = 0, block = 0, pitch = 1; size_t smem = 0; int grid[3] = {0, 0, 0};
                int cp[MAX_BINS]; int key_begin[MAX_BINS]; int key_end[MAX_BINS]; int ov_stride[MAX_BINS]; int max_stride = 0; };
  std::vector<Tier> tiers;
  // Small batches (<= SMALL_MAX pairs) run every class pair in ONE launch with the row store of the largest tier: at
  // that size the second launch, its side stream and the events between them cost more than the smaller L1 (only when
  // the table has more than one tier and four warps of the merged row store fit an SM)
  Tier small_tier; bool has_small = false;
  // per-call scratch.  Three slots, so that asynchronous submissions (dcol_submit) overlap: one staging its inputs,
  // one computing, one returning results; the synchronous entry points use slot 0.
  static constexpr int MAX_CHUNKS = 8;
  static constexpr int N_SLOTS = 3;
  struct Slot {
    DevBuf d_in, d_out, d_recs, d_inv, d_sorted, d_hist, d_tables, d_ov;
    cudaEvent_t ev_in[MAX_CHUNKS] = {}, ev_done[MAX_CHUNKS] = {}, ev_t0[MAX_CHUNKS] = {}, ev_t1[MAX_CHUNKS] = {};
    cudaEvent_t ev_out[MAX_CHUNKS] = {};  // results of a chunk have landed in the pinned staging area (pageable-memory pipeline)
    cudaEvent_t ev_end = nullptr;       // recorded behind the last operation of the submission (its own join point)
    bool pending = false; int ticket = -1, result = 0;
    int nchunks = 0, nk = 0; size_t hist_stride = 0; bool piped = false;
    // bad-index flags: mapped pinned memory the binning kernels write directly (error path only), so no copy of a
    // 4-byte flag sits behind every call
    int* h_bad = nullptr; int* d_bad_map = nullptr;
    double ms = -1.0;
  };
  Slot slots[N_SLOTS];
  int next_ticket = 0;
  // outcome of submissions whose slot has been reused since (dcol_wait on an older ticket still reports it)
  static constexpr int N_DONE = 16;
  int done_ticket[N_DONE]; int done_result[N_DONE]; std::string done_err[N_DONE]; int done_head = 0;
  dcol_ctx() { for (int i = 0; i < N_DONE; i++) { done_ticket[i] = -1; done_result[i] = 0; } }
  DevBuf d_scene, d_pairs, d_plan, d_wrench, d_brad, d_cull;
  double cull_alpha = 0.0;              // > 0: broad-phase cull of scene operators at this alpha (dcol_scene_set_cull)
  int* h_count = nullptr;               // pinned landing area of the survivor count
  // pinned staging of small host-pointer calls: one H2D and one D2H copy per call instead of six and five
  char* h_stage = nullptr; size_t h_stage_cap = 0;
  // pinned staging + copy threads of large calls on pageable caller memory (created on first use)
  char* h_pin = nullptr; size_t h_pin_cap = 0;
  HostCopier* copier = nullptr;
  int copy_threads = 0;                 // 0: min(8, cores / 2); groups divide that between their members
  int host_pipe = -1;                   // -1: not probed yet (DCOL_HOST_PIPE=0 disables, DCOL_HOST_THREADS sets the thread count)
  cudaStream_t s_copy = nullptr, s_out = nullptr, s_comp = nullptr;
  int n_sm = 148;
  int64_t launches = 0;
  double last_ms = -1.0;
  int max_iters = MAX_ITERS;            // testing knob (dcol_set_max_iters); the reference's 50 otherwise
  double* warm = nullptr; int64_t warm_cap = 0; double warm_mu = 1e-2;   // opt-in warm start (dcol_set_warm_start)
  bool copy_only = false;               // measurement knob (dcol_set_copy_only): host<->device copies without kernels
  double* mu_out = nullptr;             // optional per-pair final duality measure (dcol_set_mu_output), same memory space as the query's outputs
  cudaStream_t stream() const { return user ? user : own; }
};

static int finish_query(dcol_ctx* ctx, int slot);
static int drain_all(dcol_ctx* ctx);

// ---------------------------------------------------------------------------
// binning kernels: counting sort of pair indices by the UNORDERED class pair (lo, hi)[, then shape1, shape2]: a pair
// whose classes arrive as (hi, lo) lands in the bin of (lo, hi) and is packed with its primitives exchanged (k_pack), so
// the solver only ever sees class(prim 1) <= class(prim 2)
// ---------------------------------------------------------------------------
__device__ __forceinline__ int pair_key(const int32_t* cls, const int32_t* keymap, int n_shapes, int s1, int s2) {
  if (keymap) return keymap[s1 * n_shapes + s2];
  const int a = cls[s1], b = cls[s2];
  return min(a, b) * NCLS + max(a, b);
}

__global__ void k_hist(const int32_t* __restrict__ shape1, const int32_t* __restrict__ shape2, int64_t n,
                       const int32_t* __restrict__ cls, const int32_t* __restrict__ keymap, int n_shapes, int n_keys,
                       int* __restrict__ hist, int* __restrict__ bad) {
  extern __shared__ int sh[];
  for (int k = threadIdx.x; k < n_keys; k += blockDim.x) sh[k] = 0;
  __syncthreads();
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int s1 = shape1[i], s2 = shape2[i];
    if ((unsigned)s1 >= (unsigned)n_shapes || (unsigned)s2 >= (unsigned)n_shapes) { atomicExch(bad, 1); continue; }
    atomicAdd(&sh[pair_key(cls, keymap, n_shapes, s1, s2)], 1);
  }
  __syncthreads();
  for (int k = threadIdx.x; k < n_keys; k += blockDim.x) if (sh[k]) atomicAdd(&hist[k], sh[k]);
}

// Where the pairs of a query come from: per-pair pose rows (pose1 == NULL) or pose tables indexed per pair.
struct PairSource {
  const int32_t* shape1; const int32_t* shape2;
  const int32_t* pose1; const int32_t* pose2;      // optional: row of (r1, att1) / (r2, att2) used by the pair
  const double* r1; const double* att1; const double* r2; const double* att2;
  int na;                                          // attitude stride: 4 quaternion, 3 MRP
  const int32_t* oidx;                             // optional: where pair i's results go (culled scene lists); NULL = i
};

// The 16 doubles of pair i's 128-byte record.  Pairs whose classes arrive as (hi, lo) are stored with their primitives
// exchanged, so the solver only ever sees (lo, hi).
__device__ __forceinline__ void load_pair(const PairSource& src, int64_t i, int s1, int s2, bool swp, double (&v)[16]) {
  const int64_t i1 = src.pose1 ? (int64_t)src.pose1[i] : i, i2 = src.pose2 ? (int64_t)src.pose2[i] : i;
  const double* ra = src.r1 + 3 * i1; const double* aa = src.att1 + src.na * i1;
  const double* rb = src.r2 + 3 * i2; const double* ab = src.att2 + src.na * i2;
  v[0] = __ldg(ra); v[1] = __ldg(ra + 1); v[2] = __ldg(ra + 2);
  v[3] = __ldg(aa); v[4] = __ldg(aa + 1); v[5] = __ldg(aa + 2); v[6] = (src.na == 4) ? __ldg(aa + 3) : 0.0;
  v[7] = __ldg(rb); v[8] = __ldg(rb + 1); v[9] = __ldg(rb + 2);
  v[10] = __ldg(ab); v[11] = __ldg(ab + 1); v[12] = __ldg(ab + 2); v[13] = (src.na == 4) ? __ldg(ab + 3) : 0.0;
  if (swp) {
#pragma unroll
    for (int a = 0; a < 7; a++) { const double t = v[a]; v[a] = v[7 + a]; v[7 + a] = t; }
    const int t = s1; s1 = s2; s2 = t;
  }
  v[14] = __hiloint2double(s2, s1);
  v[15] = __hiloint2double(swp ? DCOL_REC_SWAPPED : 0, src.oidx ? src.oidx[i] : (int)i);
}

// counting-sort scatter: inv[i] = sorted position of pair i (coalesced writes; -1 for pairs with a bad shape index).
// k_pack then writes the records and k_unpermute uses the same array to bring the results back into caller order.
__global__ void k_scatter(const int32_t* __restrict__ shape1, const int32_t* __restrict__ shape2, int64_t n,
                          const int32_t* __restrict__ cls, const int32_t* __restrict__ keymap,
                          int n_shapes, int n_keys, int* __restrict__ cursor, int32_t* __restrict__ inv) {
  extern __shared__ int sh[];          // [0,n_keys) local counts, then [n_keys, 2 n_keys) global bases
  int64_t chunk = (int64_t)blockDim.x * 8;
  for (int64_t base = blockIdx.x * chunk; base < n; base += (int64_t)gridDim.x * chunk) {
    for (int k = threadIdx.x; k < n_keys; k += blockDim.x) sh[k] = 0;
    __syncthreads();
    int key[8], rank[8];
#pragma unroll
    for (int u = 0; u < 8; u++) {
      int64_t i = base + u * blockDim.x + threadIdx.x;
      key[u] = -1;
      if (i < n) {
        int s1 = shape1[i], s2 = shape2[i];
        if ((unsigned)s1 < (unsigned)n_shapes && (unsigned)s2 < (unsigned)n_shapes) {
          key[u] = pair_key(cls, keymap, n_shapes, s1, s2);
          rank[u] = atomicAdd(&sh[key[u]], 1);
        }
      }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < n_keys; k += blockDim.x) sh[n_keys + k] = sh[k] ? atomicAdd(&cursor[k], sh[k]) : 0;
    __syncthreads();
#pragma unroll
    for (int u = 0; u < 8; u++) {
      int64_t i = base + u * blockDim.x + threadIdx.x;
      if (i < n) inv[i] = (key[u] >= 0) ? sh[n_keys + key[u]] + rank[u] : -1;
    }
    __syncthreads();
  }
}

// Streaming pack: every thread reads its pair in caller order (coalesced), the block transposes through shared memory, and
// eight consecutive lanes then write one record (8 x 16 bytes = one 128-byte line) at its sorted position, so each warp
// store instruction writes four whole cache lines instead of 32 half-sectors of 32 different lines.
constexpr int PACK_THREADS = 256, PACK_PITCH = 18;      // doubles per staged record: 16 + 2 (keeps double2 alignment, spreads banks)
__global__ void __launch_bounds__(PACK_THREADS) k_pack(PairSource src, int64_t n, const int32_t* __restrict__ cls, const int32_t* __restrict__ inv,
                                                        PairRec* __restrict__ recs) {
  __shared__ double tile[PACK_THREADS * PACK_PITCH];
  __shared__ int spos[PACK_THREADS];
  const int tid = threadIdx.x;
  for (int64_t base = blockIdx.x * (int64_t)PACK_THREADS; base < n; base += (int64_t)gridDim.x * PACK_THREADS) {
    const int64_t i = base + tid;
    int pos = -1;
    if (i < n) pos = inv[i];
    if (pos >= 0) {
      const int s1 = src.shape1[i], s2 = src.shape2[i];
      double v[16];
      load_pair(src, i, s1, s2, cls[s1] > cls[s2], v);
      double2* q = reinterpret_cast<double2*>(tile + tid * PACK_PITCH);
#pragma unroll
      for (int a = 0; a < 8; a++) q[a] = make_double2(v[2 * a], v[2 * a + 1]);
    }
    spos[tid] = pos;
    __syncthreads();
    const int c = tid & 7;
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const int r = k * 32 + (tid >> 3);
      const int p = spos[r];
      if (p >= 0) reinterpret_cast<double2*>(recs + p)[c] = reinterpret_cast<const double2*>(tile + r * PACK_PITCH)[c];
    }
    __syncthreads();
  }
}

// results from sorted order (where the solver stored them with coalesced writes) back to the caller's pair order:
// thread i gathers its pair's results from position inv[i] (L2-resident for batches up to a few million pairs) and
// writes them coalesced
__global__ void k_unpermute(int64_t n, const int32_t* __restrict__ inv, const int32_t* __restrict__ oidx,
                            const double* __restrict__ alpha_s, const double* __restrict__ x_s, const int32_t* __restrict__ status_s,
                            const int32_t* __restrict__ iters_s, const double* __restrict__ mu_s,
                            double* __restrict__ alpha, double* __restrict__ x, int32_t* __restrict__ status, int32_t* __restrict__ iters,
                            double* __restrict__ mu) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int pos = inv[i];
    if (pos < 0) continue;
    const int64_t o = oidx ? (int64_t)oidx[i] : i;
    alpha[o] = alpha_s[pos];
    if (x) { x[3 * o] = x_s[3 * (int64_t)pos]; x[3 * o + 1] = x_s[3 * (int64_t)pos + 1]; x[3 * o + 2] = x_s[3 * (int64_t)pos + 2]; }
    if (status) status[o] = status_s[pos];
    if (iters) iters[o] = iters_s[pos];
    if (mu) mu[o] = mu_s[pos];
  }
}

// Instantiates the static per-tier launch plan (uploaded once by dcol_set_shapes; key ranges parked in begin[]/end[])
// for this call — query parameters arrive as a kernel argument, bin offsets from the counting sort — so that neither a
// host read of the histogram nor a host-to-device copy sits between the sort and the solver.  Also carves the
// straggler overflow area: bin b gets tiles_b * spill_t records (a tile hands over at most spill_t pairs).
// Called by every thread of a block (block-wide barriers inside); one thread per bin fetches the bin's range, only
// the two short prefix sums are serial.
constexpr int TABLE_THREADS = 64;
static_assert(MAX_BINS <= TABLE_THREADS, "one thread per bin");
__device__ void build_tables_dev(const LaunchTable* __restrict__ plan, LaunchTable* __restrict__ tables, int n_tiers,
                                 const int* __restrict__ off, double* ov_base, const QueryParams& P, int spill_div,
                                 int* s_nt, long long* s_ovsz, long long* s_ov) {
  const int tid = threadIdx.x;
  if (tid == 0) *s_ov = 0;
  for (int t = 0; t < n_tiers; t++) {
    {
      const int* src = reinterpret_cast<const int*>(plan + t);
      int* dst = reinterpret_cast<int*>(tables + t);
      for (int i = tid; i < (int)(sizeof(LaunchTable) / sizeof(int)); i += blockDim.x) dst[i] = src[i];
    }
    __syncthreads();
    LaunchTable& T = tables[t];
    const int tile = T.n_tiles;            // the plan parks the tier's CTA size here
    const int spill_t = spill_div > 0 ? tile / spill_div : 0;
    const int n_bins = T.n_bins;
    if (tid < n_bins) {
      const int b = off[T.begin[tid]], e = off[T.end[tid]];
      const int nt = (e - b + tile - 1) / tile;
      T.begin[tid] = b; T.end[tid] = e;
      T.ov_cap[tid] = nt * spill_t; T.ov_count[tid] = 0;
      s_nt[tid] = nt; s_ovsz[tid] = (long long)nt * spill_t * T.ov_stride[tid];
    }
    __syncthreads();
    if (tid == 0) {
      int tiles = 0;
      long long ov = *s_ov;
      for (int k = 0; k < n_bins; k++) {
        T.tile_prefix[k] = tiles; T.ov_off[k] = ov;
        tiles += s_nt[k]; ov += s_ovsz[k];
      }
      for (int k = n_bins; k <= MAX_BINS; k++) T.tile_prefix[k] = tiles;
      *s_ov = ov;
      const int pitch = T.P.row_pitch;
      T.P = P; T.P.row_pitch = pitch;
      T.spill_t = spill_t;
      T.n_tiles = tiles; T.ov_base = ov_base; T.mode = 0; T.in_base = ov_base;
      T.tile = tile;
    }
    __syncthreads();
  }
}

// single block: exclusive scan of hist[0..n_keys) -> off[0..n_keys], cursor = off; then the launch tables
__global__ void k_scan_tables(const int* __restrict__ hist, int n_keys, int* __restrict__ off, int* __restrict__ cursor,
                              const LaunchTable* __restrict__ plan, LaunchTable* __restrict__ tables, int n_tiers,
                              double* ov_base, QueryParams P, int spill_div) {
  __shared__ int part[1024];
  __shared__ int s_nt[MAX_BINS];
  __shared__ long long s_ovsz[MAX_BINS];
  __shared__ long long s_ov;
  int per = (n_keys + blockDim.x - 1) / blockDim.x;
  int b = threadIdx.x * per, e = min(b + per, n_keys);
  int s = 0;
  for (int k = b; k < e; k++) s += hist[k];
  part[threadIdx.x] = s;
  __syncthreads();
  for (int d = 1; d < blockDim.x; d <<= 1) {
    int v = (threadIdx.x >= d) ? part[threadIdx.x - d] : 0;
    __syncthreads();
    part[threadIdx.x] += v;
    __syncthreads();
  }
  int run = part[threadIdx.x] - s;
  for (int k = b; k < e; k++) { off[k] = run; cursor[k] = run; run += hist[k]; }
  if (threadIdx.x == blockDim.x - 1) off[n_keys] = part[threadIdx.x];
  __syncthreads();                       // off[] is read back by other threads of this block below
  build_tables_dev(plan, tables, n_tiers, off, ov_base, P, spill_div, s_nt, s_ovsz, &s_ov);
}

// Small batches (n <= SMALL_MAX): histogram, scan, sorted positions and the launch tables in ONE single-block kernel, so a
// query is this launch, the packing launch and one solver launch per shared-memory tier (no hand-over, hence no resume
// launch, and results are stored at the caller's pair index directly: no un-permute pass).  Latency path of the single-pair drop-in calls
// (src/proximity.jl:3-14 is what every reference example calls in its inner loop).
constexpr int SMALL_MAX = 16384;
constexpr int SMALL_THREADS = 1024;
constexpr int SMALL_ITEMS = SMALL_MAX / SMALL_THREADS;
constexpr int SMALL_FUSED = 4096;       // up to this many pairs the preparation kernel writes the packed records itself
__global__ void __launch_bounds__(SMALL_THREADS) k_small_prep(PairSource src, int n, const int32_t* __restrict__ cls, const int32_t* __restrict__ keymap,
                             int n_shapes, int n_keys, int* __restrict__ off, int* __restrict__ counters, int n_counters, int* __restrict__ bad,
                             int32_t* __restrict__ inv, const LaunchTable* __restrict__ plan, LaunchTable* __restrict__ tables, int n_tiers,
                             double* ov_base, QueryParams P, PairRec* __restrict__ recs) {
  extern __shared__ int sh[];            // [0, n_keys]: counts, then exclusive offsets
  __shared__ int part[SMALL_THREADS];
  __shared__ int s_nt[MAX_BINS];
  __shared__ long long s_ovsz[MAX_BINS];
  __shared__ long long s_ov;
  const int tid = threadIdx.x;
  asm volatile("griddepcontrol.launch_dependents;");      // the solver launch behind this kernel may become resident now
  for (int k = tid; k <= n_keys; k += blockDim.x) sh[k] = 0;
  for (int k = tid; k < n_counters; k += blockDim.x) counters[k] = 0;
  __syncthreads();
  int key[SMALL_ITEMS], rank[SMALL_ITEMS];
#pragma unroll
  for (int u = 0; u < SMALL_ITEMS; u++) {
    const int i = u * blockDim.x + tid;
    key[u] = -1;
    if (i < n) {
      const int s1 = src.shape1[i], s2 = src.shape2[i];
      if ((unsigned)s1 < (unsigned)n_shapes && (unsigned)s2 < (unsigned)n_shapes) {
        key[u] = pair_key(cls, keymap, n_shapes, s1, s2);
        rank[u] = atomicAdd(&sh[key[u]], 1);
      } else *bad = 1;
    }
  }
  __syncthreads();
  {  // exclusive scan of the counts, in place
    const int per = (n_keys + blockDim.x - 1) / blockDim.x;
    const int b = tid * per, e = min(b + per, n_keys);
    int s = 0;
    for (int k = b; k < e; k++) s += sh[k];
    part[tid] = s;
    __syncthreads();
    for (int d = 1; d < blockDim.x; d <<= 1) {
      int v = (tid >= d) ? part[tid - d] : 0;
      __syncthreads();
      part[tid] += v;
      __syncthreads();
    }
    int run = part[tid] - s;
    for (int k = b; k < e; k++) { const int c = sh[k]; sh[k] = run; off[k] = run; run += c; }
    if (tid == blockDim.x - 1) { sh[n_keys] = part[tid]; off[n_keys] = part[tid]; }
  }
  __syncthreads();
#pragma unroll
  for (int u = 0; u < SMALL_ITEMS; u++) {
    const int i = u * blockDim.x + tid;
    if (i < n) {
      const int pos = (key[u] >= 0) ? sh[key[u]] + rank[u] : -1;
      inv[i] = pos;
      if (recs && pos >= 0) {            // a few thousand pairs at most: pack them here and save the k_pack launch
        const int s1 = src.shape1[i], s2 = src.shape2[i];
        double v[16];
        load_pair(src, i, s1, s2, cls[s1] > cls[s2], v);
        double2* q = reinterpret_cast<double2*>(recs + pos);
#pragma unroll
        for (int a = 0; a < 8; a++) q[a] = make_double2(v[2 * a], v[2 * a + 1]);
      }
    }
  }
  build_tables_dev(plan, tables, n_tiers, off, ov_base, P, 0, s_nt, s_ovsz, &s_ov);
}

// resume-pass table of ONE tier from the overflow counts its previous pass left behind (launched right behind that
// pass on the tier's stream, so the resume pass of one tier overlaps the first pass of the next); the resume pass may
// itself hand its stragglers over (spill > 0) into the other overflow buffer, which reuses the previous layout
__global__ void k_build_tables2(const LaunchTable* __restrict__ Tp, LaunchTable* __restrict__ Up, double* ov_next, int spill) {
  __shared__ int s_nt[MAX_BINS];
  const LaunchTable& T = *Tp;
  LaunchTable& U = *Up;
  const int tid = threadIdx.x;
  const int tile = T.tile, n_bins = T.n_bins;
  if (tid < n_bins) {
    const int k = tid;
    const int cnt = min(T.ov_count[k], T.ov_cap[k]);
    U.begin[k] = 0; U.end[k] = cnt; U.cp[k] = T.cp[k];
    s_nt[k] = (cnt + tile - 1) / tile;
    U.in_cap[k] = T.ov_cap[k]; U.in_off[k] = T.ov_off[k]; U.ov_stride[k] = T.ov_stride[k];
    U.ov_cap[k] = spill ? T.ov_cap[k] : 0; U.ov_off[k] = T.ov_off[k]; U.ov_count[k] = 0;
  }
  __syncthreads();
  if (tid == 0) {
    int tiles = 0;
    for (int k = 0; k < n_bins; k++) { U.tile_prefix[k] = tiles; tiles += s_nt[k]; }
    for (int k = n_bins; k <= MAX_BINS; k++) U.tile_prefix[k] = tiles;
    U.P = T.P; U.tile = tile;
    U.n_bins = n_bins; U.n_tiles = tiles; U.in_base = T.ov_base; U.ov_base = ov_next; U.mode = 1;
    U.spill_t = spill ? T.spill_t : 0; U.rows_cap = T.rows_cap;
  }
}

// (i, j), i < j in combinations(1:N, 2) order (the pair list of examples/contact_physics/ncp_contact_simulation.jl:48),
// generated on the device: pair k of row i starts at k_i = i(2N - i - 1)/2.
__device__ __forceinline__ void pair_of_index(int n_bodies, int64_t k, int64_t& i, int64_t& j) {
  const double N = (double)n_bodies;
  i = (int64_t)floor(((2.0 * N - 1.0) - sqrt((2.0 * N - 1.0) * (2.0 * N - 1.0) - 8.0 * (double)k)) * 0.5);
  i = max((int64_t)0, min(i, (int64_t)n_bodies - 2));
  while (i > 0 && i * (2 * (int64_t)n_bodies - i - 1) / 2 > k) i--;
  while ((i + 1) * (2 * (int64_t)n_bodies - i - 2) / 2 <= k) i++;
  j = k - i * (2 * (int64_t)n_bodies - i - 1) / 2 + i + 1;
}
// The n_pairs pairs from index k0 on (a slice of the enumeration: one member of a device group, one rank); row k of the
// outputs belongs to pair k0 + k.
__global__ void k_allpairs(int n_bodies, int64_t k0, int64_t n_pairs, const int32_t* __restrict__ shape,
                           int32_t* __restrict__ s1, int32_t* __restrict__ s2, int32_t* __restrict__ p1, int32_t* __restrict__ p2) {
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n_pairs; k += (int64_t)gridDim.x * blockDim.x) {
    int64_t i, j;
    pair_of_index(n_bodies, k0 + k, i, j);
    s1[k] = shape[i]; s2[k] = shape[j]; p1[k] = (int32_t)i; p2[k] = (int32_t)j;
  }
}

// Broad phase of the all-pairs scene operators (opt-in, dcol_scene_set_cull): every primitive scales about its body
// origin r~ = r + Qn r_offset and lies inside the ball of radius rho around it, so alpha* >= |r~_i - r~_j| / (rho_i + rho_j).
// Pairs whose bound exceeds alpha_cut are answered here (alpha = that lower bound, status DCOL_STATUS_CULLED, zero
// derivative rows, NaN witness point); the others are appended to a compact pair list for the conic solver, together with
// the index their results go to.  centre[] holds r~ per body (k_body_centres).
__global__ void k_body_centres(int n_bodies, const int32_t* __restrict__ shape, const DevShape* __restrict__ shapes, int n_shapes,
                               const double* __restrict__ r, const double* __restrict__ att, int att_kind, double* __restrict__ centre, int* __restrict__ bad) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n_bodies) return;
  const int s = shape[b];
  double c[3] = {r[3 * b], r[3 * b + 1], r[3 * b + 2]};
  if ((unsigned)s >= (unsigned)n_shapes) { *bad = 1; }
  else if (shapes[s].has_offset) {
    double Q[3][3];
    if (att_kind == 0) { double q[4] = {att[4 * b], att[4 * b + 1], att[4 * b + 2], att[4 * b + 3]}; dcm_from_q<double>(q, Q); }
    else { double q[3] = {att[3 * b], att[3 * b + 1], att[3 * b + 2]}; dcm_from_mrp<double>(q, Q); }
    for (int a = 0; a < 3; a++) c[a] += Q[a][0] * shapes[s].r_off[0] + Q[a][1] * shapes[s].r_off[1] + Q[a][2] * shapes[s].r_off[2];
  }
  centre[3 * b] = c[0]; centre[3 * b + 1] = c[1]; centre[3 * b + 2] = c[2];
}
__global__ void k_allpairs_cull(int n_bodies, int64_t k0, int64_t n_pairs, const int32_t* __restrict__ shape, int n_shapes, const double* __restrict__ brad,
                                const double* __restrict__ centre, double alpha_cut, int nd,
                                int32_t* __restrict__ s1, int32_t* __restrict__ s2, int32_t* __restrict__ p1, int32_t* __restrict__ p2,
                                int32_t* __restrict__ oidx, int* __restrict__ count,
                                double* __restrict__ alpha, double* __restrict__ x, double* __restrict__ deriv, int32_t* __restrict__ status,
                                int32_t* __restrict__ iters) {
  const int64_t span = (int64_t)gridDim.x * blockDim.x;
  const int64_t rounds = (n_pairs + span - 1) / span;
  for (int64_t rd = 0; rd < rounds; rd++) {
    const int64_t k = rd * span + blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    bool keep = false;
    int64_t i = 0, j = 0; int si = 0, sj = 0;
    if (k < n_pairs) {
      pair_of_index(n_bodies, k0 + k, i, j);
      si = shape[i]; sj = shape[j];
      keep = true;
      if ((unsigned)si < (unsigned)n_shapes && (unsigned)sj < (unsigned)n_shapes) {
        const double dx = centre[3 * i] - centre[3 * j], dy = centre[3 * i + 1] - centre[3 * j + 1], dz = centre[3 * i + 2] - centre[3 * j + 2];
        const double lb = sqrt(dx * dx + dy * dy + dz * dz) / (brad[si] + brad[sj]);
        if (lb > alpha_cut) {
          keep = false;
          alpha[k] = lb;
          if (status) status[k] = DCOL_STATUS_CULLED;
          if (iters) iters[k] = 0;
          if (x) { const double qn = __longlong_as_double(0x7ff8000000000000LL); x[3 * k] = qn; x[3 * k + 1] = qn; x[3 * k + 2] = qn; }
          if (deriv) for (int a = 0; a < nd; a++) deriv[(int64_t)nd * k + a] = 0.0;
        }
      }
    }
    // warp-aggregated append of the survivors
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    if (m) {
      const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
      int base = 0;
      if (lane == leader) base = atomicAdd(count, __popc(m));
      base = __shfl_sync(0xffffffffu, base, leader);
      if (keep) {
        const int slot = base + __popc(m & ((1u << lane) - 1));
        s1[slot] = si; s2[slot] = sj; p1[slot] = (int32_t)i; p2[slot] = (int32_t)j; oidx[slot] = (int32_t)k;
      }
    }
  }
}

// Per-body accumulation of the weighted pair gradients of an all-pairs scene (the loop of
// examples/contact_physics/ncp_contact_simulation.jl:133-139):  out[b] = sum over pairs k containing body b of
// lambda[k] * [ d alpha_k / d r_b ; P(att_b)' d alpha_k / d att_b ],  P = G(q) = L(q) H for quaternions
// (variational_utils.jl:24-26), identity for MRPs.  One block per body, no atomics: the pairs (b, j), j > b are the
// contiguous run starting at k_b = b(2N-b-1)/2 (body b is side 1), the pairs (i, b), i < b sit at k_i + (b-i-1)
// (body b is side 2).  Fixed summation order, so the result is reproducible.
// With a slice [k0, k1) of the pair enumeration (grad / lambda / culled hold rows k0 .. k1-1), out[b] is the slice's part of
// the sum; the parts of all slices add up to the whole (dcol_group_scene_wrench, sharded ranks).
__global__ void k_scene_wrench(int n_bodies, int64_t k0, int64_t k1, const double* __restrict__ grad, const double* __restrict__ lambda,
                               const int32_t* __restrict__ culled, const double* __restrict__ att, int att_kind, double* __restrict__ out) {
  const int b = blockIdx.x;
  const int na = (att_kind == 0) ? 4 : 3, nt = 6 + 2 * na, ns = 3 + na;
  const int64_t N = n_bodies;
  double acc[7] = {0, 0, 0, 0, 0, 0, 0};
  for (int j = threadIdx.x; j < n_bodies; j += blockDim.x) {
    if (j == b) continue;
    int64_t k; int side;
    if (j > b) { k = (int64_t)b * (2 * N - b - 1) / 2 + (j - b - 1); side = 0; }
    else       { k = (int64_t)j * (2 * N - j - 1) / 2 + (b - j - 1); side = 1; }
    if (k < k0 || k >= k1) continue;                             // another slice's pair
    k -= k0;
    if (culled && culled[k] == DCOL_STATUS_CULLED) continue;     // far pair: its gradient row was never computed
    const double w = lambda[k];
    const double* g = grad + k * nt + side * ns;
    for (int a = 0; a < ns; a++) acc[a] += w * g[a];
  }
  __shared__ double red[7][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  for (int a = 0; a < 7; a++) {
    double v = acc[a];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (lane == 0) red[a][warp] = v;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t[7];
    for (int a = 0; a < 7; a++) { double v = 0.0; for (int w = 0; w < nw; w++) v += red[a][w]; t[a] = v; }
    double* o = out + 6 * (int64_t)b;
    o[0] = t[0]; o[1] = t[1]; o[2] = t[2];
    if (att_kind == 0) {
      // G(q)' g = -v g0 + s gv - v x gv   (q = [s; v], g = [g0; gv])
      const double* q = att + 4 * (int64_t)b;
      const double s0 = q[0], v1 = q[1], v2 = q[2], v3 = q[3];
      const double g0 = t[3], g1 = t[4], g2 = t[5], g3 = t[6];
      o[3] = -v1 * g0 + s0 * g1 - (v2 * g3 - v3 * g2);
      o[4] = -v2 * g0 + s0 * g2 - (v3 * g1 - v1 * g3);
      o[5] = -v3 * g0 + s0 * g3 - (v1 * g2 - v2 * g1);
    } else {
      o[3] = t[3]; o[4] = t[4]; o[5] = t[5];
    }
  }
}

// Planar vehicle against static obstacles (examples/trajectory_optimization/piano_mover.jl:39-53): state (x, y, theta)
// -> pose r = (x, y, 0), MRP p = (0, 0, tan(theta/4)); pair k = state * n_obs + obstacle.
__global__ void k_se2_pairs(int64_t n_states, int n_obs, const double* __restrict__ xyt, int32_t vehicle, const int32_t* __restrict__ obs_shape,
                            double* __restrict__ r1, double* __restrict__ p1,
                            int32_t* __restrict__ s1, int32_t* __restrict__ s2, int32_t* __restrict__ i1, int32_t* __restrict__ i2) {
  const int64_t np = n_states * n_obs;
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < np; k += (int64_t)gridDim.x * blockDim.x) {
    const int64_t s = k / n_obs; const int o = (int)(k - s * n_obs);
    s1[k] = vehicle; s2[k] = obs_shape[o]; i1[k] = (int32_t)s; i2[k] = o;
    if (o == 0) {
      r1[3 * s] = xyt[3 * s]; r1[3 * s + 1] = xyt[3 * s + 1]; r1[3 * s + 2] = 0.0;
      p1[3 * s] = 0.0; p1[3 * s + 1] = 0.0; p1[3 * s + 2] = tan(0.25 * xyt[3 * s + 2]);
    }
  }
}
// c = 1 - alpha ; dc/d(x, y, theta) = (-da/dr1_x, -da/dr1_y, -da/dp1_z / (4 cos^2(theta/4)))   (piano_mover.jl:45-52)
__global__ void k_se2_rows(int64_t np, int n_obs, const double* __restrict__ xyt, const double* __restrict__ alpha,
                           const double* __restrict__ grad, double* __restrict__ c, double* __restrict__ Jc) {
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < np; k += (int64_t)gridDim.x * blockDim.x) {
    const int64_t s = k / n_obs;
    const double ct = cos(0.25 * xyt[3 * s + 2]);
    const double* g = grad + 12 * k;
    c[k] = 1.0 - alpha[k];
    Jc[3 * k] = -g[0]; Jc[3 * k + 1] = -g[1]; Jc[3 * k + 2] = -g[5] / (4.0 * ct * ct);
  }
}

// Rigid vehicle against static obstacles (examples/trajectory_optimization/cone_through_wall.jl:56-68,
// cluttered_hallway_quadrotor.jl:83-104): every state row holds the vehicle's position at columns r_col.. and its MRP at
// p_col.. (the examples' 12-state [r; v; p; omega]: r_col 0, p_col 6); pair k = state * n_obs + obstacle.
__global__ void k_se3_pairs(int64_t n_states, int n_obs, const double* __restrict__ X, int stride, int r_col, int p_col, int32_t vehicle,
                            const int32_t* __restrict__ obs_shape, double* __restrict__ r1, double* __restrict__ p1,
                            int32_t* __restrict__ s1, int32_t* __restrict__ s2, int32_t* __restrict__ i1, int32_t* __restrict__ i2) {
  const int64_t np = n_states * n_obs;
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < np; k += (int64_t)gridDim.x * blockDim.x) {
    const int64_t s = k / n_obs; const int o = (int)(k - s * n_obs);
    s1[k] = vehicle; s2[k] = obs_shape[o]; i1[k] = (int32_t)s; i2[k] = o;
    if (o == 0) {
      const double* x = X + (int64_t)stride * s;
#pragma unroll
      for (int a = 0; a < 3; a++) { r1[3 * s + a] = x[r_col + a]; p1[3 * s + a] = x[p_col + a]; }
    }
  }
}
// c = 1 - alpha ; the non-zero columns of the examples' constraint Jacobian rows: -d alpha/d r1, -d alpha/d p1
__global__ void k_se3_rows(int64_t np, const double* __restrict__ alpha, const double* __restrict__ grad, double* __restrict__ c, double* __restrict__ Jc) {
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < np; k += (int64_t)gridDim.x * blockDim.x) {
    const double* g = grad + 12 * k;
    c[k] = 1.0 - alpha[k];
    if (Jc) {
#pragma unroll
      for (int a = 0; a < 6; a++) Jc[6 * k + a] = -g[a]; }
  }
}

// FP64 FMA throughput micro-benchmark (roofline denominator; MEASURED_PEAKS.json has no FP64 figure): 8 independent
// chains per thread, 16 rounds per loop trip (128 DFMA against 3 loop instructions).  Measured on B200 with ncu
// (profiles/r02z_k_dfma_ncu.txt, scripts/dfma_probe.cu):
//   TWO_SRC  x = fma(x, a, x)   two distinct 64-bit register sources:   37.0 TFLOP/s, FP64 pipe ~100 % busy -- the pipe's peak
//                                                                       (nominal 148 SM x 64 DFMA/clk x 2 x 1.965 GHz = 37.2)
//   !TWO_SRC x = fma(x, a, b)   three distinct sources:                 34.1 TFLOP/s, pipe 92 % busy -- whatever the chain count,
//                                                                       unrolling or CTA shape: the third 64-bit operand read
// The roofline uses the first (higher) figure; the second is what a stream of general DFMAs can reach.
template <bool TWO_SRC>
__global__ void k_dfma(double* out, int iters, double a, double b) {
  double x0 = threadIdx.x * 1e-3, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
#pragma unroll 1
  for (int i = 0; i < iters; i += 16) {
#pragma unroll
    for (int u = 0; u < 16; u++) {
      if (TWO_SRC) {
        x0 = fma(x0, a, x0); x1 = fma(x1, a, x1); x2 = fma(x2, a, x2); x3 = fma(x3, a, x3);
        x4 = fma(x4, a, x4); x5 = fma(x5, a, x5); x6 = fma(x6, a, x6); x7 = fma(x7, a, x7);
      } else {
        x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
        x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
      }
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

// ---------------------------------------------------------------------------
// shape table
// ---------------------------------------------------------------------------
static int type_to_cls(int type) {
  switch (type) {
    case T_POLYTOPE: return CLS_POLYTOPE;
    case T_CAPSULE: case T_CYLINDER: return CLS_CAPCYL;
    case T_CONE: return CLS_CONE;
    case T_SPHERE: return CLS_SPHERE;
    case T_ELLIPSOID: return CLS_ELL;
    case T_POLYGON: return CLS_POLYGON;
  }
  return -1;
}

// Radius of the smallest origin-centred ball that contains a primitive in its body frame (the primitive scales about
// that origin, so alpha times this ball contains the scaled primitive): the broad-phase bound of the scene operators.
// Polytopes / polygons: largest vertex norm by enumerating the intersections of constraint triples / pairs (done once
// per shape table; a few thousand candidates for the face counts the reference uses).
static double bounding_radius(int type, const double* p) {
  switch (type) {
    case T_SPHERE: return p[0];
    case T_CAPSULE: return 0.5 * p[1] + p[0];
    case T_CYLINDER: return std::sqrt(0.25 * p[1] * p[1] + p[0] * p[0]);
    case T_CONE: { const double H = p[0], tb = std::tan(p[1]); return std::max(0.75 * H, std::sqrt(H * H / 16 + H * H * tb * tb)); }
    case T_ELLIPSOID: {          // ||U x|| <= 1  =>  |x| <= 1/sqrt(lambda_min(U'U)); eigenvalues of the symmetric 3x3 in closed form
      double M[3][3];
      for (int a = 0; a < 3; a++) for (int c = 0; c < 3; c++) M[a][c] = p[a] * p[c] + p[3 + a] * p[3 + c] + p[6 + a] * p[6 + c];
      const double q = (M[0][0] + M[1][1] + M[2][2]) / 3.0;
      const double p1 = M[0][1] * M[0][1] + M[0][2] * M[0][2] + M[1][2] * M[1][2];
      const double p2 = (M[0][0] - q) * (M[0][0] - q) + (M[1][1] - q) * (M[1][1] - q) + (M[2][2] - q) * (M[2][2] - q) + 2.0 * p1;
      double lmin = std::min(M[0][0], std::min(M[1][1], M[2][2]));
      if (p2 > 0.0) {
        const double pp = std::sqrt(p2 / 6.0);
        double B[3][3];
        for (int a = 0; a < 3; a++) for (int c = 0; c < 3; c++) B[a][c] = (M[a][c] - (a == c ? q : 0.0)) / pp;
        const double detB = B[0][0] * (B[1][1] * B[2][2] - B[1][2] * B[2][1]) - B[0][1] * (B[1][0] * B[2][2] - B[1][2] * B[2][0])
                          + B[0][2] * (B[1][0] * B[2][1] - B[1][1] * B[2][0]);
        const double rr = std::max(-1.0, std::min(1.0, 0.5 * detB));
        const double phi = std::acos(rr) / 3.0;
        lmin = q + 2.0 * pp * std::cos(phi + 2.0 * M_PI / 3.0);          // smallest of the three roots
      }
      return lmin > 0.0 ? (1.0 + 1e-12) / std::sqrt(lmin) : INFINITY;
    }
    case T_POLYTOPE: {
      const int n = (int)p[0]; const double* A = p + 1; const double* b = p + 1 + 3 * n;
      double best = 0.0; bool any = false;
      for (int i = 0; i < n; i++) for (int j = i + 1; j < n; j++) for (int k = j + 1; k < n; k++) {
        const double* a1 = A + 3 * i; const double* a2 = A + 3 * j; const double* a3 = A + 3 * k;
        const double c1[3] = {a2[1] * a3[2] - a2[2] * a3[1], a2[2] * a3[0] - a2[0] * a3[2], a2[0] * a3[1] - a2[1] * a3[0]};
        const double det = a1[0] * c1[0] + a1[1] * c1[1] + a1[2] * c1[2];
        if (std::fabs(det) < 1e-12) continue;
        const double c2[3] = {a3[1] * a1[2] - a3[2] * a1[1], a3[2] * a1[0] - a3[0] * a1[2], a3[0] * a1[1] - a3[1] * a1[0]};
        const double c3[3] = {a1[1] * a2[2] - a1[2] * a2[1], a1[2] * a2[0] - a1[0] * a2[2], a1[0] * a2[1] - a1[1] * a2[0]};
        double v[3];
        for (int q = 0; q < 3; q++) v[q] = (b[i] * c1[q] + b[j] * c2[q] + b[k] * c3[q]) / det;
        bool inside = true;
        for (int m = 0; m < n && inside; m++) inside = A[3 * m] * v[0] + A[3 * m + 1] * v[1] + A[3 * m + 2] * v[2] <= b[m] + 1e-9 * (1.0 + std::fabs(b[m]));
        if (inside) { best = std::max(best, std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2])); any = true; }
      }
      return any ? best : INFINITY;          // unbounded or degenerate: never culled
    }
    case T_POLYGON: {
      const int n = (int)p[0]; const double* A = p + 1; const double* b = p + 1 + 2 * n; const double R = p[1 + 3 * n];
      double best = 0.0; bool any = false;
      for (int i = 0; i < n; i++) for (int j = i + 1; j < n; j++) {
        const double det = A[2 * i] * A[2 * j + 1] - A[2 * i + 1] * A[2 * j];
        if (std::fabs(det) < 1e-12) continue;
        const double v[2] = {(b[i] * A[2 * j + 1] - A[2 * i + 1] * b[j]) / det, (A[2 * i] * b[j] - b[i] * A[2 * j]) / det};
        bool inside = true;
        for (int m = 0; m < n && inside; m++) inside = A[2 * m] * v[0] + A[2 * m + 1] * v[1] <= b[m] + 1e-9 * (1.0 + std::fabs(b[m]));
        if (inside) { best = std::max(best, std::sqrt(v[0] * v[0] + v[1] * v[1])); any = true; }
      }
      return any ? best + R : INFINITY;
    }
  }
  return INFINITY;
}

extern "C" {

const char* dcol_version(void) { return "dcol_b200 0.1.0 (sm_100a, fp64)"; }

int dcol_create(dcol_ctx** out, int device_id) {
  if (!out) { g_create_err = "dcol_create: out is NULL"; return -1; }
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    g_create_err = std::string("dcol_create: no CUDA device available (") + cudaGetErrorString(e) +
                   "); this library has no CPU fallback";
    return -2;
  }
  if (device_id < 0 || device_id >= ndev) { g_create_err = "dcol_create: bad device id"; return -1; }
  dcol_ctx* ctx = new dcol_ctx();
  ctx->device = device_id;
  auto fail = [&](const char* what, cudaError_t err) { g_create_err = std::string(what) + ": " + cudaGetErrorString(err); delete ctx; return -2; };
  if ((e = cudaSetDevice(device_id)) != cudaSuccess) return fail("cudaSetDevice", e);
  if ((e = cudaStreamCreateWithFlags(&ctx->own, cudaStreamNonBlocking)) != cudaSuccess) return fail("cudaStreamCreate", e);
  for (int i = 0; i < NSIDE; i++) {
    if ((e = cudaStreamCreateWithFlags(&ctx->side[i], cudaStreamNonBlocking)) != cudaSuccess) return fail("cudaStreamCreate", e);
    if ((e = cudaEventCreateWithFlags(&ctx->ev_join[i], cudaEventDisableTiming)) != cudaSuccess) return fail("cudaEventCreate", e);
  }
  if ((e = cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming)) != cudaSuccess) return fail("cudaEventCreate", e);
  if ((e = cudaStreamCreateWithFlags(&ctx->s_copy, cudaStreamNonBlocking)) != cudaSuccess) return fail("cudaStreamCreate", e);
  if ((e = cudaStreamCreateWithFlags(&ctx->s_out, cudaStreamNonBlocking)) != cudaSuccess) return fail("cudaStreamCreate", e);
  if ((e = cudaStreamCreateWithFlags(&ctx->s_comp, cudaStreamNonBlocking)) != cudaSuccess) return fail("cudaStreamCreate", e);
  for (int k = 0; k < dcol_ctx::N_SLOTS; k++)
    for (int i = 0; i < dcol_ctx::MAX_CHUNKS; i++) {
      dcol_ctx::Slot& L = ctx->slots[k];
      if ((e = cudaEventCreateWithFlags(&L.ev_in[i], cudaEventDisableTiming)) != cudaSuccess) return fail("cudaEventCreate", e);
      if ((e = cudaEventCreateWithFlags(&L.ev_done[i], cudaEventDisableTiming)) != cudaSuccess) return fail("cudaEventCreate", e);
      if ((e = cudaEventCreateWithFlags(&L.ev_out[i], cudaEventDisableTiming)) != cudaSuccess) return fail("cudaEventCreate", e);
      if ((e = cudaEventCreate(&L.ev_t0[i])) != cudaSuccess) return fail("cudaEventCreate", e);
      if ((e = cudaEventCreate(&L.ev_t1[i])) != cudaSuccess) return fail("cudaEventCreate", e);
    }
  for (int k = 0; k < dcol_ctx::N_SLOTS; k++) {
    dcol_ctx::Slot& L = ctx->slots[k];
    if ((e = cudaHostAlloc((void**)&L.h_bad, sizeof(int) * dcol_ctx::MAX_CHUNKS, cudaHostAllocMapped)) != cudaSuccess) return fail("cudaHostAlloc", e);
    if ((e = cudaHostGetDevicePointer((void**)&L.d_bad_map, L.h_bad, 0)) != cudaSuccess) return fail("cudaHostGetDevicePointer", e);
    if ((e = cudaEventCreateWithFlags(&L.ev_end, cudaEventDisableTiming)) != cudaSuccess) return fail("cudaEventCreate", e);
  }
  if ((e = cudaEventCreate(&ctx->ev_k0)) != cudaSuccess) return fail("cudaEventCreate", e);
  if ((e = cudaEventCreate(&ctx->ev_k1)) != cudaSuccess) return fail("cudaEventCreate", e);
  { cudaDeviceProp prop; if (cudaGetDeviceProperties(&prop, device_id) == cudaSuccess) ctx->n_sm = prop.multiProcessorCount; }
  *out = ctx;
  return 0;
}

void dcol_destroy(dcol_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaDeviceSynchronize();
  DevBuf* bufs[] = {&ctx->d_shapes, &ctx->d_rows, &ctx->d_cls, &ctx->d_keymap, &ctx->d_scene, &ctx->d_pairs, &ctx->d_plan, &ctx->d_wrench, &ctx->d_brad, &ctx->d_cull};
  if (ctx->h_count) cudaFreeHost(ctx->h_count);
  for (DevBuf* b : bufs) b->release();
  for (int i = 0; i < NSIDE; i++) { if (ctx->side[i]) cudaStreamDestroy(ctx->side[i]); if (ctx->ev_join[i]) cudaEventDestroy(ctx->ev_join[i]); }
  for (int k = 0; k < dcol_ctx::N_SLOTS; k++) {
    dcol_ctx::Slot& L = ctx->slots[k];
    for (int i = 0; i < dcol_ctx::MAX_CHUNKS; i++) {
      if (L.ev_in[i]) cudaEventDestroy(L.ev_in[i]);
      if (L.ev_done[i]) cudaEventDestroy(L.ev_done[i]);
      if (L.ev_out[i]) cudaEventDestroy(L.ev_out[i]);
      if (L.ev_t0[i]) cudaEventDestroy(L.ev_t0[i]);
      if (L.ev_t1[i]) cudaEventDestroy(L.ev_t1[i]);
    }
    if (L.h_bad) cudaFreeHost(L.h_bad);
    if (L.ev_end) cudaEventDestroy(L.ev_end);
    DevBuf* lb[] = {&L.d_in, &L.d_out, &L.d_recs, &L.d_inv, &L.d_sorted, &L.d_hist, &L.d_tables, &L.d_ov};
    for (DevBuf* b : lb) b->release();
  }
  if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
  if (ctx->h_pin) cudaFreeHost(ctx->h_pin);
  delete ctx->copier;
  if (ctx->s_copy) cudaStreamDestroy(ctx->s_copy);
  if (ctx->s_out) cudaStreamDestroy(ctx->s_out);
  if (ctx->s_comp) cudaStreamDestroy(ctx->s_comp);
  if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
  if (ctx->ev_k0) cudaEventDestroy(ctx->ev_k0);
  if (ctx->ev_k1) cudaEventDestroy(ctx->ev_k1);
  if (ctx->own) cudaStreamDestroy(ctx->own);
  delete ctx;
}

const char* dcol_last_error(const dcol_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }

int dcol_set_stream(dcol_ctx* ctx, void* cuda_stream) {
  if (!ctx) return -1;
  ctx->user = (cudaStream_t)cuda_stream;
  return 0;
}

// Page-lock caller-owned host memory (thin wrappers for callers without CUDA bindings: Julia Arrays, numpy arrays).
int dcol_host_register(dcol_ctx* ctx, void* ptr, int64_t bytes) {
  if (!ctx) return -1;
  ctx->err.clear();
  if (!ptr || bytes <= 0) { ctx->err = "dcol_host_register: bad arguments"; return -1; }
  CK(cudaSetDevice(ctx->device));
  cudaError_t e = cudaHostRegister(ptr, (size_t)bytes, cudaHostRegisterPortable);
  if (e != cudaSuccess) {
    (void)cudaGetLastError();      // not sticky: clear it so that the next launch check does not report it
    ctx->err = std::string("dcol_host_register: ") + cudaGetErrorString(e); return -2;
  }
  return 0;
}
int dcol_host_unregister(dcol_ctx* ctx, void* ptr) {
  if (!ctx) return -1;
  ctx->err.clear();
  if (!ptr) { ctx->err = "dcol_host_unregister: bad arguments"; return -1; }
  CK(cudaSetDevice(ctx->device));
  cudaError_t e = cudaHostUnregister(ptr);
  if (e != cudaSuccess) {
    (void)cudaGetLastError();
    ctx->err = std::string("dcol_host_unregister: ") + cudaGetErrorString(e); return -2;
  }
  return 0;
}

// Testing knob: iteration cap of the interior-point loop.  The reference hard-codes 50 (static_solver2.jl:132) and so does
// every production call; a smaller cap lets a test land a pair's convergence exactly on the last allowed step.
int dcol_set_max_iters(dcol_ctx* ctx, int max_iters) {
  if (!ctx) return -1;
  if (max_iters < 1 || max_iters > MAX_ITERS) { ctx->err = "dcol_set_max_iters: 1..50"; return -1; }
  ctx->max_iters = max_iters;
  return 0;
}
// Optional per-pair output of the duality measure mu = s'z/deg at the iteration that ended the solve (the quantity the
// reference's verbose mode prints, static_solver2.jl:179-183).  mu lives where the following queries' outputs live (host
// pointer for DCOL_MEM_HOST calls, device pointer otherwise) and must hold n doubles; NULL switches it off.
int dcol_set_mu_output(dcol_ctx* ctx, double* mu) {
  if (!ctx) return -1;
  ctx->mu_out = mu;
  return 0;
}

// Opt-in warm start (NOT in the reference, which always starts from `initialize`, and excluded from
// parity: iteration counts change, results agree with a cold solve to the solver tolerance).  `state` is DEVICE memory
// of capacity_pairs x dcol_warm_stride(ctx) doubles, zero-initialised by the caller, addressed by the pair index of the
// following queries.  Each solve leaves behind its iterate at the first iteration whose duality measure fell below
// mu_save (<= 0: 1e-2); a later solve of the same pair index with the same two shapes starts from it instead of the
// reference's initial point.  state == NULL switches it off.  dcol_set_shapes invalidates the stride: set the table first.
int64_t dcol_warm_stride(const dcol_ctx* ctx) {
  if (!ctx) return 0;
  int stride = 0;
  for (const dcol_ctx::Tier& TR : ctx->tiers) stride = std::max(stride, TR.max_stride);
  return stride;
}
int dcol_set_warm_start(dcol_ctx* ctx, double* state, int64_t capacity_pairs, double mu_save) {
  if (!ctx) return -1;
  if (state && capacity_pairs <= 0) { ctx->err = "dcol_set_warm_start: capacity_pairs must be positive"; return -1; }
  ctx->warm = state; ctx->warm_cap = state ? capacity_pairs : 0; ctx->warm_mu = mu_save > 0.0 ? mu_save : 1e-2;
  return 0;
}

// Measurement knob: the following host-pointer calls issue their host<->device copies (same staging buffers, streams,
// chunking and events) but no kernels, which gives the copy-only bound of the end-to-end path (bench.py: e2e.copy_bound).
int dcol_set_copy_only(dcol_ctx* ctx, int on) {
  if (!ctx) return -1;
  ctx->copy_only = on != 0;
  return 0;
}

int64_t dcol_launch_count(const dcol_ctx* ctx) { return ctx ? ctx->launches : 0; }
double dcol_last_kernel_ms(const dcol_ctx* ctx) { return ctx ? ctx->last_ms : -1.0; }

int dcol_set_shapes(dcol_ctx* ctx, int n_shapes, const int32_t* type, const int64_t* param_off, const double* params, int64_t n_params) {
  if (!ctx) return -1;
  ctx->err.clear();
  if (n_shapes <= 0 || !type || !param_off || !params) { ctx->err = "dcol_set_shapes: bad arguments"; return -1; }
  CK(cudaSetDevice(ctx->device));
  // the table, the row constants and the launch plan are read by kernels of submissions still in flight on the
  // library's non-blocking streams: join them before anything is overwritten
  { int rc = drain_all(ctx); if (rc) return rc; }
  CK(cudaDeviceSynchronize());
  std::vector<DevShape> hs(n_shapes);
  std::vector<double> rows;
  std::vector<int32_t> cls(n_shapes);
  for (int c = 0; c < NCLS; c++) ctx->max_rows_cls[c] = 0;
  for (int i = 0; i < n_shapes; i++) {
    DevShape& S = hs[i];
    memset(&S, 0, sizeof(S));
    const double* p = params + param_off[i];
    int64_t avail = n_params - param_off[i];
    int c = type_to_cls(type[i]);
    if (c < 0 || param_off[i] < 0 || param_off[i] >= n_params) { ctx->err = "dcol_set_shapes: bad shape type or offset at shape " + std::to_string(i); return -1; }
    {
      // packed length per type: geometry + r_offset[3] + Q_offset[9]  (include/dcol_b200.h)
      int64_t need = 0;
      const int64_t nn = (type[i] == T_POLYTOPE || type[i] == T_POLYGON) ? (int64_t)p[0] : 0;
      switch (type[i]) {
        case T_CAPSULE: case T_CYLINDER: case T_CONE: need = 2 + 12; break;
        case T_SPHERE: need = 1 + 12; break;
        case T_ELLIPSOID: need = 9 + 12; break;
        case T_POLYTOPE: need = 1 + 4 * nn + 12; break;
        case T_POLYGON: need = 2 + 3 * nn + 12; break;
      }
      if (((type[i] == T_POLYTOPE || type[i] == T_POLYGON) && (!(p[0] >= 1.0) || p[0] > 1e6 || p[0] != std::floor(p[0]))) || avail < need) {
        ctx->err = "dcol_set_shapes: parameter block of shape " + std::to_string(i) + " is truncated or has a bad row count"; return -1; }
      for (int64_t k = 0; k < need; k++) if (!std::isfinite(p[k])) {
        ctx->err = "dcol_set_shapes: non-finite parameter at shape " + std::to_string(i); return -1; }
      bool pos = true;
      if (type[i] == T_CAPSULE || type[i] == T_CYLINDER || type[i] == T_CONE) pos = p[0] > 0.0 && p[1] > 0.0;
      if (type[i] == T_CONE) pos = pos && p[1] < 1.5707963267948966;
      if (type[i] == T_SPHERE) pos = p[0] > 0.0;
      if (type[i] == T_POLYGON) pos = p[1 + 3 * nn] > 0.0;
      if (!pos) { ctx->err = "dcol_set_shapes: R, L, H, beta must be positive (beta < pi/2) at shape " + std::to_string(i); return -1; }
    }
    if (rows.size() & 1) rows.push_back(0.0);          // rows of every shape start 16-byte aligned (double2 loads)
    S.cls = c; S.type = type[i]; S.rows_off = (int64_t)rows.size();
    const double* tail = nullptr;
    auto row = [&](std::initializer_list<double> v) { rows.insert(rows.end(), v); };
    switch (type[i]) {
      case T_POLYTOPE: {          // u_i = b_i*alpha - a_i . x~        (problem_matrices.jl:79-88)
        int n = (int)p[0];
        if (n < 1 || avail < 1 + 4 * (int64_t)n + 12) { ctx->err = "dcol_set_shapes: bad polytope at shape " + std::to_string(i); return -1; }
        const double* A = p + 1; const double* b = p + 1 + 3 * n;
        for (int k = 0; k < n; k++) row({-A[3 * k], -A[3 * k + 1], -A[3 * k + 2], b[k]});
        S.n_ort = n; tail = p + 1 + 4 * n;
      } break;
      case T_CAPSULE: case T_CYLINDER: {   // :30-39, :146-160
        double R = p[0], L = p[1];
        row({0, 0, 0, L / 2, -1.0});       // alpha*L/2 - gamma
        row({0, 0, 0, L / 2, 1.0});        // alpha*L/2 + gamma
        S.n_ort = 2;
        if (type[i] == T_CYLINDER) { row({1.0, 0, 0, L / 2, 0}); row({-1.0, 0, 0, L / 2, 0}); S.n_ort = 4; }
        S.soc[0] = R; tail = p + 2;
      } break;
      case T_CONE: {                        // :56-66
        double H = p[0], beta = p[1], tb = std::tan(beta);
        row({-1.0, 0, 0, H / 4});           // alpha*H/4 - x~1
        S.n_ort = 1; S.soc[0] = tb; S.soc[1] = tb * 3 * H / 4; tail = p + 2;
      } break;
      case T_SPHERE: S.n_ort = 0; S.soc[0] = p[0]; tail = p + 1; break;   // :99-113
      case T_ELLIPSOID: S.n_ort = 0; for (int k = 0; k < 9; k++) S.soc[k] = p[k]; tail = p + 9; break;   // :172-182
      case T_POLYGON: {                     // u_i = b_i*alpha - A_i . e   (:124-134)
        int n = (int)p[0];
        if (n < 1 || avail < 2 + 3 * (int64_t)n + 12) { ctx->err = "dcol_set_shapes: bad polygon at shape " + std::to_string(i); return -1; }
        const double* A = p + 1; const double* b = p + 1 + 2 * n;
        for (int k = 0; k < n; k++) row({0, 0, 0, b[k], -A[2 * k], -A[2 * k + 1]});
        S.n_ort = n; S.soc[0] = p[1 + 3 * n]; tail = p + 2 + 3 * n;
      } break;
    }
    for (int k = 0; k < 3; k++) S.r_off[k] = tail[k];
    for (int k = 0; k < 9; k++) S.Q_off[k] = tail[3 + k];
    bool ident = true;
    for (int k = 0; k < 3; k++) ident = ident && tail[k] == 0.0;
    for (int k = 0; k < 9; k++) ident = ident && tail[3 + k] == ((k % 4 == 0) ? 1.0 : 0.0);
    S.has_offset = ident ? 0 : 1;
    cls[i] = c;
    ctx->max_rows_cls[c] = std::max(ctx->max_rows_cls[c], S.n_ort);
  }
  if (rows.empty()) rows.push_back(0.0);
  // dense sort keys ordered by (unordered class pair, shape1, shape2) when the table is small enough
  ctx->use_keymap = (int64_t)n_shapes * n_shapes <= MAX_KEYS;
  std::vector<int32_t> keymap;
  ctx->key_cp.clear();
  if (ctx->use_keymap) {
    std::vector<std::pair<int, int>> order;
    for (int a = 0; a < n_shapes; a++) for (int b = 0; b < n_shapes; b++) order.push_back({a, b});
    auto canon = [&](const std::pair<int, int>& u) {
      return std::min(cls[u.first], cls[u.second]) * NCLS + std::max(cls[u.first], cls[u.second]); };
    std::stable_sort(order.begin(), order.end(), [&](const std::pair<int, int>& u, const std::pair<int, int>& v) { return canon(u) < canon(v); });
    keymap.resize((size_t)n_shapes * n_shapes);
    for (size_t k = 0; k < order.size(); k++) {
      keymap[(size_t)order[k].first * n_shapes + order[k].second] = (int32_t)k;
      ctx->key_cp.push_back(canon(order[k]));
    }
    ctx->n_keys = (int)order.size();
  } else {
    ctx->n_keys = NCLS * NCLS;
    for (int k = 0; k < ctx->n_keys; k++) ctx->key_cp.push_back(k);
  }
  CK(ctx->d_shapes.need(sizeof(DevShape) * n_shapes));
  CK(ctx->d_rows.need(sizeof(double) * rows.size()));
  CK(ctx->d_cls.need(sizeof(int32_t) * n_shapes));
  CK(cudaMemcpy(ctx->d_shapes.p, hs.data(), sizeof(DevShape) * n_shapes, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(ctx->d_rows.p, rows.data(), sizeof(double) * rows.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(ctx->d_cls.p, cls.data(), sizeof(int32_t) * n_shapes, cudaMemcpyHostToDevice));
  {
    std::vector<double> brad(n_shapes);
    for (int i = 0; i < n_shapes; i++) brad[i] = bounding_radius(type[i], params + param_off[i]);
    CK(ctx->d_brad.need(sizeof(double) * n_shapes));
    CK(cudaMemcpy(ctx->d_brad.p, brad.data(), sizeof(double) * n_shapes, cudaMemcpyHostToDevice));
  }
  if (ctx->use_keymap) {
    CK(ctx->d_keymap.need(sizeof(int32_t) * keymap.size()));
    CK(cudaMemcpy(ctx->d_keymap.p, keymap.data(), sizeof(int32_t) * keymap.size(), cudaMemcpyHostToDevice));
  }
  ctx->h_shapes = hs;
  ctx->n_shapes = n_shapes;

  // ---- static launch plan: class pairs present in the table, grouped into shared-memory tiers by the orthant-row
  // capacity a pair of that class pair can need; inside a tier the heaviest class pairs are queued first
  ctx->tiers.clear();
  {
    // Row capacities of the shared-memory tiers.  Up to 16 rows per pair two 4-warp CTAs fit an SM (the register limit's
    // 8 warps); the <= 8 and <= 16 row tiers stay separate launches all the same: the smaller row store leaves the SM a
    // larger L1 for the spilled registers (measured -1 % proximity, -2.5 % Jacobian against one merged tier).
    const int cap_max = (MAX_SMEM - 1024) / (NFIELD * (int)sizeof(double) * 32) - 8;
    std::vector<int> tier_cap = {8, 16, 32, 64, 128};
    if (const char* e = getenv("DCOL_TIERS")) {          // tuning knob: comma-separated row capacities
      tier_cap.clear();
      for (const char* q = e; *q;) { int v = atoi(q); if (v > 0 && v < cap_max) tier_cap.push_back(v); while (*q && *q != ',') q++; if (*q == ',') q++; }
      std::sort(tier_cap.begin(), tier_cap.end());
    }
    tier_cap.push_back(cap_max);
    const int NTIER = (int)tier_cap.size();
    bool present[NCLS] = {};
    for (int i = 0; i < n_shapes; i++) present[cls[i]] = true;
    // first key of every class pair (keys are class-pair major); key_first[cp+1] closes the range
    std::vector<int> key_first(NCLS * NCLS + 1, ctx->n_keys);
    for (int k = ctx->n_keys - 1; k >= 0; k--) key_first[ctx->key_cp[k]] = k;
    for (int cp = NCLS * NCLS - 1; cp >= 0; cp--) if (key_first[cp] == ctx->n_keys || key_first[cp] > key_first[cp + 1]) key_first[cp] = key_first[cp + 1];
    ctx->has_small = false;
    for (int tier = 0; tier <= NTIER; tier++) {
      // tier == NTIER: the merged tier of small batches (every class pair present, row store of the largest)
      const bool merged = (tier == NTIER);
      if (merged && ctx->tiers.size() < 2) break;
      dcol_ctx::Tier T;
      std::vector<std::pair<int, int>> members;   // (weight, cp)
      for (int c1 = 0; c1 < NCLS; c1++) for (int c2 = c1; c2 < NCLS; c2++) {      // unordered class pairs
        if (!present[c1] || !present[c2]) continue;
        int need = std::max(1, ctx->max_rows_cls[c1] + ctx->max_rows_cls[c2]);
        if (need > tier_cap[NTIER - 1]) { ctx->err = "dcol_set_shapes: too many polytope faces / polygon sides for one pair (shared memory)"; ctx->n_shapes = 0; return -3; }
        if (!merged && !(need <= tier_cap[tier] && (tier == 0 || need > tier_cap[tier - 1]))) continue;
        int weight = need + 6 * (cls_d(c1) + cls_d(c2)) + 8 * (cls_ne(c1) + cls_ne(c2));
        members.push_back({-weight, c1 * NCLS + c2});
        T.rows_cap = std::max(T.rows_cap, need);
      }
      if (members.empty()) continue;
      std::sort(members.begin(), members.end());
      for (auto& m : members) {
        int k = T.n_bins++;
        T.cp[k] = m.second; T.key_begin[k] = key_first[m.second]; T.key_end[k] = key_first[m.second + 1];
        const int c1 = m.second / NCLS, c2 = m.second % NCLS;
        T.ov_stride[k] = 3 + 4 + cls_ne(c1) + cls_ne(c2) + 2 * (ctx->max_rows_cls[c1] + ctx->max_rows_cls[c2]) + 2 * (cls_d(c1) + cls_d(c2));
        T.max_stride = std::max(T.max_stride, T.ov_stride[k]);
      }
      T.pitch = (POSE_SLOTS + T.rows_cap * NFIELD) | 1; // pose slots + rows; odd pitch: conflict-free 64-bit shared-memory accesses
      // 8 warps per SM (255 registers per thread).  Two CTAs of 4 warps when the row stores of both fit: each CTA is
      // its own lockstep group, so one computes while the other sits in the exposed latency of a tile's first gathers
      // and last scatters (measured -3..5 % against one CTA of 8 warps; 4 x 2 and 8 x 1 warps are slower again: fewer
      // warps share each fetched instruction line).  Larger row stores: one CTA with as many warps as fit.
      const int fit = std::max(1, std::min(DCOL_MAXTHREADS / 32, (MAX_SMEM - 1024) / (T.pitch * (int)sizeof(double) * 32)));
      int cand[2] = {fit, fit}, need_occ[2] = {0, 0}, n_cand = 1;
      if (const char* e = getenv("DCOL_WARPS")) cand[0] = std::max(1, std::min(fit, atoi(e)));   // tuning knob
      else if (fit >= 8) { cand[0] = 4; need_occ[0] = 2; n_cand = 2; }
      for (int a = 0; a < n_cand; a++) {
        T.block = 32 * cand[a];
        T.smem = (size_t)T.pitch * sizeof(double) * T.block;
        int min_occ = 1 << 30;
        for (int jac = 0; jac < 3; jac++) {
          const void* kfn = persistent_kernel_ptr(jac);
          if (T.smem > 48 * 1024) CK(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(MAX_SMEM - 1024)));
          int occ = 0;
          CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kfn, T.block, T.smem));
          T.grid[jac] = std::max(1, occ) * ctx->n_sm;
          min_occ = std::min(min_occ, occ);
        }
        if (min_occ >= need_occ[a]) break;      // 4-warp CTAs only if two of them are resident per SM
      }
      if (merged) {
        if (fit >= 4) {                          // four warps of the merged row store fit: 128-pair tiles as in the regular tiers
          T.block = 128; T.smem = (size_t)T.pitch * sizeof(double) * T.block;
          for (int jac = 0; jac < 3; jac++) {
            const void* kfn = persistent_kernel_ptr(jac);
            if (T.smem > 48 * 1024) CK(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(MAX_SMEM - 1024)));
            int occ = 0;
            CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kfn, T.block, T.smem));
            T.grid[jac] = std::max(1, occ) * ctx->n_sm;
          }
          ctx->small_tier = T; ctx->has_small = true;
        }
      } else ctx->tiers.push_back(T);
    }
    if (const char* e = getenv("DCOL_SMALL_MERGED")) if (atoi(e) == 0) ctx->has_small = false;   // tuning knob
    // device copy of the plan: one LaunchTable template per tier (key ranges in begin/end, CTA size in n_tiles, row pitch in P);
    // the merged small-batch tier, if any, sits behind the regular ones
    std::vector<LaunchTable> plan(std::max<size_t>(1, ctx->tiers.size()) + 1);
    for (size_t t = 0; t < ctx->tiers.size() + (ctx->has_small ? 1 : 0); t++) {
      const dcol_ctx::Tier& TR = (t < ctx->tiers.size()) ? ctx->tiers[t] : ctx->small_tier;
      LaunchTable& T = plan[t];
      memset(&T, 0, sizeof(T));
      T.P.row_pitch = TR.pitch; T.n_bins = TR.n_bins; T.rows_cap = TR.rows_cap; T.n_tiles = TR.block;
      for (int k = 0; k < TR.n_bins; k++) { T.cp[k] = TR.cp[k]; T.begin[k] = TR.key_begin[k]; T.end[k] = TR.key_end[k]; T.ov_stride[k] = TR.ov_stride[k]; }
    }
    CK(ctx->d_plan.need(sizeof(LaunchTable) * plan.size()));
    CK(cudaMemcpy(ctx->d_plan.p, plan.data(), sizeof(LaunchTable) * plan.size(), cudaMemcpyHostToDevice));
  }
  // a cudaMemcpy from pageable memory may return once the data is staged; the solver launches go to non-blocking
  // streams that are not ordered behind the legacy stream, so make the uploads visible before returning
  CK(cudaDeviceSynchronize());
  return 0;
}

// DFMA throughput of the device: `seconds` of back-to-back launches of an FMA-only kernel (8 independent chains per
// thread, 8 CTAs of 256 threads per SM), median launch.  bench.py samples the SM clock while this runs.
// three_source = 0: x = fma(x, a, x), the pipe's peak; 1: x = fma(x, a, b), a stream of general DFMAs (k_dfma above).
double dcol_measure_fp64_tflops_mode(dcol_ctx* ctx, double seconds, int three_source) {
  if (!ctx) return -1.0;
  if (cudaSetDevice(ctx->device) != cudaSuccess) return -1.0;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, ctx->device) != cudaSuccess) return -1.0;
  int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 20000;
  double* d = nullptr;
  if (cudaMalloc(&d, sizeof(double) * blocks * threads) != cudaSuccess) return -1.0;
  cudaStream_t s = ctx->stream();
  // (iteration counts are multiples of 16; x grows by (1 + 1e-9) per step in the two-source form: no overflow)
  auto launch = [&](int it) {
    if (three_source) k_dfma<false><<<blocks, threads, 0, s>>>(d, it, 1.0000001, 1e-9);
    else k_dfma<true><<<blocks, threads, 0, s>>>(d, it, 1e-9, 0.0);
  };
  launch(2000);
  cudaStreamSynchronize(s);
  std::vector<float> ms_all;
  double total = 0.0;
  while (total < seconds * 1e3 && ms_all.size() < 4000) {
    cudaEventRecord(ctx->ev_k0, s);
    launch(iters);
    cudaEventRecord(ctx->ev_k1, s);
    cudaEventSynchronize(ctx->ev_k1);
    float ms = 0; cudaEventElapsedTime(&ms, ctx->ev_k0, ctx->ev_k1);
    ms_all.push_back(ms); total += ms; ctx->launches++;
  }
  cudaFree(d);
  if (ms_all.empty()) return -1.0;
  std::sort(ms_all.begin(), ms_all.end());
  const double med = ms_all[ms_all.size() / 2];
  double flops = 2.0 * 8.0 * (double)iters * blocks * threads;
  return flops / (med * 1e-3) / 1e12;
}
double dcol_measure_fp64_tflops_for(dcol_ctx* ctx, double seconds) { return dcol_measure_fp64_tflops_mode(ctx, seconds, 0); }
double dcol_measure_fp64_tflops(dcol_ctx* ctx) { return dcol_measure_fp64_tflops_for(ctx, 0.6); }

}  // extern "C"

// ---------------------------------------------------------------------------
// query driver
// ---------------------------------------------------------------------------
static size_t align256(size_t b) { return (b + 255) & ~(size_t)255; }
static HostCopier* host_copier(dcol_ctx* ctx) {     // copy threads, created on first use
  if (!ctx->copier) {
    int nth = ctx->copy_threads > 0 ? ctx->copy_threads : (int)std::min(8u, std::max(2u, std::thread::hardware_concurrency() / 2));
    if (const char* e = getenv("DCOL_HOST_THREADS")) nth = std::max(1, std::min(64, atoi(e)));
    ctx->copier = new HostCopier();
    ctx->copier->start(nth - 1);                    // the calling thread copies too
  }
  return ctx->copier;
}

static bool host_unpinned(const void* q) {
  cudaPointerAttributes pa;
  if (cudaPointerGetAttributes(&pa, q) != cudaSuccess) { (void)cudaGetLastError(); return true; }
  return pa.type == cudaMemoryTypeUnregistered;
}
static bool host_pipe_on(dcol_ctx* ctx) {
  if (ctx->host_pipe < 0) { const char* e = getenv("DCOL_HOST_PIPE"); ctx->host_pipe = (e && atoi(e) == 0) ? 0 : 1; }
  return ctx->host_pipe != 0;
}

// Results of an operator (device scratch) to the caller's HOST arrays, then join the stream.  items: {host destination,
// device source, bytes}.  Page-locked destinations and small totals: plain asynchronous copies.  Pageable destinations
// above 8 MB: the bytes travel in up to 8 slabs through the pinned staging area, and the copy threads move slab k to the
// caller's arrays while slab k+1 is still crossing PCIe (the same pipeline run_query uses for its chunks).
static int results_to_host(dcol_ctx* ctx, cudaStream_t S, const std::vector<CopyJob>& items) {
  size_t total = 0; bool unpinned = false;
  for (const CopyJob& it : items) { total += it.bytes; unpinned = unpinned || host_unpinned(it.dst); }
  size_t lim_mb = 8192;
  if (const char* e = getenv("DCOL_PIN_MAX_MB")) lim_mb = (size_t)std::max(0, atoi(e));
  bool pipe = host_pipe_on(ctx) && unpinned && total > (8u << 20) && total <= (lim_mb << 20) && !ctx->copy_only;
  if (pipe && ctx->h_pin_cap < total) {
    if (ctx->h_pin) cudaFreeHost(ctx->h_pin);
    ctx->h_pin = nullptr; ctx->h_pin_cap = 0;
    if (cudaHostAlloc((void**)&ctx->h_pin, total + total / 8, cudaHostAllocDefault) == cudaSuccess) ctx->h_pin_cap = total + total / 8;
    else { (void)cudaGetLastError(); ctx->h_pin = nullptr; pipe = false; }
  }
  if (!pipe) {
    for (const CopyJob& it : items) CK(cudaMemcpyAsync(it.dst, it.src, it.bytes, cudaMemcpyDeviceToHost, S));
    CK(cudaStreamSynchronize(S));
    return 0;
  }
  HostCopier* cp = host_copier(ctx);
  dcol_ctx::Slot& L = ctx->slots[0];                    // operators run with every submission drained: slot 0's events are free
  const size_t slab = std::max<size_t>(4u << 20, (total + dcol_ctx::MAX_CHUNKS - 1) / dcol_ctx::MAX_CHUNKS);
  // cut the concatenated byte range [0, total) into slabs; a slab may span several items
  struct Piece { char* host; const char* dev; size_t stage_off, bytes; int slab; };
  std::vector<Piece> pieces;
  size_t pos = 0;
  for (const CopyJob& it : items)
    for (size_t o = 0; o < it.bytes;) {
      const size_t room = slab - (pos % slab), take = std::min(room, it.bytes - o);
      pieces.push_back({it.dst + o, it.src + o, pos, take, (int)(pos / slab)});
      o += take; pos += take;
    }
  const int n_slabs = (int)((total + slab - 1) / slab);
  size_t ip = 0;
  for (int k = 0; k < n_slabs; k++) {
    for (; ip < pieces.size() && pieces[ip].slab == k; ip++)
      CK(cudaMemcpyAsync(ctx->h_pin + pieces[ip].stage_off, pieces[ip].dev, pieces[ip].bytes, cudaMemcpyDeviceToHost, S));
    CK(cudaEventRecord(L.ev_out[k], S));
  }
  std::vector<CopyJob> jobs;
  ip = 0;
  for (int k = 0; k < n_slabs; k++) {
    CK(cudaEventSynchronize(L.ev_out[k]));
    jobs.clear();
    for (; ip < pieces.size() && pieces[ip].slab == k; ip++) HostCopier::add(jobs, pieces[ip].host, ctx->h_pin + pieces[ip].stage_off, pieces[ip].bytes);
    cp->run(jobs);
  }
  CK(cudaStreamSynchronize(S));
  return 0;
}

// pose1/pose2 == NULL: r*/att* have one row per pair.  Otherwise they are tables of np1 / np2 rows and pair k uses
// rows pose1[k] / pose2[k].  in_mem: where the INPUT arrays live (lets the scene operator feed device-generated
// pair lists while the outputs still go to host buffers).
static int run_query(dcol_ctx* ctx, int query, int64_t n, const int32_t* shape1, const int32_t* shape2,
                     const double* r1, const double* att1, const double* r2, const double* att2,
                     int att_kind, double pdip_tol, int mem,
                     double* alpha, double* x, double* deriv, int32_t* status, int32_t* iters,
                     const int32_t* pose1 = nullptr, const int32_t* pose2 = nullptr, int64_t np1 = 0, int64_t np2 = 0,
                     int in_mem = -1, int slot = 0, bool async = false, const int32_t* oidx = nullptr) {
  if (!ctx) return -1;
  if (in_mem < 0) in_mem = mem;
  dcol_ctx::Slot& L = ctx->slots[slot];
  // a slot about to be reused: join the submission it still holds; its outcome stays retrievable by ticket (dcol_wait)
  if (L.pending) { int rc = finish_query(ctx, slot); if (rc && !async) return rc; }
  L.ticket = -1;
  const bool indexed = (pose1 != nullptr);
  ctx->err.clear();
  ctx->last_ms = -1.0;
  if (ctx->n_shapes == 0) { ctx->err = "dcol: shape table not set (call dcol_set_shapes first)"; return -1; }
  if (n < 0 || n > 0x7fffffff) { ctx->err = "dcol: n out of range"; return -1; }
  if (att_kind != DCOL_ATT_QUAT && att_kind != DCOL_ATT_MRP) { ctx->err = "dcol: att_kind must be 0 (quaternion) or 1 (MRP)"; return -1; }
  if (mem != DCOL_MEM_HOST && mem != DCOL_MEM_DEVICE) { ctx->err = "dcol: mem must be 0 (host) or 1 (device)"; return -1; }
  if (n == 0) return 0;
  if (!shape1 || !shape2 || !r1 || !att1 || !r2 || !att2 || !alpha) { ctx->err = "dcol: NULL input/output pointer"; return -1; }
  if (oidx && (mem != DCOL_MEM_DEVICE || in_mem == DCOL_MEM_HOST)) { ctx->err = "dcol: result indirection needs device-resident pairs and outputs"; return -1; }
  if (indexed && (!pose2 || np1 <= 0 || np2 <= 0)) { ctx->err = "dcol: indexed query needs pose1, pose2 and the pose table sizes"; return -1; }
  if (query == Q_PROX && !x) { ctx->err = "dcol_proximity: x is NULL"; return -1; }
  if (query != Q_PROX && !deriv) { ctx->err = "dcol: derivative output is NULL"; return -1; }
  if (query == Q_JAC && !x) { ctx->err = "dcol_proximity_jacobian: x is NULL"; return -1; }
  if (query == Q_JAC && mem == DCOL_MEM_DEVICE && ((uintptr_t)deriv & 15)) { ctx->err = "dcol_proximity_jacobian: J must be 16-byte aligned"; return -1; }
  CK(cudaSetDevice(ctx->device));
  const int na = (att_kind == DCOL_ATT_QUAT) ? 4 : 3;
  const int nt = 6 + 2 * na;
  const size_t dsz = (query == Q_GRAD) ? (size_t)nt : (query == Q_JAC ? (size_t)4 * nt : 0);

  // ---- chunking: device-pointer calls run as one chunk; host-pointer calls are split so that the H2D copy of
  // chunk k+1 and the D2H copy of chunk k-1 overlap the kernels of chunk k (three streams)
  // Large synchronous calls on PAGEABLE caller memory run the library's own staging pipeline (HostCopier): per chunk,
  // copy threads move the caller's inputs into pinned memory ahead of the H2D copy and the results out of pinned memory
  // behind the D2H copy, overlapping the GPU work of the neighbouring chunks.  Page-locked arrays (cudaHostAlloc,
  // dcol_host_register) keep the direct path.
  constexpr size_t STAGE_MAX = 8u << 20;      // below this, one packed staging copy per direction (the small-call path further down)
  constexpr size_t STAGE_THREADS_MIN = 1u << 20;   // ... filled by the copy threads from this size on
  bool pageable = false;
  {
    const size_t approx = (size_t)n * (8 + 8 * (6 + 2 * (size_t)na) + 8 + 24 + 8 * dsz + 8 + (ctx->mu_out ? 8 : 0));
    if (host_pipe_on(ctx) && !async && mem == DCOL_MEM_HOST && !ctx->copy_only && approx > STAGE_MAX) {
      auto unpinned = host_unpinned;
      pageable = unpinned(alpha) || (deriv && unpinned(deriv)) || (x && unpinned(x));
      if (in_mem == DCOL_MEM_HOST) pageable = pageable || unpinned(r1) || unpinned(att2);   // (device-generated pair lists: results only)
    }
  }
  int nchunks = 1;
  if ((mem == DCOL_MEM_HOST || in_mem == DCOL_MEM_HOST) && !async) {      // asynchronous submissions overlap with each other instead
    int64_t target = pageable ? 200000 : 350000;
    if (const char* e = getenv("DCOL_CHUNK")) target = std::max<int64_t>(1024, atoll(e));
    nchunks = (int)std::max<int64_t>(1, std::min<int64_t>(dcol_ctx::MAX_CHUNKS, (n + target / 2) / target));
  }
  const int nk = ctx->n_keys;
  const int nt_tiers = (int)ctx->tiers.size();
  constexpr int MAX_LEVELS = 3;                                     // resume passes after the first pass
  const size_t hist_ints = 3 * (size_t)nk + 16 + (2 + MAX_LEVELS) * NSIDE;   // hist | off | cursor | bad | tile counters per pass and tier
  const size_t hist_stride = (hist_ints + 63) & ~(size_t)63;
  CK(L.d_hist.need(sizeof(int) * hist_stride * nchunks));
  const bool small = (n <= SMALL_MAX) && nchunks == 1;      // single-block preparation, no hand-over, direct result stores
  int max_block = 128;
  for (int t = 0; t < nt_tiers; t++) max_block = std::max(max_block, ctx->tiers[t].block);
  CK(L.d_recs.need(sizeof(PairRec) * ((size_t)n + max_block)));      // one tile of slack behind the last record
  CK(L.d_inv.need(sizeof(int32_t) * n));
  if (!small) CK(L.d_sorted.need((size_t)n * (8 + 24 + 4 + 4 + 8) + 5 * 256));
  CK(L.d_tables.need(sizeof(LaunchTable) * (1 + MAX_LEVELS) * std::max(1, nt_tiers) * nchunks));
  // straggler overflow area per chunk (upper bound: every tile hands over spill_t pairs of the widest record)
  int spill_div = 8;
  if (const char* e = getenv("DCOL_SPILL")) spill_div = atoi(e);          // 0 disables the hand-over; default tile/8
  int levels = 1;                                                           // resume passes (each may hand over again, except the last)
  if (const char* e = getenv("DCOL_LEVELS")) levels = std::max(1, std::min(MAX_LEVELS, atoi(e)));
  const int64_t chunk_max = (nchunks > 1) ? (n + nchunks - 2) / (nchunks - 1) : n;
  size_t ov_doubles = 0;
  for (int t = 0; t < nt_tiers && !small; t++) {
    const dcol_ctx::Tier& TR = ctx->tiers[t];
    const int spill_t = spill_div > 0 ? TR.block / spill_div : 0;
    ov_doubles += (size_t)(chunk_max / TR.block + TR.n_bins + 1) * spill_t * TR.max_stride;
  }
  ov_doubles = (ov_doubles + 31) & ~(size_t)31;
  CK(L.d_ov.need(sizeof(double) * std::max<size_t>(ov_doubles, 32) * 2 * nchunks));

  size_t o_s1 = 0, o_s2 = 0, o_r1 = 0, o_a1 = 0, o_r2 = 0, o_a2 = 0;
  size_t o_alpha = 0, o_x = 0, o_der = 0, o_st = 0, o_it = 0, o_mu = 0;
  char* di = nullptr; char* dout = nullptr;
  size_t o_p1 = 0, o_p2 = 0;
  const int64_t rows1 = indexed ? np1 : n, rows2 = indexed ? np2 : n;
  if (in_mem == DCOL_MEM_HOST) {
    size_t b = 0;
    o_s1 = b; b += align256(sizeof(int32_t) * n);
    o_s2 = b; b += align256(sizeof(int32_t) * n);
    o_p1 = b; b += align256(sizeof(int32_t) * (indexed ? n : 0));
    o_p2 = b; b += align256(sizeof(int32_t) * (indexed ? n : 0));
    o_r1 = b; b += align256(sizeof(double) * 3 * rows1);
    o_a1 = b; b += align256(sizeof(double) * na * rows1);
    o_r2 = b; b += align256(sizeof(double) * 3 * rows2);
    o_a2 = b; b += align256(sizeof(double) * na * rows2);
    CK(L.d_in.need(b));
    di = (char*)L.d_in.p;
  }
  if (mem == DCOL_MEM_HOST) {
    size_t b = 0;
    o_alpha = b; b += align256(sizeof(double) * n);
    o_x = b; b += align256(sizeof(double) * 3 * n);
    o_der = b; b += align256(sizeof(double) * dsz * n);
    o_st = b; b += align256(sizeof(int32_t) * n);
    o_it = b; b += align256(sizeof(int32_t) * n);
    o_mu = b; b += align256(sizeof(double) * (ctx->mu_out ? n : 0));
    CK(L.d_out.need(b));
    dout = (char*)L.d_out.p;
  }
  // ---- small synchronous host calls: pack the inputs into one pinned staging area (host memcpy), one H2D copy, one D2H
  // copy of all results, unpack on the host.  Pageable caller memory would otherwise cost one driver-staged copy per array.
  size_t b_in = 0, b_out = 0;
  if (in_mem == DCOL_MEM_HOST) b_in = o_a2 + align256(sizeof(double) * na * rows2);
  if (mem == DCOL_MEM_HOST) b_out = o_mu + align256(sizeof(double) * (ctx->mu_out ? n : 0));
  const bool staged = !async && nchunks == 1 && mem == DCOL_MEM_HOST && in_mem == DCOL_MEM_HOST && b_in + b_out <= STAGE_MAX;
  if (staged) {
    if (ctx->h_stage_cap < 2 * STAGE_MAX) {
      if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
      ctx->h_stage = nullptr; ctx->h_stage_cap = 0;
      CK(cudaHostAlloc((void**)&ctx->h_stage, 2 * STAGE_MAX, cudaHostAllocDefault));
      ctx->h_stage_cap = 2 * STAGE_MAX;
    }
    char* hs = ctx->h_stage;
    std::vector<CopyJob> J;
    HostCopier::add(J, hs + o_s1, shape1, sizeof(int32_t) * n); HostCopier::add(J, hs + o_s2, shape2, sizeof(int32_t) * n);
    if (indexed) { HostCopier::add(J, hs + o_p1, pose1, sizeof(int32_t) * n); HostCopier::add(J, hs + o_p2, pose2, sizeof(int32_t) * n); }
    HostCopier::add(J, hs + o_r1, r1, sizeof(double) * 3 * rows1); HostCopier::add(J, hs + o_a1, att1, sizeof(double) * na * rows1);
    HostCopier::add(J, hs + o_r2, r2, sizeof(double) * 3 * rows2); HostCopier::add(J, hs + o_a2, att2, sizeof(double) * na * rows2);
    if (b_in + b_out >= STAGE_THREADS_MIN && ctx->host_pipe) host_copier(ctx)->run(J);      // megabytes: the copy threads help
    else for (const CopyJob& j : J) memcpy(j.dst, j.src, j.bytes);
  }
  // ---- pageable-memory pipeline: pinned staging for the whole call (inputs | results), copy threads on first use
  char* hp_in = nullptr; char* hp_out = nullptr;
  if (pageable) {
    size_t lim_mb = 8192;
    if (const char* e = getenv("DCOL_PIN_MAX_MB")) lim_mb = (size_t)std::max(0, atoi(e));
    const size_t want = align256(b_in) + align256(b_out);
    if (want > (lim_mb << 20)) pageable = false;
    else if (ctx->h_pin_cap < want) {
      if (ctx->h_pin) cudaFreeHost(ctx->h_pin);
      ctx->h_pin = nullptr; ctx->h_pin_cap = 0;
      const size_t cap = want + want / 8;
      if (cudaHostAlloc((void**)&ctx->h_pin, cap, cudaHostAllocDefault) == cudaSuccess) ctx->h_pin_cap = cap;
      else { (void)cudaGetLastError(); ctx->h_pin = nullptr; pageable = false; }      // no pinned memory to spare: the driver's pageable path
    }
    if (pageable) (void)host_copier(ctx);
    if (pageable) { hp_in = ctx->h_pin; hp_out = ctx->h_pin + align256(b_in); }
  }
  // host-side source / destination of the copies: the caller's arrays, or their image in the pinned staging area
  const int32_t* hs1 = pageable ? (const int32_t*)(hp_in + o_s1) : shape1;
  const int32_t* hs2 = pageable ? (const int32_t*)(hp_in + o_s2) : shape2;
  const int32_t* hq1 = pageable ? (const int32_t*)(hp_in + o_p1) : pose1;
  const int32_t* hq2 = pageable ? (const int32_t*)(hp_in + o_p2) : pose2;
  const double* hr1 = pageable ? (const double*)(hp_in + o_r1) : r1;
  const double* ha1 = pageable ? (const double*)(hp_in + o_a1) : att1;
  const double* hr2 = pageable ? (const double*)(hp_in + o_r2) : r2;
  const double* ha2 = pageable ? (const double*)(hp_in + o_a2) : att2;
  double* ho_alpha = pageable ? (double*)(hp_out + o_alpha) : alpha;
  double* ho_x = pageable ? (double*)(hp_out + o_x) : x;
  double* ho_der = pageable ? (double*)(hp_out + o_der) : deriv;
  int32_t* ho_st = pageable ? (int32_t*)(hp_out + o_st) : status;
  int32_t* ho_it = pageable ? (int32_t*)(hp_out + o_it) : iters;
  double* ho_mu = pageable ? (double*)(hp_out + o_mu) : ctx->mu_out;
  std::vector<CopyJob> jobs;
  int drained = 0;                              // result chunks already copied out of the staging area
  std::vector<int64_t> cb;                      // chunk boundaries (filled below)
  auto unstage = [&](int k) {                   // results of chunk k: pinned staging -> the caller's arrays
    const int64_t k0 = cb[k], nk_ = cb[k + 1] - cb[k];
    jobs.clear();
    HostCopier::add(jobs, alpha + k0, ho_alpha + k0, sizeof(double) * nk_);
    if (x) HostCopier::add(jobs, x + 3 * k0, ho_x + 3 * k0, sizeof(double) * 3 * nk_);
    if (dsz) HostCopier::add(jobs, deriv + dsz * k0, ho_der + dsz * k0, sizeof(double) * dsz * nk_);
    if (status) HostCopier::add(jobs, status + k0, ho_st + k0, sizeof(int32_t) * nk_);
    if (iters) HostCopier::add(jobs, iters + k0, ho_it + k0, sizeof(int32_t) * nk_);
    if (ctx->mu_out) HostCopier::add(jobs, ctx->mu_out + k0, ho_mu + k0, sizeof(double) * nk_);
    ctx->copier->run(jobs);
  };
  const int jac = query;                       // kernel variant = query kind
  cudaStream_t S = ctx->stream();
  const bool piped = (nchunks > 1) || async;                  // three-stream pipeline (always for asynchronous submissions)
  cudaStream_t SC = piped ? ctx->s_comp : S;                  // compute stream of the pipeline
  if (piped) { CK(cudaEventRecord(ctx->ev_fork, S)); CK(cudaStreamWaitEvent(ctx->s_copy, ctx->ev_fork, 0)); CK(cudaStreamWaitEvent(SC, ctx->ev_fork, 0)); }

  // Tiny calls (the single-pair drop-in calls above all): the kernels read the inputs from and write the results to the pinned
  // staging area directly (zero-copy over PCIe), which saves both copy launches and their latency.  Measured through the
  // C ABI on B200 (profiles/r02_latency.txt): single pair 63.0 -> 52.8 us, 100 pairs 77.8 -> 73.4, 1 000 pairs 97.9 -> 92.5.
  static const int zc_max = getenv("DCOL_ZEROCOPY_MAX") ? atoi(getenv("DCOL_ZEROCOPY_MAX")) : 1024;    // pairs; 0 switches it off
  bool zero_copy = false;
  if (staged && n <= zc_max) {
    void* dp = nullptr;
    if (cudaHostGetDevicePointer(&dp, ctx->h_stage, 0) == cudaSuccess && dp) { di = (char*)dp; dout = (char*)dp + STAGE_MAX; zero_copy = true; }
    else (void)cudaGetLastError();
  }
  // A handful of pairs on host pointers (above all ONE pair: the drop-in calls of the reference's examples): the host knows
  // the shapes' classes, so it packs the 128-byte records itself into the staging area -- grouped by class pair, with the
  // exchange of the two primitives the packing kernels apply -- and ONE launch (proximity_packed: a block per class pair
  // present) reads them from there and writes the results back the same way.  No preparation kernel, no launch table.
  // Through the C ABI on B200: single pair 50 -> 38 us (profiles/r02_latency.txt).
  static const int packed_max = getenv("DCOL_PACKED_MAX") ? std::min(DCOL_PACKED_MAX, atoi(getenv("DCOL_PACKED_MAX"))) : DCOL_PACKED_MAX;   // tuning knob; 0 = off
  const PairRec* packed_recs = nullptr; PackedGroups packed; int packed_block = 32;
  packed.n = 0;
  if (zero_copy && n <= packed_max && !indexed && !oidx && (ctx->has_small || nt_tiers == 1) && !ctx->copy_only && !ctx->warm) {
    int cpk[DCOL_PACKED_MAX], cnt[MAX_BINS] = {};
    for (int64_t i = 0; i < n; i++) {
      const int s1 = shape1[i], s2 = shape2[i];
      if ((unsigned)s1 >= (unsigned)ctx->n_shapes || (unsigned)s2 >= (unsigned)ctx->n_shapes) { ctx->err = "dcol: shape index out of range in shape1/shape2"; return -1; }
      const int c1 = ctx->h_shapes[s1].cls, c2 = ctx->h_shapes[s2].cls;
      cpk[i] = std::min(c1, c2) * NCLS + std::max(c1, c2);
      cnt[cpk[i]]++;
    }
    int first[MAX_BINS];
    for (int cp = 0, at = 0; cp < MAX_BINS; cp++) if (cnt[cp]) {
      packed.cp[packed.n] = cp; packed.first[packed.n] = at; packed.count[packed.n] = cnt[cp];
      first[cp] = at; at += cnt[cp]; packed.n++;
      packed_block = std::max(packed_block, (cnt[cp] + 31) & ~31);
    }
    PairRec* hr = (PairRec*)(ctx->h_stage + align256(b_in));
    for (int64_t i = 0; i < n; i++) {
      const int s1 = shape1[i], s2 = shape2[i];
      const bool swp = ctx->h_shapes[s1].cls > ctx->h_shapes[s2].cls;
      PairRec& R = hr[first[cpk[i]]++];
      const double* ra = (swp ? r2 : r1) + 3 * i; const double* aa = (swp ? att2 : att1) + na * i;
      const double* rb = (swp ? r1 : r2) + 3 * i; const double* ab = (swp ? att1 : att2) + na * i;
      for (int a = 0; a < 3; a++) { R.r1[a] = ra[a]; R.r2[a] = rb[a]; }
      R.a1[3] = 0.0; R.a2[3] = 0.0;
      for (int a = 0; a < na; a++) { R.a1[a] = aa[a]; R.a2[a] = ab[a]; }
      R.s1 = swp ? s2 : s1; R.s2 = swp ? s1 : s2; R.pair = (int32_t)i; R.flags = swp ? DCOL_REC_SWAPPED : 0;
    }
    // (a tier whose row store leaves room for fewer threads per block than the largest group needs keeps the general path)
    if (packed_block <= (ctx->has_small ? ctx->small_tier : ctx->tiers[0]).block) packed_recs = (const PairRec*)(di + align256(b_in));
  }
  if (staged && !zero_copy) CK(cudaMemcpyAsync(di, ctx->h_stage, b_in, cudaMemcpyHostToDevice, S));
  if (indexed && in_mem == DCOL_MEM_HOST && !staged) {       // pose tables go up once, ahead of the first chunk's pair lists
    cudaStream_t SI = piped ? ctx->s_copy : S;
    if (pageable) {
      jobs.clear();
      HostCopier::add(jobs, hp_in + o_r1, r1, sizeof(double) * 3 * np1); HostCopier::add(jobs, hp_in + o_a1, att1, sizeof(double) * na * np1);
      HostCopier::add(jobs, hp_in + o_r2, r2, sizeof(double) * 3 * np2); HostCopier::add(jobs, hp_in + o_a2, att2, sizeof(double) * na * np2);
      ctx->copier->run(jobs);
    }
    CK(cudaMemcpyAsync(di + o_r1, hr1, sizeof(double) * 3 * np1, cudaMemcpyHostToDevice, SI));
    CK(cudaMemcpyAsync(di + o_a1, ha1, sizeof(double) * na * np1, cudaMemcpyHostToDevice, SI));
    CK(cudaMemcpyAsync(di + o_r2, hr2, sizeof(double) * 3 * np2, cudaMemcpyHostToDevice, SI));
    CK(cudaMemcpyAsync(di + o_a2, ha2, sizeof(double) * na * np2, cudaMemcpyHostToDevice, SI));
  }
  // chunk boundaries: a short first chunk lets the kernels start early while the bulk is still crossing PCIe
  for (int c = 0; c < nchunks; c++) L.h_bad[c] = 0;
  cb.assign(nchunks + 1, 0);
  {
    const int64_t first = (nchunks > 1) ? std::min<int64_t>(n / (2 * nchunks), 120000) : n;
    cb[nchunks] = n;
    for (int c = 1; c < nchunks; c++) cb[c] = first + (n - first) * (c - 1) / (nchunks - 1);
    if (nchunks > 1) cb[1] = first;
  }
  for (int c = 0; c < nchunks; c++) {
    const int64_t c0 = cb[c], c1 = cb[c + 1], nc = c1 - c0;
    QueryParams P;
    memset(&P, 0, sizeof(P));
    P.shapes = (const DevShape*)ctx->d_shapes.p; P.rows = (const double*)ctx->d_rows.p;
    P.att_kind = att_kind; P.pdip_tol = pdip_tol; P.query = query; P.max_iters = ctx->max_iters; P.direct = small ? 1 : 0;
    if (ctx->warm) {      // records are addressed by the caller's pair index of this call
      int stride = 0;
      for (const dcol_ctx::Tier& TR : ctx->tiers) stride = std::max(stride, TR.max_stride);
      P.warm = ctx->warm + (size_t)c0 * stride; P.warm_stride = stride;
      P.warm_cap = (int)std::max<int64_t>(0, std::min<int64_t>(ctx->warm_cap - c0, 0x7fffffff)); P.warm_mu = ctx->warm_mu;
    }
    PairSource src;
    memset(&src, 0, sizeof(src));
    src.na = na;
    src.oidx = oidx ? oidx + c0 : nullptr;
    if (in_mem == DCOL_MEM_HOST) {
      cudaStream_t SI = piped ? ctx->s_copy : S;
      if (pageable) {      // this chunk's inputs: the caller's arrays -> pinned staging (the GPU is busy with the previous chunk meanwhile)
        jobs.clear();
        HostCopier::add(jobs, hp_in + o_s1 + sizeof(int32_t) * c0, shape1 + c0, sizeof(int32_t) * nc);
        HostCopier::add(jobs, hp_in + o_s2 + sizeof(int32_t) * c0, shape2 + c0, sizeof(int32_t) * nc);
        if (indexed) {
          HostCopier::add(jobs, hp_in + o_p1 + sizeof(int32_t) * c0, pose1 + c0, sizeof(int32_t) * nc);
          HostCopier::add(jobs, hp_in + o_p2 + sizeof(int32_t) * c0, pose2 + c0, sizeof(int32_t) * nc);
        } else {
          HostCopier::add(jobs, hp_in + o_r1 + sizeof(double) * 3 * c0, r1 + 3 * c0, sizeof(double) * 3 * nc);
          HostCopier::add(jobs, hp_in + o_a1 + sizeof(double) * na * c0, att1 + na * c0, sizeof(double) * na * nc);
          HostCopier::add(jobs, hp_in + o_r2 + sizeof(double) * 3 * c0, r2 + 3 * c0, sizeof(double) * 3 * nc);
          HostCopier::add(jobs, hp_in + o_a2 + sizeof(double) * na * c0, att2 + na * c0, sizeof(double) * na * nc);
        }
        ctx->copier->run(jobs);
      }
      if (!staged) {
      CK(cudaMemcpyAsync(di + o_s1 + sizeof(int32_t) * c0, hs1 + c0, sizeof(int32_t) * nc, cudaMemcpyHostToDevice, SI));
      CK(cudaMemcpyAsync(di + o_s2 + sizeof(int32_t) * c0, hs2 + c0, sizeof(int32_t) * nc, cudaMemcpyHostToDevice, SI));
      if (indexed) {
        CK(cudaMemcpyAsync(di + o_p1 + sizeof(int32_t) * c0, hq1 + c0, sizeof(int32_t) * nc, cudaMemcpyHostToDevice, SI));
        CK(cudaMemcpyAsync(di + o_p2 + sizeof(int32_t) * c0, hq2 + c0, sizeof(int32_t) * nc, cudaMemcpyHostToDevice, SI));
      } else {
        CK(cudaMemcpyAsync(di + o_r1 + sizeof(double) * 3 * c0, hr1 + 3 * c0, sizeof(double) * 3 * nc, cudaMemcpyHostToDevice, SI));
        CK(cudaMemcpyAsync(di + o_a1 + sizeof(double) * na * c0, ha1 + na * c0, sizeof(double) * na * nc, cudaMemcpyHostToDevice, SI));
        CK(cudaMemcpyAsync(di + o_r2 + sizeof(double) * 3 * c0, hr2 + 3 * c0, sizeof(double) * 3 * nc, cudaMemcpyHostToDevice, SI));
        CK(cudaMemcpyAsync(di + o_a2 + sizeof(double) * na * c0, ha2 + na * c0, sizeof(double) * na * nc, cudaMemcpyHostToDevice, SI));
      }
      }
      if (piped) { CK(cudaEventRecord(L.ev_in[c], SI)); CK(cudaStreamWaitEvent(SC, L.ev_in[c], 0)); }
      src.shape1 = (const int32_t*)(di + o_s1) + c0; src.shape2 = (const int32_t*)(di + o_s2) + c0;
      if (indexed) {
        src.pose1 = (const int32_t*)(di + o_p1) + c0; src.pose2 = (const int32_t*)(di + o_p2) + c0;
        src.r1 = (const double*)(di + o_r1); src.att1 = (const double*)(di + o_a1);
        src.r2 = (const double*)(di + o_r2); src.att2 = (const double*)(di + o_a2);
      } else {
        src.r1 = (const double*)(di + o_r1) + 3 * c0; src.att1 = (const double*)(di + o_a1) + na * c0;
        src.r2 = (const double*)(di + o_r2) + 3 * c0; src.att2 = (const double*)(di + o_a2) + na * c0;
      }
    } else {
      src.shape1 = shape1 + c0; src.shape2 = shape2 + c0;
      if (indexed) { src.pose1 = pose1 + c0; src.pose2 = pose2 + c0; src.r1 = r1; src.att1 = att1; src.r2 = r2; src.att2 = att2; }
      else { src.r1 = r1 + 3 * c0; src.att1 = att1 + na * c0; src.r2 = r2 + 3 * c0; src.att2 = att2 + na * c0; }
    }
    // where the caller's results of this chunk go (caller order): the caller's device arrays or the D2H staging area
    double* f_alpha; double* f_x; double* f_der; int32_t* f_st; int32_t* f_it; double* f_mu = nullptr;
    if (mem == DCOL_MEM_HOST) {
      f_alpha = (double*)(dout + o_alpha) + c0; f_x = (double*)(dout + o_x) + 3 * c0;
      f_der = dsz ? (double*)(dout + o_der) + dsz * c0 : nullptr;
      f_st = (int32_t*)(dout + o_st) + c0; f_it = (int32_t*)(dout + o_it) + c0;
      if (ctx->mu_out) f_mu = (double*)(dout + o_mu) + c0;
    } else {
      f_alpha = alpha + c0; f_x = x ? x + 3 * c0 : nullptr; f_der = deriv ? deriv + dsz * c0 : nullptr;
      f_st = status ? status + c0 : nullptr; f_it = iters ? iters + c0 : nullptr;
      if (ctx->mu_out) f_mu = ctx->mu_out + c0;
    }
    // the solver stores the per-pair scalars in SORTED order (coalesced); k_unpermute brings them into caller order.
    // Derivative rows (120-488 contiguous bytes per pair) go to their final place directly.
    PairRec* d_recs = (PairRec*)L.d_recs.p + c0;
    int32_t* d_inv = (int32_t*)L.d_inv.p + c0;
    double* s_alpha = nullptr; double* s_x = nullptr; int32_t* s_st = nullptr; int32_t* s_it = nullptr; double* s_mu = nullptr;
    if (small) {
      P.alpha = f_alpha; P.x = f_x; P.status = f_st; P.iters = f_it; P.mu = f_mu;
    } else {
      char* sb = (char*)L.d_sorted.p;
      const size_t nn = (size_t)n;
      s_alpha = (double*)sb + c0; sb += align256(8 * nn);
      s_x = (double*)sb + 3 * c0; sb += align256(24 * nn);
      s_mu = (double*)sb + c0; sb += align256(8 * nn);
      s_st = (int32_t*)sb + c0; sb += align256(4 * nn);
      s_it = (int32_t*)sb + c0;
      P.alpha = s_alpha; P.x = f_x ? s_x : nullptr; P.status = f_st ? s_st : nullptr; P.iters = f_it ? s_it : nullptr; P.mu = f_mu ? s_mu : nullptr;
    }
    P.deriv = f_der;
    P.recs = d_recs;
    // ---- bin the chunk's pairs by key (device counting sort), pack them in sorted order, build the tile tables on the device
    int* d_hist = (int*)L.d_hist.p + hist_stride * c;
    int* d_off = d_hist + nk;            // nk+1
    int* d_cursor = d_off + nk + 1;      // nk
    int* d_bad = L.d_bad_map + c;        // mapped host flag of this chunk
    int* d_counter = d_cursor + nk + 1;  // one tile counter per shared-memory tier and pass
    LaunchTable* d_tab = (LaunchTable*)L.d_tables.p + (size_t)(1 + MAX_LEVELS) * std::max(1, nt_tiers) * c;
    double* d_ovc = (double*)L.d_ov.p + 2 * ov_doubles * c;
    const int32_t* d_keymap = ctx->use_keymap ? (const int32_t*)ctx->d_keymap.p : nullptr;
    bool t0_recorded = false;
    if (ctx->copy_only) {
      // measurement mode: the copies of this call with the same buffers, streams and chunking, and no kernels
    } else if (packed_recs) {
      P.recs = packed_recs;
      CK(cudaEventRecord(L.ev_t0[c], SC)); t0_recorded = true;
    } else if (small) {
      // one preparation launch (which also packs up to SMALL_FUSED pairs) and one solver launch over the merged tier
      const bool fused = nc <= SMALL_FUSED;
      // (nothing may sit between the two kernels for the solver launch to become resident early, see below: the kernel
      // timer of such a call starts ahead of the preparation kernel)
      if (fused && (ctx->has_small || nt_tiers == 1)) { CK(cudaEventRecord(L.ev_t0[c], SC)); t0_recorded = true; }
      const LaunchTable* plan = (const LaunchTable*)ctx->d_plan.p + (ctx->has_small ? nt_tiers : 0);
      // (a handful of pairs: four warps instead of thirty-two reach the block-wide barriers of the scan sooner)
      const int prep_threads = (nc <= 128 && !getenv("DCOL_PREP_1024")) ? 128 : SMALL_THREADS;
      k_small_prep<<<1, prep_threads, sizeof(int) * (nk + 1), SC>>>(src, (int)nc, (const int32_t*)ctx->d_cls.p, d_keymap, ctx->n_shapes, nk,
                                                                      d_off, d_counter, (2 + MAX_LEVELS) * NSIDE, d_bad, d_inv,
                                                                      plan, d_tab, ctx->has_small ? 1 : nt_tiers, d_ovc, P, fused ? d_recs : nullptr);
      ctx->launches += 1;
      if (!fused) {
        k_pack<<<(int)std::min<int64_t>((nc + PACK_THREADS - 1) / PACK_THREADS, 148 * 8), PACK_THREADS, 0, SC>>>(src, nc, (const int32_t*)ctx->d_cls.p, d_inv, d_recs);
        ctx->launches += 1;
      }
    } else {
      CK(cudaMemsetAsync(d_hist, 0, sizeof(int) * hist_ints, SC));
      int threads = 256;
      int blocks = (int)std::min<int64_t>((nc + threads * 8 - 1) / (threads * 8), 1184);
      k_hist<<<blocks, threads, sizeof(int) * nk, SC>>>(src.shape1, src.shape2, nc, (const int32_t*)ctx->d_cls.p, d_keymap, ctx->n_shapes, nk, d_hist, d_bad);
      k_scan_tables<<<1, 1024, 0, SC>>>(d_hist, nk, d_off, d_cursor, (const LaunchTable*)ctx->d_plan.p, d_tab, nt_tiers, d_ovc, P, spill_div);
      k_scatter<<<blocks, threads, sizeof(int) * 2 * nk, SC>>>(src.shape1, src.shape2, nc, (const int32_t*)ctx->d_cls.p, d_keymap, ctx->n_shapes, nk, d_cursor, d_inv);
      k_pack<<<(int)std::min<int64_t>((nc + PACK_THREADS - 1) / PACK_THREADS, 148 * 8), PACK_THREADS, 0, SC>>>(src, nc, (const int32_t*)ctx->d_cls.p, d_inv, d_recs);
      ctx->launches += 4;
    }
    // ---- persistent solver launches, one per shared-memory tier present in the shape table
    if (!t0_recorded) CK(cudaEventRecord(L.ev_t0[c], SC));
    // per tier, on the tier's stream: first pass, then (right behind it) the resume passes that finish the handed-over
    // stragglers in dense tiles — the resume pass of one tier overlaps the first pass of the next
    const bool one_launch = small && ctx->has_small;
    const int n_launch = one_launch ? 1 : nt_tiers;
    if (n_launch > 1) CK(cudaEventRecord(ctx->ev_join[NSIDE - 1], SC));
    // (tiers in ascending row capacity; launching the heavier tier first was measured 0.5-1 % slower, profiles/r02_variants.txt)
    for (int t = 0; t < n_launch && !ctx->copy_only; t++) {
      const dcol_ctx::Tier& TR = one_launch ? ctx->small_tier : ctx->tiers[t];
      cudaStream_t st = (n_launch == 1) ? SC : ctx->side[t % NSIDE];
      if (st != SC) CK(cudaStreamWaitEvent(st, ctx->ev_join[NSIDE - 1], 0));
      // small batches: no more CTAs than there can be tiles (every bin may end in a partial tile)
      const int grid = small ? (int)std::min<int64_t>(TR.grid[jac], nc / TR.block + TR.n_bins + 1) : TR.grid[jac];
      // a small batch's single solver launch sits right behind k_small_prep on the same stream: programmatic stream
      // serialisation lets its blocks become resident while that kernel still runs (they wait at griddepcontrol.wait).
      // Together with the kernel timer's start moved ahead of the preparation kernel (an event record between two kernels
      // costs ~2 us of front-end time): single pair 49.5 -> 45.8 us (device pointers), 1 000 pairs 76 -> 71 us
      static const bool pdl = !(getenv("DCOL_PDL") && atoi(getenv("DCOL_PDL")) == 0);      // tuning knob
      const int early = (pdl && small && n_launch == 1 && nc <= SMALL_FUSED) ? 1 : 0;
      if (packed_recs) {
        P.row_pitch = TR.pitch;
        const size_t smp = (size_t)TR.pitch * sizeof(double) * packed_block;
        CK(jac == 2 ? launch_packed_v2(packed_block, smp, st, P, packed) : (jac == 1 ? launch_packed_v1(packed_block, smp, st, P, packed) : launch_packed_v0(packed_block, smp, st, P, packed)));
      } else
      CK(launch_persistent(jac, grid, TR.block, TR.smem, st, d_tab + t, d_counter + t, early));
      ctx->launches++;
      static const bool skip_resume = getenv("DCOL_SKIP_RESUME") != nullptr;      // timing experiments only: stragglers stay unfinished
      for (int lv = 1; lv <= levels && spill_div > 0 && !small && !skip_resume; lv++) {
        LaunchTable* prev = d_tab + (size_t)(lv - 1) * std::max(1, nt_tiers) + t;
        LaunchTable* cur = d_tab + (size_t)lv * std::max(1, nt_tiers) + t;
        k_build_tables2<<<1, TABLE_THREADS, 0, st>>>(prev, cur, d_ovc + (size_t)(lv & 1) * ov_doubles, lv < levels ? 1 : 0);
        CK(launch_persistent(jac, TR.grid[jac], TR.block, TR.smem, st, cur, d_counter + lv * NSIDE + t));
        ctx->launches += 2;
      }
      if (st != SC) { CK(cudaEventRecord(ctx->ev_join[t % (NSIDE - 1)], st)); CK(cudaStreamWaitEvent(SC, ctx->ev_join[t % (NSIDE - 1)], 0)); }
    }
    CK(cudaEventRecord(L.ev_t1[c], SC));             // dcol_last_kernel_ms brackets the solver launches only (first pass + resume, all tiers)
    if (!small && !ctx->copy_only) {
      const int blocks = (int)std::min<int64_t>((nc + 255) / 256, 148 * 16);
      k_unpermute<<<blocks, 256, 0, SC>>>(nc, d_inv, src.oidx, s_alpha, s_x, s_st, s_it, s_mu, f_alpha, f_x, f_st, f_it, f_mu);
      ctx->launches += 1;
    }
    // ---- results of this chunk back (host mode), overlapping the next chunk's kernels
    if (staged) {
      if (!zero_copy) CK(cudaMemcpyAsync(ctx->h_stage + STAGE_MAX, dout, b_out, cudaMemcpyDeviceToHost, S));
    } else if (mem == DCOL_MEM_HOST) {
      cudaStream_t SO = piped ? ctx->s_out : S;
      if (piped) { CK(cudaEventRecord(L.ev_done[c], SC)); CK(cudaStreamWaitEvent(SO, L.ev_done[c], 0)); }
      CK(cudaMemcpyAsync(ho_alpha + c0, dout + o_alpha + sizeof(double) * c0, sizeof(double) * nc, cudaMemcpyDeviceToHost, SO));
      if (x) CK(cudaMemcpyAsync(ho_x + 3 * c0, dout + o_x + sizeof(double) * 3 * c0, sizeof(double) * 3 * nc, cudaMemcpyDeviceToHost, SO));
      if (dsz) CK(cudaMemcpyAsync(ho_der + dsz * c0, dout + o_der + sizeof(double) * dsz * c0, sizeof(double) * dsz * nc, cudaMemcpyDeviceToHost, SO));
      if (status) CK(cudaMemcpyAsync(ho_st + c0, dout + o_st + sizeof(int32_t) * c0, sizeof(int32_t) * nc, cudaMemcpyDeviceToHost, SO));
      if (iters) CK(cudaMemcpyAsync(ho_it + c0, dout + o_it + sizeof(int32_t) * c0, sizeof(int32_t) * nc, cudaMemcpyDeviceToHost, SO));
      if (ctx->mu_out) CK(cudaMemcpyAsync(ho_mu + c0, dout + o_mu + sizeof(double) * c0, sizeof(double) * nc, cudaMemcpyDeviceToHost, SO));
      if (pageable) {
        CK(cudaEventRecord(L.ev_out[c], SO));
        // results of earlier chunks that have landed meanwhile go to the caller's arrays now (never waits)
        while (drained < c) {
          if (cudaEventQuery(L.ev_out[drained]) != cudaSuccess) { (void)cudaGetLastError(); break; }   // not ready is not an error (a real one is sticky and resurfaces at the join)
          unstage(drained++);
        }
      }
    }
  }
  // ---- the join is finish_query (the only host synchronisation); the bad-shape flags are mapped host memory
  cudaStream_t SL = piped ? ctx->s_out : S;
  if (piped && mem != DCOL_MEM_HOST) { CK(cudaEventRecord(L.ev_done[0], SC)); CK(cudaStreamWaitEvent(SL, L.ev_done[0], 0)); }
  // join point of THIS submission: everything it enqueued precedes ev_end on SL (the compute stream's last work
  // is ordered before it through ev_done), so waiting for it does not wait for later submissions
  if (piped) { CK(cudaEventRecord(L.ev_done[nchunks - 1], SC)); CK(cudaStreamWaitEvent(SL, L.ev_done[nchunks - 1], 0)); }
  CK(cudaEventRecord(L.ev_end, SL));
  L.pending = true; L.nchunks = nchunks; L.piped = piped; L.result = 0;
  if (async) return 0;
  if (pageable) {
    for (; drained < nchunks; drained++) { CK(cudaEventSynchronize(L.ev_out[drained])); unstage(drained); }
  }
  const int rc = finish_query(ctx, slot);
  if (staged && rc == 0) {
    const char* ho = ctx->h_stage + STAGE_MAX;
    std::vector<CopyJob> J;
    HostCopier::add(J, alpha, ho + o_alpha, sizeof(double) * n);
    if (x) HostCopier::add(J, x, ho + o_x, sizeof(double) * 3 * n);
    if (dsz) HostCopier::add(J, deriv, ho + o_der, sizeof(double) * dsz * n);
    if (status) HostCopier::add(J, status, ho + o_st, sizeof(int32_t) * n);
    if (iters) HostCopier::add(J, iters, ho + o_it, sizeof(int32_t) * n);
    if (ctx->mu_out) HostCopier::add(J, ctx->mu_out, ho + o_mu, sizeof(double) * n);
    if (b_in + b_out >= STAGE_THREADS_MIN && ctx->host_pipe) host_copier(ctx)->run(J);
    else for (const CopyJob& j : J) memcpy(j.dst, j.src, j.bytes);
  }
  return rc;
}

// Join one submission: wait for its streams, collect the kernel time and the bad-index flags.
static int finish_query(dcol_ctx* ctx, int slot) {
  dcol_ctx::Slot& L = ctx->slots[slot];
  if (!L.pending) return L.result;
  L.pending = false;
  auto fail = [&](const char* what, cudaError_t e) { ctx->err = std::string(what) + ": " + cudaGetErrorString(e); L.result = -2; return -2; };
  cudaError_t e;
  if ((e = cudaEventSynchronize(L.ev_end)) != cudaSuccess) return fail("cudaEventSynchronize", e);
  double total_ms = 0.0;
  for (int c = 0; c < L.nchunks; c++) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, L.ev_t0[c], L.ev_t1[c]) == cudaSuccess) total_ms += ms;
  }
  L.ms = total_ms; ctx->last_ms = total_ms;
  L.result = 0;
  for (int c = 0; c < L.nchunks; c++)
    if (L.h_bad[c]) { ctx->err = "dcol: shape index out of range in shape1/shape2"; L.result = -1; }
  if (L.ticket >= 0) {
    const int h = ctx->done_head; ctx->done_head = (h + 1) % dcol_ctx::N_DONE;
    ctx->done_ticket[h] = L.ticket; ctx->done_result[h] = L.result; ctx->done_err[h] = L.result ? ctx->err : std::string();
  }
  return L.result;
}

// Join every submission still in flight (entry points that rewrite context-level buffers the kernels read)
static int drain_all(dcol_ctx* ctx) {
  int rc = 0;
  for (int k = 0; k < dcol_ctx::N_SLOTS; k++)
    if (ctx->slots[k].pending) { int r = finish_query(ctx, k); if (r && !rc) rc = r; }
  return rc;
}

extern "C" {

int dcol_proximity(dcol_ctx* ctx, int64_t n, const int32_t* shape1, const int32_t* shape2,
                   const double* r1, const double* att1, const double* r2, const double* att2,
                   int att_kind, double pdip_tol, int mem,
                   double* alpha, double* x, int32_t* status, int32_t* iters) {
  return run_query(ctx, Q_PROX, n, shape1, shape2, r1, att1, r2, att2, att_kind, pdip_tol, mem, alpha, x, nullptr, status, iters);
}

int dcol_proximity_gradient(dcol_ctx* ctx, int64_t n, const int32_t* shape1, const int32_t* shape2,
                            const double* r1, const double* att1, const double* r2, const double* att2,
                            int att_kind, double pdip_tol, int mem,
                            double* alpha, double* grad, int32_t* status, int32_t* iters) {
  return run_query(ctx, Q_GRAD, n, shape1, shape2, r1, att1, r2, att2, att_kind, pdip_tol, mem, alpha, nullptr, grad, status, iters);
}

int dcol_proximity_jacobian(dcol_ctx* ctx, int64_t n, const int32_t* shape1, const int32_t* shape2,
                            const double* r1, const double* att1, const double* r2, const double* att2,
                            int att_kind, double pdip_tol, int mem,
                            double* alpha, double* x, double* J, int32_t* status, int32_t* iters) {
  return run_query(ctx, Q_JAC, n, shape1, shape2, r1, att1, r2, att2, att_kind, pdip_tol, mem, alpha, x, J, status, iters);
}


static int check_query_kind(dcol_ctx* ctx, int query) {
  if (!ctx) return -1;
  if (query != Q_PROX && query != Q_GRAD && query != Q_JAC) { ctx->err = "dcol: query must be 0 (proximity), 1 (gradient) or 2 (jacobian)"; return -1; }
  return 0;
}

int dcol_query_indexed(dcol_ctx* ctx, int query, int64_t n, const int32_t* shape1, const int32_t* shape2,
                       const int32_t* pose1, const int32_t* pose2,
                       int64_t n_pose1, const double* r1, const double* att1,
                       int64_t n_pose2, const double* r2, const double* att2,
                       int att_kind, double pdip_tol, int mem,
                       double* alpha, double* x, double* deriv, int32_t* status, int32_t* iters) {
  if (check_query_kind(ctx, query)) return -1;
  if (!pose1 || !pose2) { ctx->err = "dcol_query_indexed: pose1/pose2 are NULL"; return -1; }
  return run_query(ctx, query, n, shape1, shape2, r1, att1, r2, att2, att_kind, pdip_tol, mem, alpha, x, deriv, status, iters,
                   pose1, pose2, n_pose1, n_pose2);
}

// Broad phase of dcol_scene_allpairs / dcol_scene_wrench (opt-in; the reference's contact example solves every pair):
// alpha_cut > 0 answers every pair whose bounding balls prove alpha* > alpha_cut without running the conic solver
// (alpha = the lower bound |r~_i - r~_j| / (rho_i + rho_j) > alpha_cut, status DCOL_STATUS_CULLED = 3, iters 0, zero derivative rows,
// NaN witness point); all other pairs are solved exactly as before, bit for bit.  alpha_cut <= 0 switches it off.
int dcol_scene_set_cull(dcol_ctx* ctx, double alpha_cut) {
  if (!ctx) return -1;
  ctx->cull_alpha = (alpha_cut > 0.0 && alpha_cut == alpha_cut) ? alpha_cut : 0.0;
  return 0;
}

// centres, cull kernel, survivor count: fills the compact pair list (s1, s2, p1, p2, oidx) and returns its length in *n_keep
static int scene_cull(dcol_ctx* ctx, cudaStream_t S, int n_bodies, int64_t k0, int64_t np, const int32_t* d_shape, const double* d_r, const double* d_att,
                      int att_kind, int nd, int32_t* s1, int32_t* s2, int32_t* p1, int32_t* p2, int32_t* oidx,
                      double* alpha, double* x, double* deriv, int32_t* status, int32_t* iters, int64_t* n_keep) {
  if (!ctx->h_count) CK(cudaHostAlloc((void**)&ctx->h_count, 2 * sizeof(int), cudaHostAllocDefault));
  CK(ctx->d_cull.need(sizeof(double) * 3 * (size_t)n_bodies + 256));
  double* d_centre = (double*)ctx->d_cull.p;
  int* d_count = (int*)((char*)ctx->d_cull.p + sizeof(double) * 3 * (size_t)n_bodies);      // count, bad flag
  CK(cudaMemsetAsync(d_count, 0, 2 * sizeof(int), S));
  k_body_centres<<<(n_bodies + 127) / 128, 128, 0, S>>>(n_bodies, d_shape, (const DevShape*)ctx->d_shapes.p, ctx->n_shapes, d_r, d_att, att_kind,
                                                        d_centre, d_count + 1);
  k_allpairs_cull<<<(int)std::min<int64_t>((np + 255) / 256, 148 * 16), 256, 0, S>>>(n_bodies, k0, np, d_shape, ctx->n_shapes, (const double*)ctx->d_brad.p,
                                                                                       d_centre, ctx->cull_alpha, nd, s1, s2, p1, p2, oidx, d_count,
                                                                                       alpha, x, deriv, status, iters);
  CK(cudaGetLastError());
  ctx->launches += 2;
  CK(cudaMemcpyAsync(ctx->h_count, d_count, 2 * sizeof(int), cudaMemcpyDeviceToHost, S));
  CK(cudaStreamSynchronize(S));
  *n_keep = ctx->h_count[0];
  return 0;
}

int64_t dcol_scene_pair_count(int n_bodies) { return n_bodies < 2 ? 0 : (int64_t)n_bodies * (n_bodies - 1) / 2; }

int dcol_scene_allpairs(dcol_ctx* ctx, int query, int n_bodies, const int32_t* shape, const double* r, const double* att,
                        int att_kind, double pdip_tol, int mem,
                        double* alpha, double* x, double* deriv, int32_t* status, int32_t* iters) {
  return dcol_scene_allpairs_range(ctx, query, n_bodies, 0, dcol_scene_pair_count(n_bodies), shape, r, att, att_kind, pdip_tol, mem,
                                   alpha, x, deriv, status, iters);
}

int dcol_scene_allpairs_range(dcol_ctx* ctx, int query, int n_bodies, int64_t k_begin, int64_t k_end,
                              const int32_t* shape, const double* r, const double* att,
                              int att_kind, double pdip_tol, int mem,
                              double* alpha, double* x, double* deriv, int32_t* status, int32_t* iters) {
  if (check_query_kind(ctx, query)) return -1;
  ctx->err.clear();
  if (n_bodies < 0 || !shape || !r || !att) { ctx->err = "dcol_scene_allpairs: bad arguments"; return -1; }
  if (k_begin < 0 || k_end < k_begin || k_end > dcol_scene_pair_count(n_bodies)) { ctx->err = "dcol_scene_allpairs_range: need 0 <= k_begin <= k_end <= N(N-1)/2"; return -1; }
  if (att_kind != DCOL_ATT_QUAT && att_kind != DCOL_ATT_MRP) { ctx->err = "dcol: att_kind must be 0 (quaternion) or 1 (MRP)"; return -1; }
  if (mem != DCOL_MEM_HOST && mem != DCOL_MEM_DEVICE) { ctx->err = "dcol: mem must be 0 (host) or 1 (device)"; return -1; }
  const int64_t np = k_end - k_begin;
  if (np == 0) return 0;
  if (np > 0x7fffffff) { ctx->err = "dcol_scene_allpairs: too many pairs for one call"; return -1; }
  CK(cudaSetDevice(ctx->device));
  { int rc = drain_all(ctx); if (rc) return rc; }      // context-level scratch (d_scene, d_pairs) is about to be rewritten
  cudaStream_t S = ctx->stream();
  const int na = (att_kind == DCOL_ATT_QUAT) ? 4 : 3;
  // A handful of pairs on host arrays (the contact example's own scene: 13 bodies, 78 pairs): the host writes the pair
  // list itself and the call takes the one-launch path of small host-pointer queries (same records, same results).
  if (mem == DCOL_MEM_HOST && np <= DCOL_PACKED_MAX && ctx->cull_alpha <= 0.0) {
    std::vector<int32_t> s1(np), s2(np);
    std::vector<double> r1(3 * np), a1((size_t)na * np), r2(3 * np), a2((size_t)na * np);
    int64_t i = 0, k = 0;
    while ((i + 1) * (2 * (int64_t)n_bodies - i - 2) / 2 <= k_begin) i++;             // row of the first pair
    int64_t j = k_begin - i * (2 * (int64_t)n_bodies - i - 1) / 2 + i + 1;
    for (; k < np; k++) {
      s1[k] = shape[i]; s2[k] = shape[j];
      for (int a = 0; a < 3; a++) { r1[3 * k + a] = r[3 * i + a]; r2[3 * k + a] = r[3 * j + a]; }
      for (int a = 0; a < na; a++) { a1[na * k + a] = att[na * i + a]; a2[na * k + a] = att[na * j + a]; }
      if (++j == n_bodies) { i++; j = i + 1; }
    }
    return run_query(ctx, query, np, s1.data(), s2.data(), r1.data(), a1.data(), r2.data(), a2.data(), att_kind, pdip_tol, DCOL_MEM_HOST,
                     alpha, x, deriv, status, iters);
  }
  // body arrays on the device
  const int32_t* d_shape = shape; const double* d_r = r; const double* d_att = att;
  if (mem == DCOL_MEM_HOST) {
    size_t b0 = align256(sizeof(int32_t) * n_bodies), b1 = align256(sizeof(double) * 3 * n_bodies), b2 = align256(sizeof(double) * na * n_bodies);
    CK(ctx->d_scene.need(b0 + b1 + b2));
    char* d = (char*)ctx->d_scene.p;
    CK(cudaMemcpyAsync(d, shape, sizeof(int32_t) * n_bodies, cudaMemcpyHostToDevice, S));
    CK(cudaMemcpyAsync(d + b0, r, sizeof(double) * 3 * n_bodies, cudaMemcpyHostToDevice, S));
    CK(cudaMemcpyAsync(d + b0 + b1, att, sizeof(double) * na * n_bodies, cudaMemcpyHostToDevice, S));
    d_shape = (const int32_t*)d; d_r = (const double*)(d + b0); d_att = (const double*)(d + b0 + b1);
  }
  // pair lists generated on the device
  CK(ctx->d_pairs.need(sizeof(int32_t) * 5 * (size_t)np));
  int32_t* s1 = (int32_t*)ctx->d_pairs.p; int32_t* s2 = s1 + np; int32_t* p1 = s2 + np; int32_t* p2 = p1 + np; int32_t* oidx = p2 + np;
  if (ctx->cull_alpha > 0.0) {
    // broad phase: results of all np pairs live in device arrays (the caller's, or scratch when the caller's are host memory);
    // far pairs are answered by the cull kernel, the survivors by the solver through the result indirection oidx
    const int nt = 6 + 2 * na;
    const int nd = (query == Q_GRAD) ? nt : (query == Q_JAC ? 4 * nt : 0);
    const bool host = (mem == DCOL_MEM_HOST);
    double* q_alpha = alpha; double* q_x = x; double* q_der = deriv; int32_t* q_st = status; int32_t* q_it = iters;
    const size_t w_a = align256(sizeof(double) * (size_t)np), w_x = align256(sizeof(double) * 3 * (size_t)np),
                 w_d = align256(sizeof(double) * (size_t)nd * (size_t)np), w_i = align256(sizeof(int32_t) * (size_t)np);
    if (host) {
      CK(ctx->d_wrench.need(w_a + w_x + w_d + 2 * w_i));
      char* wb = (char*)ctx->d_wrench.p;
      q_alpha = (double*)wb; q_x = x ? (double*)(wb + w_a) : nullptr; q_der = nd ? (double*)(wb + w_a + w_x) : nullptr;
      q_st = (int32_t*)(wb + w_a + w_x + w_d); q_it = (int32_t*)(wb + w_a + w_x + w_d + w_i);
    }
    if (!q_alpha || (query != Q_GRAD && !q_x) || (nd && !q_der)) { ctx->err = "dcol_scene_allpairs: NULL output pointer"; return -1; }
    int64_t n_keep = 0;
    { int rc = scene_cull(ctx, S, n_bodies, k_begin, np, d_shape, d_r, d_att, att_kind, nd, s1, s2, p1, p2, oidx, q_alpha, q_x, q_der, q_st, q_it, &n_keep); if (rc) return rc; }
    if (ctx->h_count[1]) { ctx->err = "dcol: shape index out of range in shape"; return -1; }
    if (n_keep > 0) {
      int rc = run_query(ctx, query, n_keep, s1, s2, d_r, d_att, d_r, d_att, att_kind, pdip_tol, DCOL_MEM_DEVICE, q_alpha, q_x, q_der, q_st, q_it,
                         p1, p2, n_bodies, n_bodies, DCOL_MEM_DEVICE, 0, false, oidx);
      if (rc) return rc;
    }
    if (host) {
      std::vector<CopyJob> items;
      items.push_back({(char*)alpha, (const char*)q_alpha, sizeof(double) * (size_t)np});
      if (x) items.push_back({(char*)x, (const char*)q_x, sizeof(double) * 3 * (size_t)np});
      if (nd) items.push_back({(char*)deriv, (const char*)q_der, sizeof(double) * (size_t)nd * (size_t)np});
      if (status) items.push_back({(char*)status, (const char*)q_st, sizeof(int32_t) * (size_t)np});
      if (iters) items.push_back({(char*)iters, (const char*)q_it, sizeof(int32_t) * (size_t)np});
      return results_to_host(ctx, S, items);
    }
    return 0;
  }
  k_allpairs<<<(int)std::min<int64_t>((np + 255) / 256, 148 * 16), 256, 0, S>>>(n_bodies, k_begin, np, d_shape, s1, s2, p1, p2);
  CK(cudaGetLastError());
  ctx->launches++;
  return run_query(ctx, query, np, s1, s2, d_r, d_att, d_r, d_att, att_kind, pdip_tol, mem, alpha, x, deriv, status, iters,
                   p1, p2, n_bodies, n_bodies, DCOL_MEM_DEVICE);
}


int dcol_scene_wrench(dcol_ctx* ctx, int n_bodies, const int32_t* shape, const double* r, const double* att, const double* lambda,
                      int att_kind, double pdip_tol, int mem,
                      double* alpha, double* wrench, int32_t* status, int32_t* iters) {
  return dcol_scene_wrench_range(ctx, n_bodies, 0, dcol_scene_pair_count(n_bodies), shape, r, att, lambda, att_kind, pdip_tol, mem,
                                 alpha, wrench, status, iters);
}

int dcol_scene_wrench_range(dcol_ctx* ctx, int n_bodies, int64_t k_begin, int64_t k_end,
                            const int32_t* shape, const double* r, const double* att, const double* lambda,
                            int att_kind, double pdip_tol, int mem,
                            double* alpha, double* wrench, int32_t* status, int32_t* iters) {
  if (!ctx) return -1;
  ctx->err.clear();
  if (n_bodies < 0 || !shape || !r || !att || !lambda || !wrench) { ctx->err = "dcol_scene_wrench: bad arguments"; return -1; }
  if (att_kind != DCOL_ATT_QUAT && att_kind != DCOL_ATT_MRP) { ctx->err = "dcol: att_kind must be 0 (quaternion) or 1 (MRP)"; return -1; }
  if (mem != DCOL_MEM_HOST && mem != DCOL_MEM_DEVICE) { ctx->err = "dcol: mem must be 0 (host) or 1 (device)"; return -1; }
  if (k_begin < 0 || k_end < k_begin || k_end > dcol_scene_pair_count(n_bodies)) { ctx->err = "dcol_scene_wrench_range: need 0 <= k_begin <= k_end <= N(N-1)/2"; return -1; }
  const int64_t np = k_end - k_begin;
  if (np > 0x7fffffff) { ctx->err = "dcol_scene_wrench: too many pairs for one call"; return -1; }
  CK(cudaSetDevice(ctx->device));
  { int rc = drain_all(ctx); if (rc) return rc; }
  cudaStream_t S = ctx->stream();
  const int na = (att_kind == DCOL_ATT_QUAT) ? 4 : 3, nt = 6 + 2 * na;
  if (np == 0) {
    if (n_bodies > 0) {
      if (mem == DCOL_MEM_HOST) memset(wrench, 0, sizeof(double) * 6 * n_bodies);
      else { CK(cudaMemsetAsync(wrench, 0, sizeof(double) * 6 * n_bodies, S)); CK(cudaStreamSynchronize(S)); }
    }
    return 0;
  }
  // per-pair results stay on the device: gradient rows always; alpha / status / iters too when the caller's are host memory
  const size_t w0 = align256(sizeof(double) * nt * (size_t)np), w1 = align256(sizeof(double) * (size_t)np),
               w2 = align256(sizeof(int32_t) * (size_t)np), w3 = align256(sizeof(double) * 6 * (size_t)n_bodies);
  CK(ctx->d_wrench.need(w0 + 2 * w1 + 2 * w2 + w3));
  char* wb = (char*)ctx->d_wrench.p;
  double* d_grad = (double*)wb;
  double* d_lambda = (double*)(wb + w0);
  double* d_alpha = (double*)(wb + w0 + w1);
  int32_t* d_status = (int32_t*)(wb + w0 + 2 * w1);
  int32_t* d_iters = (int32_t*)(wb + w0 + 2 * w1 + w2);
  double* d_out = (double*)(wb + w0 + 2 * w1 + 2 * w2);
  const int32_t* d_shape = shape; const double* d_r = r; const double* d_att = att; const double* d_lam = lambda;
  if (mem == DCOL_MEM_HOST) {
    size_t b0 = align256(sizeof(int32_t) * n_bodies), b1 = align256(sizeof(double) * 3 * n_bodies), b2 = align256(sizeof(double) * na * n_bodies);
    CK(ctx->d_scene.need(b0 + b1 + b2));
    char* d = (char*)ctx->d_scene.p;
    CK(cudaMemcpyAsync(d, shape, sizeof(int32_t) * n_bodies, cudaMemcpyHostToDevice, S));
    CK(cudaMemcpyAsync(d + b0, r, sizeof(double) * 3 * n_bodies, cudaMemcpyHostToDevice, S));
    CK(cudaMemcpyAsync(d + b0 + b1, att, sizeof(double) * na * n_bodies, cudaMemcpyHostToDevice, S));
    CK(cudaMemcpyAsync(d_lambda, lambda, sizeof(double) * (size_t)np, cudaMemcpyHostToDevice, S));
    d_shape = (const int32_t*)d; d_r = (const double*)(d + b0); d_att = (const double*)(d + b0 + b1); d_lam = d_lambda;
  }
  CK(ctx->d_pairs.need(sizeof(int32_t) * 5 * (size_t)np));
  int32_t* s1 = (int32_t*)ctx->d_pairs.p; int32_t* s2 = s1 + np; int32_t* p1 = s2 + np; int32_t* p2 = p1 + np; int32_t* oidx = p2 + np;
  const bool host = (mem == DCOL_MEM_HOST);
  const bool cull = ctx->cull_alpha > 0.0;
  double* q_alpha = (host || !alpha) ? d_alpha : alpha;
  int32_t* q_status = host ? ((status || cull) ? d_status : nullptr) : (status ? status : (cull ? d_status : nullptr));
  int32_t* q_iters = host ? (iters ? d_iters : nullptr) : iters;
  int rc = 0;
  if (cull) {
    int64_t n_keep = 0;
    rc = scene_cull(ctx, S, n_bodies, k_begin, np, d_shape, d_r, d_att, att_kind, nt, s1, s2, p1, p2, oidx, q_alpha, nullptr, nullptr, q_status, q_iters, &n_keep);
    if (rc) return rc;
    if (ctx->h_count[1]) { ctx->err = "dcol: shape index out of range in shape"; return -1; }
    if (n_keep > 0) rc = run_query(ctx, Q_GRAD, n_keep, s1, s2, d_r, d_att, d_r, d_att, att_kind, pdip_tol, DCOL_MEM_DEVICE,
                                   q_alpha, nullptr, d_grad, q_status, q_iters, p1, p2, n_bodies, n_bodies, DCOL_MEM_DEVICE, 0, false, oidx);
  } else {
    k_allpairs<<<(int)std::min<int64_t>((np + 255) / 256, 148 * 16), 256, 0, S>>>(n_bodies, k_begin, np, d_shape, s1, s2, p1, p2);
    CK(cudaGetLastError());
    ctx->launches++;
    rc = run_query(ctx, Q_GRAD, np, s1, s2, d_r, d_att, d_r, d_att, att_kind, pdip_tol, DCOL_MEM_DEVICE,
                   q_alpha, nullptr, d_grad, q_status, q_iters, p1, p2, n_bodies, n_bodies, DCOL_MEM_DEVICE);
  }
  if (rc) return rc;
  double* o = host ? d_out : wrench;
  k_scene_wrench<<<n_bodies, 128, 0, S>>>(n_bodies, k_begin, k_end, d_grad, d_lam, cull ? q_status : nullptr, d_att, att_kind, o);
  CK(cudaGetLastError());
  ctx->launches++;
  if (host) {
    std::vector<CopyJob> items;
    items.push_back({(char*)wrench, (const char*)d_out, sizeof(double) * 6 * (size_t)n_bodies});
    if (alpha) items.push_back({(char*)alpha, (const char*)d_alpha, sizeof(double) * (size_t)np});
    if (status) items.push_back({(char*)status, (const char*)d_status, sizeof(int32_t) * (size_t)np});
    if (iters) items.push_back({(char*)iters, (const char*)d_iters, sizeof(int32_t) * (size_t)np});
    return results_to_host(ctx, S, items);
  }
  CK(cudaStreamSynchronize(S));
  return 0;
}

int dcol_se2_constraints(dcol_ctx* ctx, int64_t n_states, const double* xytheta, int32_t vehicle_shape,
                         int n_obs, const int32_t* obs_shape, const double* r_obs, const double* p_obs,
                         double pdip_tol, int mem, double* c, double* Jc, int32_t* status) {
  if (!ctx) return -1;
  ctx->err.clear();
  if (n_states < 0 || n_obs < 0 || !xytheta || !obs_shape || !r_obs || !p_obs || !c) { ctx->err = "dcol_se2_constraints: bad arguments"; return -1; }
  if (mem != DCOL_MEM_HOST && mem != DCOL_MEM_DEVICE) { ctx->err = "dcol: mem must be 0 (host) or 1 (device)"; return -1; }
  const int64_t np = n_states * (int64_t)n_obs;
  if (np == 0) return 0;
  if (np > 0x7fffffff) { ctx->err = "dcol_se2_constraints: too many pairs for one call"; return -1; }
  CK(cudaSetDevice(ctx->device));
  { int rc = drain_all(ctx); if (rc) return rc; }
  cudaStream_t S = ctx->stream();
  const bool host = (mem == DCOL_MEM_HOST);
  // device scratch: gradient rows | alpha | status | vehicle poses | (host mode) states, c, Jc
  const size_t w0 = align256(sizeof(double) * 12 * (size_t)np), w1 = align256(sizeof(double) * (size_t)np),
               w2 = align256(sizeof(int32_t) * (size_t)np), w3 = align256(sizeof(double) * 3 * (size_t)n_states),
               w4 = align256(sizeof(double) * 3 * (size_t)np);
  CK(ctx->d_wrench.need(w0 + 2 * w1 + w2 + 3 * w3 + w4));
  char* wb = (char*)ctx->d_wrench.p;
  double* d_grad = (double*)wb;
  double* d_alpha = (double*)(wb + w0);
  double* d_c = (double*)(wb + w0 + w1);
  int32_t* d_status = (int32_t*)(wb + w0 + 2 * w1);
  double* d_r1 = (double*)(wb + w0 + 2 * w1 + w2);
  double* d_p1 = (double*)(wb + w0 + 2 * w1 + w2 + w3);
  double* d_xyt = (double*)(wb + w0 + 2 * w1 + w2 + 2 * w3);
  double* d_J = (double*)(wb + w0 + 2 * w1 + w2 + 3 * w3);
  const double* x_dev = xytheta; const int32_t* os_dev = obs_shape; const double* ro_dev = r_obs; const double* po_dev = p_obs;
  if (host) {
    size_t b0 = align256(sizeof(int32_t) * n_obs), b1 = align256(sizeof(double) * 3 * n_obs);
    CK(ctx->d_scene.need(b0 + 2 * b1));
    char* d = (char*)ctx->d_scene.p;
    CK(cudaMemcpyAsync(d, obs_shape, sizeof(int32_t) * n_obs, cudaMemcpyHostToDevice, S));
    CK(cudaMemcpyAsync(d + b0, r_obs, sizeof(double) * 3 * n_obs, cudaMemcpyHostToDevice, S));
    CK(cudaMemcpyAsync(d + b0 + b1, p_obs, sizeof(double) * 3 * n_obs, cudaMemcpyHostToDevice, S));
    CK(cudaMemcpyAsync(d_xyt, xytheta, sizeof(double) * 3 * (size_t)n_states, cudaMemcpyHostToDevice, S));
    os_dev = (const int32_t*)d; ro_dev = (const double*)(d + b0); po_dev = (const double*)(d + b0 + b1); x_dev = d_xyt;
  }
  CK(ctx->d_pairs.need(sizeof(int32_t) * 4 * (size_t)np));
  int32_t* s1 = (int32_t*)ctx->d_pairs.p; int32_t* s2 = s1 + np; int32_t* i1 = s2 + np; int32_t* i2 = i1 + np;
  const int blocks = (int)std::min<int64_t>((np + 255) / 256, 148 * 16);
  k_se2_pairs<<<blocks, 256, 0, S>>>(n_states, n_obs, x_dev, vehicle_shape, os_dev, d_r1, d_p1, s1, s2, i1, i2);
  CK(cudaGetLastError());
  ctx->launches++;
  int32_t* q_status = host ? (status ? d_status : nullptr) : status;
  int rc = run_query(ctx, Q_GRAD, np, s1, s2, d_r1, d_p1, ro_dev, po_dev, DCOL_ATT_MRP, pdip_tol, DCOL_MEM_DEVICE,
                     d_alpha, nullptr, d_grad, q_status, nullptr, i1, i2, n_states, n_obs, DCOL_MEM_DEVICE);
  if (rc) return rc;
  double* o_c = host ? d_c : c;
  double* o_J = (host || !Jc) ? d_J : Jc;
  k_se2_rows<<<blocks, 256, 0, S>>>(np, n_obs, x_dev, d_alpha, d_grad, o_c, o_J);
  CK(cudaGetLastError());
  ctx->launches++;
  if (host) {
    std::vector<CopyJob> items;
    items.push_back({(char*)c, (const char*)d_c, sizeof(double) * (size_t)np});
    if (Jc) items.push_back({(char*)Jc, (const char*)d_J, sizeof(double) * 3 * (size_t)np});
    if (status) items.push_back({(char*)status, (const char*)d_status, sizeof(int32_t) * (size_t)np});
    return results_to_host(ctx, S, items);
  }
  CK(cudaStreamSynchronize(S));
  return 0;
}

int dcol_se3_constraints(dcol_ctx* ctx, int64_t n_states, const double* states, int state_stride, int r_col, int p_col,
                         int32_t vehicle_shape, int n_obs, const int32_t* obs_shape, const double* r_obs, const double* p_obs,
                         double pdip_tol, int mem, double* c, double* Jc, int32_t* status) {
  if (!ctx) return -1;
  ctx->err.clear();
  if (n_states < 0 || n_obs < 0 || !states || !obs_shape || !r_obs || !p_obs || !c) { ctx->err = "dcol_se3_constraints: bad arguments"; return -1; }
  if (state_stride < 6 || r_col < 0 || p_col < 0 || r_col + 3 > state_stride || p_col + 3 > state_stride ||
      (r_col < p_col + 3 && p_col < r_col + 3)) { ctx->err = "dcol_se3_constraints: position and MRP columns must be disjoint triples inside a state row"; return -1; }
  if (mem != DCOL_MEM_HOST && mem != DCOL_MEM_DEVICE) { ctx->err = "dcol: mem must be 0 (host) or 1 (device)"; return -1; }
  const int64_t np = n_states * (int64_t)n_obs;
  if (np == 0) return 0;
  if (np > 0x7fffffff) { ctx->err = "dcol_se3_constraints: too many pairs for one call"; return -1; }
  CK(cudaSetDevice(ctx->device));
  { int rc = drain_all(ctx); if (rc) return rc; }
  cudaStream_t S = ctx->stream();
  const bool host = (mem == DCOL_MEM_HOST);
  // device scratch: gradient rows | alpha | c | status | vehicle poses (r, p) | (host mode) states, Jc
  const size_t w0 = align256(sizeof(double) * 12 * (size_t)np), w1 = align256(sizeof(double) * (size_t)np),
               w2 = align256(sizeof(int32_t) * (size_t)np), w3 = align256(sizeof(double) * 3 * (size_t)n_states),
               w4 = align256(sizeof(double) * (size_t)state_stride * (size_t)n_states), w5 = align256(sizeof(double) * 6 * (size_t)np);
  CK(ctx->d_wrench.need(w0 + 2 * w1 + w2 + 2 * w3 + (host ? w4 + w5 : 0)));
  char* wb = (char*)ctx->d_wrench.p;
  double* d_grad = (double*)wb;
  double* d_alpha = (double*)(wb + w0);
  double* d_c = (double*)(wb + w0 + w1);
  int32_t* d_status = (int32_t*)(wb + w0 + 2 * w1);
  double* d_r1 = (double*)(wb + w0 + 2 * w1 + w2);
  double* d_p1 = (double*)(wb + w0 + 2 * w1 + w2 + w3);
  double* d_X = (double*)(wb + w0 + 2 * w1 + w2 + 2 * w3);
  double* d_J = (double*)(wb + w0 + 2 * w1 + w2 + 2 * w3 + w4);
  const double* x_dev = states; const int32_t* os_dev = obs_shape; const double* ro_dev = r_obs; const double* po_dev = p_obs;
  if (host) {
    size_t b0 = align256(sizeof(int32_t) * n_obs), b1 = align256(sizeof(double) * 3 * n_obs);
    CK(ctx->d_scene.need(b0 + 2 * b1));
    char* d = (char*)ctx->d_scene.p;
    CK(cudaMemcpyAsync(d, obs_shape, sizeof(int32_t) * n_obs, cudaMemcpyHostToDevice, S));
    CK(cudaMemcpyAsync(d + b0, r_obs, sizeof(double) * 3 * n_obs, cudaMemcpyHostToDevice, S));
    CK(cudaMemcpyAsync(d + b0 + b1, p_obs, sizeof(double) * 3 * n_obs, cudaMemcpyHostToDevice, S));
    CK(cudaMemcpyAsync(d_X, states, sizeof(double) * (size_t)state_stride * (size_t)n_states, cudaMemcpyHostToDevice, S));
    os_dev = (const int32_t*)d; ro_dev = (const double*)(d + b0); po_dev = (const double*)(d + b0 + b1); x_dev = d_X;
  }
  CK(ctx->d_pairs.need(sizeof(int32_t) * 4 * (size_t)np));
  int32_t* s1 = (int32_t*)ctx->d_pairs.p; int32_t* s2 = s1 + np; int32_t* i1 = s2 + np; int32_t* i2 = i1 + np;
  const int blocks = (int)std::min<int64_t>((np + 255) / 256, 148 * 16);
  k_se3_pairs<<<blocks, 256, 0, S>>>(n_states, n_obs, x_dev, state_stride, r_col, p_col, vehicle_shape, os_dev, d_r1, d_p1, s1, s2, i1, i2);
  CK(cudaGetLastError());
  ctx->launches++;
  int32_t* q_status = host ? (status ? d_status : nullptr) : status;
  int rc = run_query(ctx, Q_GRAD, np, s1, s2, d_r1, d_p1, ro_dev, po_dev, DCOL_ATT_MRP, pdip_tol, DCOL_MEM_DEVICE,
                     d_alpha, nullptr, d_grad, q_status, nullptr, i1, i2, n_states, n_obs, DCOL_MEM_DEVICE);
  if (rc) return rc;
  double* o_c = host ? d_c : c;
  double* o_J = host ? (Jc ? d_J : nullptr) : Jc;
  k_se3_rows<<<blocks, 256, 0, S>>>(np, d_alpha, d_grad, o_c, o_J);
  CK(cudaGetLastError());
  ctx->launches++;
  if (host) {
    std::vector<CopyJob> items;
    items.push_back({(char*)c, (const char*)d_c, sizeof(double) * (size_t)np});
    if (Jc) items.push_back({(char*)Jc, (const char*)d_J, sizeof(double) * 6 * (size_t)np});
    if (status) items.push_back({(char*)status, (const char*)d_status, sizeof(int32_t) * (size_t)np});
    return results_to_host(ctx, S, items);
  }
  CK(cudaStreamSynchronize(S));
  return 0;
}

// Generic batched conic solve (dcol_socp.cu): the reference's solve_socp on n problems of common dimensions.
static int solve_socp_impl(dcol_ctx* ctx, int64_t n, int nx, int n_ort, int d1, int d2, const double* c, const double* G, const double* h,
                           double pdip_tol, int mem, double* x, double* s, double* z, int32_t* status, int32_t* iters, double* trace) {
  if (!ctx) return -1;
  ctx->err.clear();
  const int ns = n_ort + d1 + d2;
  if (n < 0 || nx < 1 || nx > socp_limit(0) || n_ort < 0 || d1 < 0 || d2 < 0 || ns < 1 || ns > socp_limit(1) || d1 > socp_limit(2) || d2 > socp_limit(2) ||
      d1 == 1 || d2 == 1) { ctx->err = "dcol_solve_socp: dimensions out of range (nx <= 16, n_ort + d1 + d2 <= 96, d in {0, 2..8})"; return -1; }
  if (mem != DCOL_MEM_HOST && mem != DCOL_MEM_DEVICE) { ctx->err = "dcol: mem must be 0 (host) or 1 (device)"; return -1; }
  if (n == 0) return 0;
  if (!c || !G || !h || !x || !s || !z) { ctx->err = "dcol_solve_socp: NULL pointer"; return -1; }
  CK(cudaSetDevice(ctx->device));
  { int rc = drain_all(ctx); if (rc) return rc; }
  cudaStream_t S = ctx->stream();
  const double *dc_ = c, *dG = G, *dh = h; double *dx = x, *ds = s, *dz = z; int32_t *dst = status, *dit = iters;
  const size_t bc = align256(sizeof(double) * nx * (size_t)n), bG = align256(sizeof(double) * (size_t)ns * nx * (size_t)n), bh = align256(sizeof(double) * ns * (size_t)n),
               bi = align256(sizeof(int32_t) * (size_t)n), bt = trace ? align256(sizeof(double) * DCOL_TRACE_ROWS * DCOL_TRACE_COLS * (size_t)n) : 0;
  double* dtr = trace;
  if (mem == DCOL_MEM_HOST) {
    CK(ctx->d_wrench.need(2 * bc + bG + 3 * bh + 2 * bi + bt));
    char* w = (char*)ctx->d_wrench.p;
    CK(cudaMemcpyAsync(w, c, sizeof(double) * nx * (size_t)n, cudaMemcpyHostToDevice, S));
    CK(cudaMemcpyAsync(w + bc, G, sizeof(double) * (size_t)ns * nx * (size_t)n, cudaMemcpyHostToDevice, S));
    CK(cudaMemcpyAsync(w + bc + bG, h, sizeof(double) * ns * (size_t)n, cudaMemcpyHostToDevice, S));
    dc_ = (const double*)w; dG = (const double*)(w + bc); dh = (const double*)(w + bc + bG);
    dx = (double*)(w + bc + bG + bh); ds = (double*)(w + 2 * bc + bG + bh); dz = (double*)(w + 2 * bc + bG + 2 * bh);
    dst = (int32_t*)(w + 2 * bc + bG + 3 * bh); dit = (int32_t*)(w + 2 * bc + bG + 3 * bh + bi);
    if (trace) {                                  // rows past a problem's last iteration keep the caller's values
      dtr = (double*)(w + 2 * bc + bG + 3 * bh + 2 * bi);
      CK(cudaMemcpyAsync(dtr, trace, sizeof(double) * DCOL_TRACE_ROWS * DCOL_TRACE_COLS * (size_t)n, cudaMemcpyHostToDevice, S));
    }
  }
  CK(launch_solve_socp(n, nx, n_ort, d1, d2, dc_, dG, dh, pdip_tol, ctx->max_iters, dx, ds, dz, dst, dit, dtr, S));
  ctx->launches++;
  if (mem == DCOL_MEM_HOST) {
    CK(cudaMemcpyAsync(x, dx, sizeof(double) * nx * (size_t)n, cudaMemcpyDeviceToHost, S));
    CK(cudaMemcpyAsync(s, ds, sizeof(double) * ns * (size_t)n, cudaMemcpyDeviceToHost, S));
    CK(cudaMemcpyAsync(z, dz, sizeof(double) * ns * (size_t)n, cudaMemcpyDeviceToHost, S));
    if (status) CK(cudaMemcpyAsync(status, dst, sizeof(int32_t) * (size_t)n, cudaMemcpyDeviceToHost, S));
    if (iters) CK(cudaMemcpyAsync(iters, dit, sizeof(int32_t) * (size_t)n, cudaMemcpyDeviceToHost, S));
    if (trace) CK(cudaMemcpyAsync(trace, dtr, sizeof(double) * DCOL_TRACE_ROWS * DCOL_TRACE_COLS * (size_t)n, cudaMemcpyDeviceToHost, S));
  }
  CK(cudaStreamSynchronize(S));
  return 0;
}
int dcol_solve_socp(dcol_ctx* ctx, int64_t n, int nx, int n_ort, int d1, int d2, const double* c, const double* G, const double* h,
                    double pdip_tol, int mem, double* x, double* s, double* z, int32_t* status, int32_t* iters) {
  return solve_socp_impl(ctx, n, nx, n_ort, d1, d2, c, G, h, pdip_tol, mem, x, s, z, status, iters, nullptr);
}
int dcol_solve_socp_trace(dcol_ctx* ctx, int64_t n, int nx, int n_ort, int d1, int d2, const double* c, const double* G, const double* h,
                          double pdip_tol, int mem, double* x, double* s, double* z, int32_t* status, int32_t* iters, double* trace) {
  if (ctx && !trace) { ctx->err = "dcol_solve_socp_trace: NULL trace"; return -1; }
  return solve_socp_impl(ctx, n, nx, n_ort, d1, d2, c, G, h, pdip_tol, mem, x, s, z, status, iters, trace);
}

// ---- asynchronous submissions (host buffers): up to three in flight; inputs of submission k+1 cross PCIe while
// submission k computes and submission k-1's results go back
int dcol_submit(dcol_ctx* ctx, int query, int64_t n, const int32_t* shape1, const int32_t* shape2,
                const double* r1, const double* att1, const double* r2, const double* att2,
                int att_kind, double pdip_tol,
                double* alpha, double* x, double* deriv, int32_t* status, int32_t* iters) {
  if (check_query_kind(ctx, query)) return -1;
  const int ticket = ctx->next_ticket;
  const int slot = ticket % dcol_ctx::N_SLOTS;
  int rc = run_query(ctx, query, n, shape1, shape2, r1, att1, r2, att2, att_kind, pdip_tol, DCOL_MEM_HOST,
                     alpha, x, deriv, status, iters, nullptr, nullptr, 0, 0, -1, slot, true);
  if (rc) return rc;
  ctx->slots[slot].ticket = ticket;
  ctx->next_ticket = (ticket + 1) & 0x3fffffff;
  return ticket;
}

int dcol_wait(dcol_ctx* ctx, int ticket) {
  if (!ctx) return -1;
  for (int k = 0; k < dcol_ctx::N_SLOTS; k++)
    if (ctx->slots[k].ticket == ticket) return finish_query(ctx, k);
  for (int h = 0; h < dcol_ctx::N_DONE; h++)
    if (ctx->done_ticket[h] == ticket) { if (ctx->done_result[h]) ctx->err = ctx->done_err[h]; return ctx->done_result[h]; }
  ctx->err = "dcol_wait: unknown ticket (the outcomes of the last 16 submissions are kept)";
  return -1;
}

// Self-test of the host copy engine (no GPU needed; run by the CPU test suite): `rounds` batches of slices with ragged
// sizes and offsets between two buffers of `bytes` bytes on `threads` threads; returns 0 if every byte arrived.
int dcol_host_copy_selftest(int threads, int64_t bytes, int rounds) {
  if (threads < 1 || threads > 64 || bytes < 1 || rounds < 1) return -1;
  std::vector<unsigned char> src((size_t)bytes), dst((size_t)bytes);
  uint64_t x = 0x9e3779b97f4a7c15ull;
  auto rnd = [&]() { x ^= x << 13; x ^= x >> 7; x ^= x << 17; return x; };
  HostCopier cp;
  cp.start(threads - 1);
  for (int r = 0; r < rounds; r++) {
    for (size_t i = 0; i < src.size(); i++) src[i] = (unsigned char)rnd();
    std::fill(dst.begin(), dst.end(), (unsigned char)0);
    std::vector<CopyJob> jobs;
    for (size_t o = 0; o < src.size();) {      // ragged pieces: 1 byte .. 3 MB, cut into slices by add()
      const size_t len = std::min<size_t>(src.size() - o, 1 + (size_t)(rnd() % ((r & 1) ? 4096 : (3u << 20))));
      HostCopier::add(jobs, dst.data() + o, src.data() + o, len);
      o += len;
    }
    cp.run(jobs);
    if (!jobs.empty() || memcmp(src.data(), dst.data(), src.size()) != 0) return 1 + r;
  }
  return 0;
}

// ---- device groups: one context per listed device, contiguous index slices, one host thread per member and call
struct dcol_group {
  std::vector<dcol_ctx*> members;
  std::string err;
  double last_ms = -1.0;
};
thread_local std::string g_group_create_err;

int dcol_group_create(dcol_group** out, const int* device_ids, int n_devices) {
  if (!out) { g_group_create_err = "dcol_group_create: out is NULL"; return -1; }
  *out = nullptr;
  if (!device_ids || n_devices < 1 || n_devices > 64) { g_group_create_err = "dcol_group_create: need 1..64 device ids"; return -1; }
  dcol_group* g = new dcol_group();
  for (int i = 0; i < n_devices; i++) {
    dcol_ctx* c = nullptr;
    const int rc = dcol_create(&c, device_ids[i]);
    if (rc) {
      g_group_create_err = std::string("dcol_group_create: member ") + std::to_string(i) + ": " + dcol_last_error(nullptr);
      dcol_group_destroy(g);
      return rc;
    }
    // the members copy concurrently: share the host's copy threads between them
    const unsigned hw = std::max(2u, std::thread::hardware_concurrency());
    c->copy_threads = (int)std::max(2u, std::min(8u, hw / 2) / (unsigned)n_devices);
    g->members.push_back(c);
  }
  *out = g;
  return 0;
}

void dcol_group_destroy(dcol_group* g) {
  if (!g) return;
  for (dcol_ctx* c : g->members) dcol_destroy(c);
  delete g;
}

const char* dcol_group_last_error(const dcol_group* g) { return g ? g->err.c_str() : g_group_create_err.c_str(); }
int dcol_group_size(const dcol_group* g) { return g ? (int)g->members.size() : 0; }
dcol_ctx* dcol_group_ctx(dcol_group* g, int member) { return (g && member >= 0 && member < (int)g->members.size()) ? g->members[member] : nullptr; }
double dcol_group_last_kernel_ms(const dcol_group* g) { return g ? g->last_ms : -1.0; }

int dcol_group_set_shapes(dcol_group* g, int n_shapes, const int32_t* type, const int64_t* param_off, const double* params, int64_t n_params) {
  if (!g) return -1;
  g->err.clear();
  for (size_t i = 0; i < g->members.size(); i++) {
    const int rc = dcol_set_shapes(g->members[i], n_shapes, type, param_off, params, n_params);
    if (rc) { g->err = "member " + std::to_string(i) + ": " + dcol_last_error(g->members[i]); return rc; }
  }
  return 0;
}

int dcol_group_query(dcol_group* g, int query, int64_t n, const int32_t* shape1, const int32_t* shape2,
                     const double* r1, const double* att1, const double* r2, const double* att2,
                     int att_kind, double pdip_tol,
                     double* alpha, double* x, double* deriv, int32_t* status, int32_t* iters) {
  if (!g) return -1;
  g->err.clear(); g->last_ms = -1.0;
  if (query != Q_PROX && query != Q_GRAD && query != Q_JAC) { g->err = "dcol_group_query: query must be 0 (proximity), 1 (gradient) or 2 (jacobian)"; return -1; }
  if (att_kind != DCOL_ATT_QUAT && att_kind != DCOL_ATT_MRP) { g->err = "dcol: att_kind must be 0 (quaternion) or 1 (MRP)"; return -1; }
  if (n < 0) { g->err = "dcol: n out of range"; return -1; }
  const int m = (int)g->members.size();
  const int na = (att_kind == DCOL_ATT_QUAT) ? 4 : 3, nt = 6 + 2 * na;
  const int64_t dsz = (query == Q_GRAD) ? nt : (query == Q_JAC ? 4 * nt : 0);
  std::vector<int> rc(m, 0);
  auto slice = [&](int i) {
    const int64_t base = n / m, rem = n % m;
    const int64_t b = i * base + std::min<int64_t>(i, rem), e = b + base + (i < rem ? 1 : 0);
    if (e == b) return;
    // (argument checks, including NULL pointers, are run_query's: offsetting a NULL base is avoided here)
    auto off = [](auto* p, int64_t k) { return p ? p + k : p; };
    rc[i] = run_query(g->members[i], query, e - b, off(shape1, b), off(shape2, b), off(r1, 3 * b), off(att1, na * b), off(r2, 3 * b), off(att2, na * b),
                      att_kind, pdip_tol, DCOL_MEM_HOST, off(alpha, b), off(x, 3 * b), off(deriv, dsz * b), off(status, b), off(iters, b));
  };
  if (m == 1 || n < 2) {
    for (int i = 0; i < m; i++) slice(i);
  } else {
    std::vector<std::thread> th;
    for (int i = 1; i < m; i++) th.emplace_back(slice, i);
    slice(0);
    for (std::thread& t : th) t.join();
  }
  for (int i = 0; i < m; i++) {
    g->last_ms = std::max(g->last_ms, g->members[i]->last_ms);
    if (rc[i]) { g->err = "member " + std::to_string(i) + " (device " + std::to_string(g->members[i]->device) + "): " + g->members[i]->err; return rc[i]; }
  }
  return 0;
}

// Runs f(member, begin, end) on every member's contiguous slice of [0, n), one host thread per member.
extern "C++" {
template <class F>
static int group_slices(dcol_group* g, int64_t n, F&& f) {
  const int m = (int)g->members.size();
  std::vector<int> rc(m, 0);
  auto slice = [&](int i) {
    const int64_t base = n / m, rem = n % m;
    const int64_t b = i * base + std::min<int64_t>(i, rem), e = b + base + (i < rem ? 1 : 0);
    rc[i] = f(i, b, e);
  };
  if (m == 1) slice(0);
  else {
    std::vector<std::thread> th;
    for (int i = 1; i < m; i++) th.emplace_back(slice, i);
    slice(0);
    for (std::thread& t : th) t.join();
  }
  g->last_ms = -1.0;
  for (int i = 0; i < m; i++) {
    g->last_ms = std::max(g->last_ms, g->members[i]->last_ms);
    if (rc[i]) { g->err = "member " + std::to_string(i) + " (device " + std::to_string(g->members[i]->device) + "): " + g->members[i]->err; return rc[i]; }
  }
  return 0;
}
}  // extern "C++"

// All-pairs scene over the group: the body arrays go to every member, member i solves the i-th
// contiguous slice of the i < j enumeration and writes its rows of the caller's (host) arrays.
int dcol_group_scene_allpairs(dcol_group* g, int query, int n_bodies, const int32_t* shape, const double* r, const double* att,
                              int att_kind, double pdip_tol,
                              double* alpha, double* x, double* deriv, int32_t* status, int32_t* iters) {
  if (!g) return -1;
  g->err.clear();
  if (query != Q_PROX && query != Q_GRAD && query != Q_JAC) { g->err = "dcol_group_scene_allpairs: query must be 0 (proximity), 1 (gradient) or 2 (jacobian)"; return -1; }
  if (att_kind != DCOL_ATT_QUAT && att_kind != DCOL_ATT_MRP) { g->err = "dcol: att_kind must be 0 (quaternion) or 1 (MRP)"; return -1; }
  const int na = (att_kind == DCOL_ATT_QUAT) ? 4 : 3, nt = 6 + 2 * na;
  const int64_t dsz = (query == Q_GRAD) ? nt : (query == Q_JAC ? 4 * nt : 0);
  auto off = [](auto* p, int64_t k) { return p ? p + k : p; };
  return group_slices(g, dcol_scene_pair_count(n_bodies), [&](int i, int64_t b, int64_t e) {
    if (e == b) return 0;
    return dcol_scene_allpairs_range(g->members[i], query, n_bodies, b, e, shape, r, att, att_kind, pdip_tol, DCOL_MEM_HOST,
                                     off(alpha, b), off(x, 3 * b), off(deriv, dsz * b), off(status, b), off(iters, b));
  });
}

// Contact-scene accumulation over the group: every member reduces the pairs of its slice to per-body partial sums on its
// device; the partial sums (n_bodies x 6 per member) are added on the host in member order -- the exchange step of
// this operator.  The result differs from the one-device call by the summation order only.
int dcol_group_scene_wrench(dcol_group* g, int n_bodies, const int32_t* shape, const double* r, const double* att, const double* lambda,
                            int att_kind, double pdip_tol,
                            double* alpha, double* wrench, int32_t* status, int32_t* iters) {
  if (!g) return -1;
  g->err.clear();
  if (n_bodies < 0 || !wrench || !lambda) { g->err = "dcol_group_scene_wrench: bad arguments"; return -1; }
  const int m = (int)g->members.size();
  std::vector<double> part((size_t)m * 6 * (size_t)std::max(n_bodies, 0), 0.0);
  auto off = [](auto* p, int64_t k) { return p ? p + k : p; };
  const int rc = group_slices(g, dcol_scene_pair_count(n_bodies), [&](int i, int64_t b, int64_t e) {
    if (e == b) return 0;
    return dcol_scene_wrench_range(g->members[i], n_bodies, b, e, shape, r, att, lambda + b, att_kind, pdip_tol, DCOL_MEM_HOST,
                                   off(alpha, b), part.data() + (size_t)i * 6 * n_bodies, off(status, b), off(iters, b));
  });
  if (rc) return rc;
  for (int64_t k = 0; k < 6 * (int64_t)n_bodies; k++) {
    double a = 0.0;
    for (int i = 0; i < m; i++) a += part[(size_t)i * 6 * n_bodies + k];
    wrench[k] = a;
  }
  return 0;
}

}  // extern "C"
