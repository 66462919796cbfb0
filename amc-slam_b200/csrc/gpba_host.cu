// gpba_host.cu -- host side of libgpba.so: problem upload, structure builder (P0), LM controller (K7),
// and the C ABI of include/gpba.h.  Mirrors the reference's plugin surface:
//   g2o::BlockSolver<Traits>          (Thirdparty/g2o/g2o/core/block_solver.h:96-176, block_solver.hpp)  -> struct Solver, L1 calls
//   g2o::OptimizationAlgorithmLevenberg (optimization_algorithm_levenberg.cpp:61-194)                  -> Solver::lm_solve
//   g2o::SparseOptimizer::optimize    (sparse_optimizer.cpp:354-419)                                     -> Solver::optimize
// There is NO CPU fallback: every entry point fails with GPBA_ERR_NO_DEVICE / GPBA_ERR_CUDA if the
// device path is unavailable.
#include "../../include/gpba.h"
#include "gpba_chol.cuh"
#include "gpba_pcg.cuh"
#include "gpba_structure.cuh"
#include "gpba_order.h"
#include "gpba_pose.cuh"
#include "gpba_vel.cuh"
#include "gpba_posegraph.cuh"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <limits>
#include <string>
#include <vector>
#include <dlfcn.h>
#include <functional>
#include <future>
#include <map>
#include <memory>
#include <mutex>
#include <chrono>

using namespace gpba;

static thread_local std::string g_err;
#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess) {                                                                       \
      g_err = std::string(#call) + ": " + cudaGetErrorString(e_) + " (" __FILE__ ":" + std::to_string(__LINE__) + ")"; \
      return GPBA_ERR_CUDA;                                                                        \
    }                                                                                              \
  } while (0)
#define CKR(call)                   \
  do {                              \
    int r_ = (call);                \
    if (r_ != GPBA_OK) return r_;   \
  } while (0)

namespace {

// Device buffers come from the device's stream-ordered memory pool (cudaMallocAsync) on the handle's stream: the pool
// keeps freed blocks, so building the structures of the next optimize() call does not pay cudaMalloc / cudaFree again
// (they were ~200 ms of a 380 ms build_structure at C4).  g_alloc_stream is set by every entry point of a handle.
static thread_local cudaStream_t g_alloc_stream = nullptr;

// A SLAM process runs one BA after the other, each with a fresh handle (like a fresh g2o::SparseOptimizer) and nearly the same
// buffer sizes.  cudaMallocAsync's pool re-uses freed memory, but a create / destroy cycle of ~4 GB in ~150 blocks made it
// re-map physical memory unpredictably (measured: gpba_build_structure 17 .. 46 ms for the same problem, one call of 967 ms).
// Freed blocks are therefore parked in a process-wide best-fit cache; a block taken on another stream than the one it was
// released on first waits for the event recorded at release time, so the cache is as stream-ordered as cudaFreeAsync is.
// Total parked memory is capped; beyond the cap blocks go back to the driver pool.
struct BlockCache {
  struct Block { void* p; size_t bytes; int device; cudaStream_t stream; cudaEvent_t ev; };
  std::mutex m;
  std::multimap<size_t, Block> free_blocks;   // by capacity
  size_t parked = 0;
  static constexpr size_t kCap = 24ull << 30;   // parked bytes per process
  static size_t round_up(size_t b) { return b < (1u << 20) ? ((b + 511) & ~size_t(511)) : ((b + (1u << 20) - 1) & ~size_t((1u << 20) - 1)); }
  cudaError_t get(void** out, size_t bytes, size_t* capacity, cudaStream_t s) {
    bytes = round_up(bytes);
    int dev = 0;
    cudaGetDevice(&dev);
    {
      std::lock_guard<std::mutex> lock(m);
      for (auto it = free_blocks.lower_bound(bytes); it != free_blocks.end() && it->first <= bytes + bytes / 4 + (1u << 20); ++it) {
        if (it->second.device != dev) continue;
        Block b = it->second;
        free_blocks.erase(it);
        parked -= b.bytes;
        if (b.stream != s) cudaStreamWaitEvent(s, b.ev, 0);
        cudaEventDestroy(b.ev);
        *out = b.p; *capacity = b.bytes;
        return cudaSuccess;
      }
    }
    *capacity = bytes;
    return cudaMallocAsync(out, bytes, s);
  }
  void put(void* p, size_t bytes, cudaStream_t s) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaEvent_t ev = nullptr;
    std::unique_lock<std::mutex> lock(m);
    if (parked + bytes > kCap || cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) { lock.unlock(); cudaFreeAsync(p, s); return; }
    cudaEventRecord(ev, s);
    free_blocks.emplace(bytes, Block{p, bytes, dev, s, ev});
    parked += bytes;
  }
};
static BlockCache g_blocks;

template <typename T>
struct DBuf {
  T* p = nullptr;
  size_t n = 0;
  size_t cap_bytes = 0;
  cudaStream_t owner = nullptr;
  ~DBuf() { release(); }
  void release() { if (p) g_blocks.put(p, cap_bytes, owner); p = nullptr; n = 0; cap_bytes = 0; }
  int alloc(size_t count) {
    if (count <= n && p) return GPBA_OK;
    release();
    if (count == 0) count = 1;
    owner = g_alloc_stream;
    void* q = nullptr;
    CK(g_blocks.get(&q, count * sizeof(T), &cap_bytes, owner));
    p = (T*)q;
    n = count;
    return GPBA_OK;
  }
  int upload(const std::vector<T>& v, cudaStream_t s) {
    CKR(alloc(v.size()));
    if (!v.empty()) CK(cudaMemcpyAsync(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, s));
    return GPBA_OK;
  }
  int upload(const T* v, size_t count, cudaStream_t s) {
    CKR(alloc(count));
    if (count) CK(cudaMemcpyAsync(p, v, count * sizeof(T), cudaMemcpyHostToDevice, s));
    return GPBA_OK;
  }
};

// ---- NCCL through dlopen: the library stays loadable without NCCL; multi-GPU needs libnccl.so.2
typedef struct ncclComm* ncclComm_t;
struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(void*) = nullptr;
  int (*CommInitRank)(ncclComm_t*, int, unsigned char[128], int) = nullptr;  // ncclUniqueId passed by value (128 B)
  int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*Broadcast)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool load() {
    if (lib) return true;
    lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) return false;
    GetUniqueId = (int (*)(void*))dlsym(lib, "ncclGetUniqueId");
    AllReduce = (int (*)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t))dlsym(lib, "ncclAllReduce");
    Broadcast = (int (*)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t))dlsym(lib, "ncclBroadcast");
    CommDestroy = (int (*)(ncclComm_t))dlsym(lib, "ncclCommDestroy");
    GetErrorString = (const char* (*)(int))dlsym(lib, "ncclGetErrorString");
    return GetUniqueId && AllReduce && Broadcast && CommDestroy && dlsym(lib, "ncclCommInitRank");
  }
};
static NcclApi g_nccl;
enum { kNcclDouble = 8, kNcclSum = 0, kNcclMax = 2 };   // ncclDataType_t / ncclRedOp_t values of nccl.h
struct NcclId { char internal[128]; };
typedef int (*ncclCommInitRank_t)(ncclComm_t*, int, NcclId, int);

// Pinned scratch for the few scalars a handle reads back per LM trial.  cudaMallocHost / cudaFreeHost synchronise the
// device and map / unmap pages (tens of ms when memory is busy), so handles borrow 256-byte slots of one block that is
// allocated once per process.
struct PinnedSlots {
  static constexpr int N = 256, BYTES = 256;
  std::mutex m;
  char* base = nullptr;
  bool used[N] = {false};
  void* take() {
    std::lock_guard<std::mutex> lock(m);
    if (!base && cudaMallocHost(&base, (size_t)N * BYTES) != cudaSuccess) { base = nullptr; return nullptr; }
    for (int i = 0; i < N; ++i) if (!used[i]) { used[i] = true; return base + (size_t)i * BYTES; }
    return nullptr;
  }
  void give(void* p) {
    if (!p) return;
    std::lock_guard<std::mutex> lock(m);
    used[((char*)p - base) / BYTES] = false;
  }
};
static PinnedSlots g_pinned;

// The frame-rate entry points (gpba_pose_optimize, gpba_vel_ransac) are called a few times per frame from one tracking
// thread: they keep one stream per (host thread, device) instead of creating and destroying one per call.  The streams
// live as long as the process (destroying CUDA objects from thread-exit handlers races with runtime teardown).
static cudaStream_t small_call_stream(int device) {
  static thread_local std::map<int, cudaStream_t> streams;
  auto it = streams.find(device);
  if (it != streams.end()) return it->second;
  cudaStream_t s = nullptr;
  if (cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
  streams.emplace(device, s);
  return s;
}

struct StreamHolder {
  cudaStream_t s = nullptr;
  ~StreamHolder() { if (s) { cudaStreamSynchronize(s); cudaStreamDestroy(s); } }
};

// Host part of the factorization's structure: symbolic phase (gpba_order.h) and the schedule tables derived from it.  No device
// call in here: build_structure runs it on a second host thread while the main thread finishes the device-side structure.
struct CholHost {
  CholSymbolic sym;
  std::vector<int2> pan_tab, back_tab;
  std::vector<int4> upd_tab, cf_tab, prod;
  std::vector<unsigned> klist;
  std::vector<int> cf_need, lvl_pan_begin, lvl_upd_begin, lvl_back_begin, lvl_ncols;
  int64_t products = 0, update_ctas = 0;
  int rc = GPBA_OK;
  std::string err;
};
static int chol_host_phase(int n_pose, int n_hs, const int* hs_row, const int* hs_col, CholHost& H) {
  const int n = n_pose * 12;
  const int bpt = GPBA_NB / 12;  // pose blocks per tile
  // symbolic phase on the host (gpba_order.h): nested-dissection order, tile-level fill, level schedule
  CholSymbolic& sym = H.sym;
  chol_symbolic(n_pose, n_hs, hs_row, hs_col, bpt, GPBA_TILE, getenv("GPBA_CHOL_ND_DEPTH") ? atoi(getenv("GPBA_CHOL_ND_DEPTH")) : -1,
                getenv("GPBA_VERBOSE") != nullptr, sym);
  const int NT = sym.NT;
  const int64_t chol_doubles = sym.doubles;
  const std::vector<int>& chol_col_begin = sym.col_begin;
  const std::vector<int>& chol_row_begin = sym.row_begin;
  const std::vector<int>& col_rows = sym.col_rows;
  const std::vector<int>& perm = sym.perm;
  const std::vector<int64_t>& off = sym.tile_off;
  if (getenv("GPBA_VERBOSE")) fprintf(stderr, "[gpba] cholesky: n=%d NT=%d tiles=%lld (dense lower would be %d), %.1f MB\n", n, NT, (long long)(chol_doubles / GPBA_TILE), NT * (NT + 1) / 2, chol_doubles * 8 / 1e6);
  const int n_levels = sym.n_levels;
  const std::vector<int>& level = sym.level;
  std::vector<std::vector<int>> lvl_cols(n_levels);
  for (int k = 0; k < NT; ++k) lvl_cols[level[k]].push_back(k);
  std::vector<int2>&pan_tab = H.pan_tab, &back_tab = H.back_tab;
  std::vector<int4>& upd_tab = H.upd_tab;     // left-looking update: {column j, tile slot q, klist begin, klist end}
  std::vector<unsigned>& klist = H.klist;   // finished columns k with L_ik != 0 and L_jk != 0 per tile (i, j); bit 31: column of the previous level
  std::vector<int>&lvl_pan_begin = H.lvl_pan_begin, &lvl_upd_begin = H.lvl_upd_begin, &lvl_back_begin = H.lvl_back_begin, &lvl_ncols = H.lvl_ncols;
  lvl_pan_begin.assign(n_levels + 1, 0); lvl_upd_begin.assign(n_levels + 1, 0); lvl_back_begin.assign(n_levels + 1, 0);
  lvl_ncols.assign(n_levels, 0);
  std::vector<unsigned> late_scratch;
  std::vector<int4> early_chunks, late_chunks;
  std::vector<int4>& cf_tab = H.cf_tab;
  std::vector<int>& cf_need = H.cf_need;
  cf_need.assign(2 * (size_t)NT, 0);
  for (int l = 0; l < n_levels; ++l) {
    lvl_ncols[l] = (int)lvl_cols[l].size();
    early_chunks.clear(); late_chunks.clear();
    struct TileList { int k, q, kb, cnt, n_late; };
    std::vector<TileList> lists;
    for (int k : lvl_cols[l]) {
      const int nr = chol_col_begin[k + 1] - chol_col_begin[k];
      for (int q = 0; q <= nr; ++q) pan_tab.push_back(make_int2(k, q));
      // tile (i, k) of column k receives L_ic L_kc^T from every finished column c that touches both rows: the
      // intersection of the two row patterns (both ascending)
      for (int q = 0; q <= nr; ++q) {
        const int i = q == 0 ? k : col_rows[chol_col_begin[k] + q - 1];
        const int *pa = sym.row_cols.data() + chol_row_begin[i], *pae = sym.row_cols.data() + chol_row_begin[i + 1];
        const int *pb = sym.row_cols.data() + chol_row_begin[k], *pbe = sym.row_cols.data() + chol_row_begin[k + 1];
        const int kb = (int)klist.size();
        // sources of older levels first, sources of the previous level (whose panel step overlaps with this update) last
        size_t n_late = 0;
        std::vector<unsigned>& late = late_scratch;
        late.clear();
        while (pa < pae && pb < pbe) {
          if (*pa < *pb) ++pa; else if (*pb < *pa) ++pb;
          else { if (level[*pa] == l - 1) late.push_back((unsigned)*pa | GPBA_LU_LATE); else klist.push_back((unsigned)*pa); ++pa; ++pb; }
        }
        n_late = late.size();
        klist.insert(klist.end(), late.begin(), late.end());
        const int cnt = (int)klist.size() - kb;
        if (cnt == 0) continue;
        lists.push_back({k, q, kb, cnt, (int)n_late});
      }
      const int nrow = chol_row_begin[k + 1] - chol_row_begin[k];
      for (int q = 0; q <= nrow; ++q) back_tab.push_back(make_int2(k, q));
      if (nr >= 65536) { H.err = "tile column with more than 65535 rows"; return GPBA_ERR_INVALID; }
    }
    // Chunk size: at least GPBA_LU_CHUNK products (a chunk ends with 12 reductions per thread, a fence and a counter
    // update), more where the level has enough products to keep every SM busy with fewer, longer chunks.
    {
      static const int lu_chunk = getenv("GPBA_LU_CHUNK") ? std::max(1, atoi(getenv("GPBA_LU_CHUNK"))) : GPBA_LU_CHUNK;
      static const int lu_chunk_max = getenv("GPBA_LU_CHUNK_MAX") ? std::max(1, atoi(getenv("GPBA_LU_CHUNK_MAX"))) : 12;
      int64_t total = 0;
      for (const TileList& t : lists) total += t.cnt;
      const int chunk = (int)std::min<int64_t>(std::max<int64_t>(lu_chunk, total / (2 * 148)), std::max(lu_chunk, lu_chunk_max));
      // products whose source column belongs to the previous level sit on the critical path (panel -> product -> next
      // panel): they get short chunks of their own so that several SMs work on one tile's late tail at once
      static const int late_chunk = getenv("GPBA_LU_LATE_CHUNK") ? std::max(0, atoi(getenv("GPBA_LU_LATE_CHUNK"))) : 1;
      auto cut = [&](const TileList& t, int b, int cnt, int size, bool late) {
        const int nch = (cnt + size - 1) / size;
        for (int c = 0; c < nch; ++c)
          (late ? late_chunks : early_chunks).push_back(make_int4(t.k, t.q, b + (int)((int64_t)cnt * c / nch), b + (int)((int64_t)cnt * (c + 1) / nch)));
      };
      for (const TileList& t : lists) {
        if (late_chunk > 0) {
          if (t.cnt > t.n_late) cut(t, t.kb, t.cnt - t.n_late, chunk, false);
          if (t.n_late > 0) cut(t, t.kb + t.cnt - t.n_late, t.n_late, late_chunk, true);
          continue;
        }
        const int nch = (t.cnt + chunk - 1) / chunk;
        for (int c = 0; c < nch; ++c) {
          const int cb = t.kb + (int)((int64_t)t.cnt * c / nch), ce = t.kb + (int)((int64_t)t.cnt * (c + 1) / nch);
          const bool is_late = ce > t.kb + t.cnt - t.n_late;   // the chunk reaches into the late tail
          (is_late ? late_chunks : early_chunks).push_back(make_int4(t.k, t.q, cb, ce));
        }
      }
    }
    // chunks that only need finished levels are handed out first: they overlap with the previous level's panel step
    upd_tab.insert(upd_tab.end(), early_chunks.begin(), early_chunks.end());
    upd_tab.insert(upd_tab.end(), late_chunks.begin(), late_chunks.end());
    // the same work as one task list for the persistent kernel: this level's chunks, then its panel tasks
    auto tile_of = [&](int i, int j) { return (int)(off[(size_t)i * NT + j] / GPBA_TILE); };
    auto row_of = [&](int k, int q) { return q == 0 ? k : col_rows[chol_col_begin[k] + q - 1]; };
    for (const std::vector<int4>* v : {&early_chunks, &late_chunks})
      for (const int4& c : *v) {
        cf_tab.push_back(make_int4(c.x | (c.y == 0 ? GPBA_CF_DIAG : 0), tile_of(row_of(c.x, c.y), c.x), c.z, c.w));
        cf_need[c.x] += 1;
      }
    // Wide levels (more panel tasks than resident CTAs can take at once) are throughput bound: one task per column
    // factorizes the diagonal tile and publishes it, the others only solve.  Narrow levels are latency bound: every task
    // factorizes the diagonal tile itself, which saves a publish / wait / reload hop on the critical path.
    static const int split_min = getenv("GPBA_CF_SPLIT_MIN") ? atoi(getenv("GPBA_CF_SPLIT_MIN")) : 120;
    const bool split = lvl_pan_begin.size() > 0 && ((int)pan_tab.size() - lvl_pan_begin[l]) >= split_min;
    for (int pass = 0; pass < 2; ++pass)   // diagonal tasks first
      for (int k : lvl_cols[l]) {
        const int nr = chol_col_begin[k + 1] - chol_col_begin[k];
        if (pass == 0) { cf_tab.push_back(make_int4(k | GPBA_CF_DIAG, tile_of(k, k), split && nr > 0 ? -2 : -1, tile_of(k, k))); cf_need[(size_t)NT + k] = nr + 1; }
        else for (int q = 1; q <= nr; ++q) cf_tab.push_back(make_int4(k, tile_of(row_of(k, q), k), split ? -3 : -1, tile_of(k, k)));
      }
    lvl_pan_begin[l + 1] = (int)pan_tab.size(); lvl_upd_begin[l + 1] = (int)upd_tab.size(); lvl_back_begin[l + 1] = (int)back_tab.size();
  }
  H.products = (int64_t)klist.size(); H.update_ctas = (int64_t)upd_tab.size();
  if (upd_tab.empty()) upd_tab.push_back(make_int4(0, 0, 0, 0));
  if (klist.empty()) klist.push_back(0);
  // one record per product: the tiles of L_ik and L_jk and the source column k
  H.prod.assign(klist.size(), make_int4(0, 0, 0, 0));
  for (const int4& c : upd_tab) {
    if (c.z >= c.w) continue;
    const int j = c.x, i = c.y == 0 ? j : col_rows[chol_col_begin[j] + c.y - 1];
    for (int p = c.z; p < c.w; ++p) {
      const int k = (int)(klist[p] & ~GPBA_LU_LATE);
      H.prod[p] = make_int4((int)(off[(size_t)i * NT + k] / GPBA_TILE), (int)(off[(size_t)j * NT + k] / GPBA_TILE), k, 0);
    }
  }
  return GPBA_OK;
}

struct Solver {
  int device = 0;
  StreamHolder stream_holder;   // first member: destroyed after every buffer has been returned to the pool
  StreamHolder copy_holder;     // second stream of the asynchronous measurement upload (GPBA_CREATE_ASYNC_UPLOAD)
  cudaStream_t stream = nullptr;
  cudaEvent_t ev_meas = nullptr;
  bool meas_pending = false;    // the measurement copies have been enqueued on copy_holder.s and not yet waited for
  // ------------------------------------------------------------------ host copy of the problem
  int n_cam = 0, n_kf = 0, n_pt = 0, n_rec = 0, n_prior = 0, n_velp = 0;
  int64_t n_obs = 0;
  bool stereo = false;
  std::vector<double> h_pose, h_vel, h_pt, h_time;
  std::vector<uint8_t> kf_fixed, obs_flags;
  std::vector<int> rec_kf1, rec_kf2, rec_cam, prior_kf1, prior_kf2, velp_kf;   // per-observation indices live on the device only
  std::vector<double> rec_t;   // per-observation measurements live on the device only (d_all_*)
  double lambda_init = 0;
  int linear_solver = 0;
  // ------------------------------------------------------------------ device: static
  DBuf<CamConst> d_cam[2];                      // per state buffer: the extrinsics are part of the state when they are free
  DBuf<double> d_ext[2];                        // [n_cam][7] Tbc (VertexExtrinsic estimate), double buffered like the poses
  std::vector<double> h_ext0;                   // the extrinsics the handle was created with
  std::vector<CamConst> h_cam0;
  // extrinsic self-calibration (LocalGPBA's second stage, src/Optimizer.cc:983-995, 1228-1240)
  std::vector<uint8_t> ext_free, ext_prior_on;  // [n_cam]
  std::vector<double> ext_prior_qinv, ext_prior_info;
  std::vector<int> ext_h;                       // [n_cam] hessian index (after the keyframes) or -1
  DBuf<int> d_ext_h;
  DBuf<unsigned char> d_ext_prior_on;
  DBuf<double> d_ext_prior_qinv, d_ext_prior_info;
  std::vector<std::vector<double>> stack_ext;
  int n_pose_kf = 0;
  DBuf<double> d_time, d_rec_t;
  DBuf<int> d_rec_kf1, d_rec_kf2, d_rec_cam, d_prior_kf1, d_prior_kf2, d_velp_kf;
  DBuf<double> d_pt_full;                       // [n_pt*3] always-current copy in original order
  DBuf<double> d_chi2;                          // [n_obs] stored edge chi2 (original order)
  DBuf<double> d_all_u, d_all_v, d_all_w, d_all_ur; DBuf<int> d_all_rec, d_all_pt; DBuf<uint8_t> d_all_flags;  // original-order obs
  // ------------------------------------------------------------------ device: state (double buffered)
  DBuf<double> d_pose[2], d_vel[2], d_ptS[2];   // d_ptS: landmarks in sorted order
  int cur = 0;
  int last_eval = 0;                            // buffer the last error evaluation ran on
  double* chi2_store_override = nullptr;
  bool pts_gathered = false;                    // multi-GPU: d_pt_full holds all ranks' landmarks (valid until the state changes)
         // multi-GPU: per-edge chi2 goes to the staging buffer first
  std::vector<std::vector<double>> stack_pose, stack_vel, stack_pt;  // push/pop (L1)
  // ------------------------------------------------------------------ structure (host)
  bool structure_ok = false, system_ok = false;
  gpba_structure_info info{};
  std::vector<int> kf_h, lm_pt, pt_lm, lm_rank;  // lm_rank[sorted lm] = landmark index in g2o order (ascending point id)
  std::vector<int64_t> lm_obs_begin, o_orig;
  std::vector<int> hpp_row, hpp_col, hs_row, hs_col;
  int64_t n_aobs = 0;
  int n_lm = 0, n_lm_all = 0, n_pose = 0, n_hpp = 0, n_hs = 0, n_items = 0, n_rseg = 0, n_rp = 0;   // n_lm: own landmarks, n_lm_all: all ranks'
  int64_t n_con = 0;
  int64_t n_hpl = 0, n_pairs = 0;
  // ------------------------------------------------------------------ structure (device)
  DBuf<int> d_kf_h, d_o_rec, d_o_lm, d_lm_pt, d_rseg_rec, d_item_rp, d_con_begin;
  DBuf<double> d_o_u, d_o_v, d_o_ur, d_o_w, d_r_u, d_r_v, d_r_ur, d_r_w;   // landmark-major and record-major copies
  DBuf<int> d_r_lm; DBuf<uint8_t> d_r_flags;
  DBuf<uint8_t> d_o_flags, d_item_flags;
  DBuf<int64_t> d_o_orig, d_lm_obs_begin, d_rperm, d_rseg_begin, d_item_begin, d_item_end;
  DBuf<unsigned long long> d_pairs;     // observation pairs grouped by record pair (sorted order)
  DBuf<unsigned long long> d_rp_key;    // unique record pairs (r1 * n_rec + r2), ascending
  DBuf<int> d_pt_act;                   // [n_pt] point carries an active edge (on any rank)
  DBuf<double> d_gather;                // multi-GPU: staging buffer of the end-of-optimize all-reduces
  DBuf<HsContrib> d_con;
  CubTemp cub_tmp;
  DBuf<int> d_rec_hpp13, d_rec_hpp23, d_rec_hpp33;
  DBuf<int> d_rec_hpp11, d_rec_hpp12, d_rec_hpp22, d_prior_hpp11, d_prior_hpp12, d_prior_hpp22, d_pose_hpp_diag;
  DBuf<int> d_hs_from_hpp, d_hs_diag_pose, d_hs_row, d_hs_col;
  // ------------------------------------------------------------------ Hessian storage (device)
  DBuf<double> d_rec, d_rec_lite, d_recS, d_hll, d_bl, d_W, d_U, d_C, d_Y, d_ptL, d_hpp, d_bp, d_hs, d_bs, d_x, d_xl;
  DBuf<double> d_partial, d_prior_rho, d_pose_scale, d_scal;
  DBuf<int> d_fail;
  double* h_scal = nullptr;   // pinned: [0] chi2 [1] scale [2..] spare
  int* h_fail = nullptr;      // pinned
  int grid_obs = 0;
  // ------------------------------------------------------------------ Cholesky
  int NT = 0;
  int64_t chol_doubles = 0;
  DBuf<int64_t> d_tile_off;
  DBuf<int> d_col_begin, d_col_rows, d_chol_perm;
  DBuf<unsigned char> d_chol_pos_used;   // [NT * 4] position (pose block slot) of the permuted system holds a pose block
  DBuf<double> d_tiles, d_chol_work, d_chol_dinv, d_chol_x;
  DBuf<int> d_row_begin, d_row_cols;
  std::vector<int> chol_row_begin;
  // level schedule of the tile columns (columns of one level are independent) and the CTA tables of its launches
  std::vector<int> lvl_pan_begin, lvl_upd_begin, lvl_back_begin, lvl_ncols;
  DBuf<int2> d_pan_tab, d_back_tab;
  DBuf<int4> d_upd_tab;
  DBuf<int4> d_cf_prod;              // persistent factorization: one record per tile product
  DBuf<int4> d_cf_tab;               // persistent factorization: all tasks in level order (update chunks, then panel tasks, per level)
  DBuf<int> d_cf_need;               // [2 NT]: upd_need | pan_need
  DBuf<int> d_cf_cnt;                // [2 NT + 1]: upd_cnt | pan_cnt | task counter (zeroed by the graph)
  int cf_tasks = 0;
  std::future<std::unique_ptr<CholHost>> chol_future;   // host tables of the factorization, computed beside the device-side structure build
  DBuf<double> d_cf_d8;              // [NT][6][8][8] inverted diagonal blocks published by the diagonal tasks of wide levels
  DBuf<long long> d_cf_trace;        // GPBA_CF_TRACE=<file>: per-task timestamps of the last factorization (tools/cf_trace.py)
  DBuf<int> d_tile_lm, d_tile_rlo, d_tile_rcnt;   // landmark-aligned observation tiles + their record windows (K1 / K2a)
  int n_tiles = 0;
  DBuf<unsigned> d_klist;
  DBuf<int> d_lu_counter;            // [2 * levels]: chunk counter of every level's update, then completion counter of every level's panel grid
  int chol_parts = 1;
  int64_t chol_products = 0, chol_update_ctas = 0;
  cudaGraphExec_t chol_graph = nullptr;       // load + (panel, update) x NT, captured once per structure
  cudaGraphExec_t chol_back_graph = nullptr;  // backward substitution, one launch per tile row
  int chol_graph_launches = 0, chol_back_launches = 0;
  CholView chol_view() {
    return CholView{NT, n_pose * 12, d_tile_off.p, d_col_begin.p, d_col_rows.p, d_row_begin.p, d_row_cols.p, d_tiles.p,
                    d_chol_perm.p, d_chol_pos_used.p, d_chol_dinv.p, d_chol_work.p, d_chol_x.p};
  }
  std::vector<int> chol_col_begin;
  // ------------------------------------------------------------------ PCG
  PcgBuffers pcg;
  // ------------------------------------------------------------------ LM state
  double lambda_cur = -1, ni = 2, lambda_set = 0;
  int nBad = 0;
  bool lambda_applied = false;
  double huber_mono = 0, huber_stereo = 0, huber_prior = 0, qc[6], bf = 0;
  // ------------------------------------------------------------------ distribution
  int rank = 0, nranks = 1;
  ncclComm_t comm = nullptr;
  DBuf<double> d_red;   // packed [Hschur values | bschur | chi2] for the allreduce
  // ------------------------------------------------------------------ profiling
  bool profiling = false;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  double stage_ms[GPBA_N_STAGES] = {0};
  int64_t stage_launches[GPBA_N_STAGES] = {0};
  // event pairs recorded on the library stream around every stage; elapsed times are read back
  // (cudaEventElapsedTime) only in gpba_stage_stats, so the timed region is never synchronised by profiling
  struct EvPair { cudaEvent_t a, b; int stage; };
  std::vector<EvPair> ev_pool;
  size_t ev_used = 0;
  bool structure_dirty = true;

  DevView V{};

  ~Solver() {
    // the communicator is shared by every handle created with the same NCCL id (g_comms) and lives until process exit
    if (chol_graph) cudaGraphExecDestroy(chol_graph);
    if (chol_back_graph) cudaGraphExecDestroy(chol_back_graph);
    g_pinned.give(h_scal);
    if (ev0) cudaEventDestroy(ev0);
    if (ev1) cudaEventDestroy(ev1);
    for (auto& e : ev_pool) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
    // the DBuf members free into `stream` after this body ran: the stream object itself is released by StreamHolder,
    // which is declared first and therefore destroyed last
  }

  // stage timing helpers
  void t0() {
    if (!profiling) return;
    if (ev_used == ev_pool.size()) {
      EvPair e;
      cudaEventCreate(&e.a); cudaEventCreate(&e.b); e.stage = -1;
      ev_pool.push_back(e);
    }
    cudaEventRecord(ev_pool[ev_used].a, stream);
  }
  void t1(int stage, int launches) {
    stage_launches[stage] += launches;
    if (!profiling) return;
    ev_pool[ev_used].stage = stage;
    cudaEventRecord(ev_pool[ev_used].b, stream);
    ++ev_used;
  }
  void collect_events() {
    if (ev_used == 0) return;
    cudaStreamSynchronize(stream);
    for (size_t i = 0; i < ev_used; ++i) {
      float ms = 0;
      if (cudaEventElapsedTime(&ms, ev_pool[i].a, ev_pool[i].b) == cudaSuccess) stage_ms[ev_pool[i].stage] += ms;
    }
    ev_used = 0;
  }

  int init(const gpba_problem* P, int dev, bool async_upload);
  int build_structure();
  int count_hpl();
  int build_cholesky_structure();
  int capture_cholesky_graph();
  void fill_view();
  int compute_records(int buf, bool full);
  int compute_errors(int buf, bool store, double* chi2, bool trial = false, const volatile unsigned char* stop = nullptr);
  int build_system();
  int solve(double lambda);
  int read_fail(bool* ok);
  int apply_update(double lambda, double* scale);
  // what the last trial evaluation (compute_errors with trial = true) brought back in its one host read
  double trial_scale = 0;       // computeScale: sum over all ranks of x (lambda x + b)
  bool trial_failed = false;    // any rank saw a non-positive pivot (landmark block or reduced system)
  bool stop_seen = false;       // the force-stop flag as ALL ranks agreed to see it
  bool chi_cache_valid = false; // chi_cache is the robust chi2 of state buffer `cur` (the last trial was accepted)
  double chi_cache = 0;
  int lm_solve(int iteration, const gpba_lm_params& P, const volatile unsigned char* stop, gpba_lm_trace* tr, int* result);
  int optimize(int iters, const volatile unsigned char* stop, const gpba_lm_params& P, gpba_lm_trace* tr);
  int scatter_points(int buf);
  DevView Vb(int buf) { DevView v = V; v.cam = d_cam[buf].p; return v; }   // the view of state buffer `buf`
  int refresh_cams();
  bool any_ext_free() const { for (uint8_t f : ext_free) if (f) return true; return false; }
  int download_state(double* kf_pose, double* kf_vel, double* pt_xyz);
  int allreduce_system();
  int allreduce_scalar(double* v, int count = 1, int op = kNcclSum);
};

static double f32sq(double d) { return (double)(float)(d * d); }  // RobustKernelHuber::setDelta: float dsqr

int Solver::init(const gpba_problem* P, int dev, bool async_upload) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { g_err = "no CUDA device (libgpba has no CPU fallback)"; return GPBA_ERR_NO_DEVICE; }
  if (dev < 0) CK(cudaGetDevice(&dev));
  device = dev;
  CK(cudaSetDevice(device));
  const bool verbose_init = getenv("GPBA_VERBOSE") != nullptr;
  auto ti0 = std::chrono::steady_clock::now();
  auto ilap = [&](const char* what) {
    if (!verbose_init) return;
    auto t = std::chrono::steady_clock::now();
    fprintf(stderr, "[gpba] create %-32s %6.2f ms\n", what, std::chrono::duration<double, std::milli>(t - ti0).count());
    ti0 = t;
  };
  CK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
  stream_holder.s = stream;
  g_alloc_stream = stream;
  {
    cudaMemPool_t pool;
    CK(cudaDeviceGetDefaultMemPool(&pool, device));
    unsigned long long keep = ~0ull;   // never hand memory back to the driver between optimize() calls
    CK(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
  }
  CK(cudaEventCreate(&ev0));
  CK(cudaEventCreate(&ev1));
  h_scal = (double*)g_pinned.take();
  if (!h_scal) { g_err = "out of pinned scratch slots (more than 256 live handles?)"; return GPBA_ERR_CUDA; }
  h_fail = (int*)(h_scal + 16);
  n_cam = P->n_cam; n_kf = P->n_kf; n_pt = P->n_pt; n_rec = P->n_rec; n_obs = P->n_obs;
  n_prior = P->n_prior; n_velp = P->n_velp;
  if (n_cam <= 0 || n_kf <= 0 || n_pt < 0 || n_obs < 0 || n_rec < 0 || n_prior < 0 || n_velp < 0) { g_err = "empty problem / negative count"; return GPBA_ERR_INVALID; }
  // nothing malformed crosses the boundary: every array a count promises must be there, every index in range
  if (!P->cam_intr || !P->cam_Tbc || !P->kf_pose || !P->kf_vel || !P->kf_time || !P->kf_fixed || (n_pt > 0 && !P->pt_xyz) ||
      (n_rec > 0 && (!P->rec_kf1 || !P->rec_kf2 || !P->rec_cam || !P->rec_t)) ||
      (n_obs > 0 && (!P->obs_u || !P->obs_v || !P->obs_inv_sigma2 || !P->obs_rec || !P->obs_pt)) ||
      (n_prior > 0 && (!P->prior_kf1 || !P->prior_kf2)) || (n_velp > 0 && !P->velp_kf)) { g_err = "null array with a non-zero count"; return GPBA_ERR_INVALID; }
  if (n_obs > 0 && (n_rec == 0 || n_pt == 0)) { g_err = "observations without records / points"; return GPBA_ERR_INVALID; }
  for (int i = 0; i < n_prior; ++i)
    if (P->prior_kf1[i] < 0 || P->prior_kf1[i] >= n_kf || P->prior_kf2[i] < 0 || P->prior_kf2[i] >= n_kf) { g_err = "prior keyframe index out of range"; return GPBA_ERR_INVALID; }
  for (int i = 0; i < n_velp; ++i)
    if (P->velp_kf[i] < 0 || P->velp_kf[i] >= n_kf) { g_err = "velocity-prior keyframe index out of range"; return GPBA_ERR_INVALID; }
  bf = P->bf;
  h_pose.assign(P->kf_pose, P->kf_pose + 7 * (size_t)n_kf);
  h_vel.assign(P->kf_vel, P->kf_vel + 6 * (size_t)n_kf);
  h_time.assign(P->kf_time, P->kf_time + n_kf);
  kf_fixed.assign(P->kf_fixed, P->kf_fixed + n_kf);
  h_pt.assign(P->pt_xyz, P->pt_xyz + 3 * (size_t)n_pt);
  rec_kf1.assign(P->rec_kf1, P->rec_kf1 + n_rec); rec_kf2.assign(P->rec_kf2, P->rec_kf2 + n_rec);
  rec_cam.assign(P->rec_cam, P->rec_cam + n_rec); rec_t.assign(P->rec_t, P->rec_t + n_rec);
  stereo = false;
  if (P->obs_ur)
    for (int64_t i = 0; i < n_obs; ++i) if (P->obs_ur[i] >= 0) { stereo = true; break; }
  if (P->obs_flags) obs_flags.assign(P->obs_flags, P->obs_flags + n_obs); else obs_flags.assign(n_obs, 0);
  prior_kf1.assign(P->prior_kf1, P->prior_kf1 + n_prior); prior_kf2.assign(P->prior_kf2, P->prior_kf2 + n_prior);
  velp_kf.assign(P->velp_kf, P->velp_kf + n_velp);
  for (int i = 0; i < 6; ++i) qc[i] = P->qc[i];
  huber_mono = P->huber_mono; huber_stereo = P->huber_stereo; huber_prior = P->huber_prior;
  lambda_init = P->lambda_init; linear_solver = P->linear_solver;
  for (int r = 0; r < n_rec; ++r)
    if (rec_kf2[r] < 0 || rec_kf2[r] >= n_kf || rec_kf1[r] < -1 || rec_kf1[r] >= n_kf || rec_cam[r] < 0 || rec_cam[r] >= n_cam) { g_err = "record index out of range"; return GPBA_ERR_INVALID; }

  // per-camera constants
  std::vector<CamConst> cams(n_cam);
  for (int c = 0; c < n_cam; ++c) {
    CamConst& cc = cams[c];
    cc.fx = P->cam_intr[4 * c]; cc.fy = P->cam_intr[4 * c + 1]; cc.cx = P->cam_intr[4 * c + 2]; cc.cy = P->cam_intr[4 * c + 3];
    SE3 Tbc = load_se3(P->cam_Tbc + 7 * c);
    SE3 Tcb = se3_inv(Tbc);
    M3 Rcb = quat_to_R(Tcb.q), Rbc = quat_to_R(Tbc.q);
    for (int i = 0; i < 9; ++i) { cc.Rcb[i] = Rcb.a[i]; cc.Rbc[i] = Rbc.a[i]; }
    for (int i = 0; i < 3; ++i) { cc.tcb[i] = Tcb.t[i]; cc.tbc[i] = Tbc.t[i]; }
    cc.qbc[0] = Tbc.q.x; cc.qbc[1] = Tbc.q.y; cc.qbc[2] = Tbc.q.z; cc.qbc[3] = Tbc.q.w;
    const M6 A = se3_Adj(Tbc);
    for (int r = 0; r < 6; ++r) for (int k = 0; k < 6; ++k) cc.AdjTbc[r * 6 + k] = A(r, k);
  }
  h_cam0 = cams;
  h_ext0.assign(P->cam_Tbc, P->cam_Tbc + 7 * (size_t)n_cam);
  ext_free.assign(n_cam, 0); ext_prior_on.assign(n_cam, 0); ext_h.assign(n_cam, -1);
  ext_prior_qinv.assign(4 * (size_t)n_cam, 0.0); ext_prior_info.assign(9 * (size_t)n_cam, 0.0);
  for (int b = 0; b < 2; ++b) { CKR(d_cam[b].upload(cams, stream)); CKR(d_ext[b].upload(h_ext0, stream)); }
  CKR(d_ext_h.upload(ext_h, stream)); CKR(d_ext_prior_on.upload(ext_prior_on, stream));
  CKR(d_ext_prior_qinv.upload(ext_prior_qinv, stream)); CKR(d_ext_prior_info.upload(ext_prior_info, stream));
  CKR(d_time.upload(h_time, stream));
  CKR(d_rec_kf1.upload(rec_kf1, stream)); CKR(d_rec_kf2.upload(rec_kf2, stream));
  CKR(d_rec_cam.upload(rec_cam, stream)); CKR(d_rec_t.upload(rec_t, stream));
  CKR(d_prior_kf1.upload(prior_kf1, stream)); CKR(d_prior_kf2.upload(prior_kf2, stream));
  CKR(d_velp_kf.upload(velp_kf, stream));
  CKR(d_pt_full.upload(h_pt, stream));
  for (int b = 0; b < 2; ++b) { CKR(d_pose[b].upload(h_pose, stream)); CKR(d_vel[b].upload(h_vel, stream)); }
  CKR(d_chi2.alloc((size_t)n_obs));
  CK(cudaMemsetAsync(d_chi2.p, 0, sizeof(double) * (size_t)std::max<int64_t>(n_obs, 1), stream));
  CKR(d_all_rec.upload(P->obs_rec, (size_t)n_obs, stream)); CKR(d_all_pt.upload(P->obs_pt, (size_t)n_obs, stream));
  CKR(d_scal.alloc(8)); CKR(d_fail.alloc(2));  // [0] failure flag, [1] work counter of K4b
  if (n_obs > 0) {   // index validation on the device (the host never walks the observation arrays)
    CK(cudaMemsetAsync(d_fail.p, 0, sizeof(int), stream));
    k_check_indices<<<(int)std::min<int64_t>((n_obs + 255) / 256, 148 * 8), 256, 0, stream>>>(n_obs, d_all_rec.p, d_all_pt.p, n_rec, n_pt, d_fail.p);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(h_fail, d_fail.p, sizeof(int), cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    if (*h_fail) { g_err = "observation index out of range"; return GPBA_ERR_INVALID; }
  }
  ilap("host copies + index upload + check");
  if (async_upload && n_obs > 0) {
    // allocate in stream order on the main stream, copy on the second stream: build_structure works on the index
    // arrays meanwhile and waits for ev_meas right before it needs the measurements
    CKR(d_all_u.alloc((size_t)n_obs)); CKR(d_all_v.alloc((size_t)n_obs)); CKR(d_all_w.alloc((size_t)n_obs));
    if (stereo) CKR(d_all_ur.alloc((size_t)n_obs));
    CK(cudaStreamCreateWithFlags(&copy_holder.s, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&ev_meas, cudaEventDisableTiming));
    CK(cudaEventRecord(ev_meas, stream));
    CK(cudaStreamWaitEvent(copy_holder.s, ev_meas, 0));
    const size_t nb = sizeof(double) * (size_t)n_obs;
    CK(cudaMemcpyAsync(d_all_u.p, P->obs_u, nb, cudaMemcpyHostToDevice, copy_holder.s));
    CK(cudaMemcpyAsync(d_all_v.p, P->obs_v, nb, cudaMemcpyHostToDevice, copy_holder.s));
    CK(cudaMemcpyAsync(d_all_w.p, P->obs_inv_sigma2, nb, cudaMemcpyHostToDevice, copy_holder.s));
    if (stereo) CK(cudaMemcpyAsync(d_all_ur.p, P->obs_ur, nb, cudaMemcpyHostToDevice, copy_holder.s));
    CK(cudaEventRecord(ev_meas, copy_holder.s));
    meas_pending = true;
  } else {
    CKR(d_all_u.upload(P->obs_u, (size_t)n_obs, stream)); CKR(d_all_v.upload(P->obs_v, (size_t)n_obs, stream));
    CKR(d_all_w.upload(P->obs_inv_sigma2, (size_t)n_obs, stream));
    if (stereo) CKR(d_all_ur.upload(P->obs_ur, (size_t)n_obs, stream));
  }
  CKR(d_rec.alloc((size_t)n_rec * GPBA_REC_STRIDE)); CKR(d_rec_lite.alloc((size_t)n_rec * GPBA_REC_LITE_STRIDE));
  CKR(d_recS.alloc((size_t)n_rec * 27)); CKR(d_Y.alloc((size_t)n_rec * 6));
  CKR(d_prior_rho.alloc((size_t)n_prior + n_velp + n_cam));
  ilap("measurement upload enqueued");
  CK(cudaStreamSynchronize(stream));
  ilap("final sync");
  return GPBA_OK;
}

void Solver::fill_view() {
  V.n_cam = n_cam; V.n_kf = n_kf; V.n_pt = n_pt; V.n_rec = n_rec; V.n_prior = n_prior; V.n_velp = n_velp;
  V.cam = d_cam[cur].p; V.kf_time = d_time.p; V.kf_h = d_kf_h.p;
  V.n_pose_kf = n_pose_kf; V.ext_h = d_ext_h.p; V.ext_prior_on = d_ext_prior_on.p;
  V.ext_prior_qinv = d_ext_prior_qinv.p; V.ext_prior_info = d_ext_prior_info.p;
  V.rec_hpp13 = d_rec_hpp13.p; V.rec_hpp23 = d_rec_hpp23.p; V.rec_hpp33 = d_rec_hpp33.p;
  V.rec_kf1 = d_rec_kf1.p; V.rec_kf2 = d_rec_kf2.p; V.rec_cam = d_rec_cam.p; V.rec_t = d_rec_t.p;
  V.prior_kf1 = d_prior_kf1.p; V.prior_kf2 = d_prior_kf2.p; V.velp_kf = d_velp_kf.p;
  for (int i = 0; i < 6; ++i) V.qc_inv[i] = 1.0 / qc[i];  // GaussianProcess::mQcInv of a diagonal Qc
  V.bf = bf;
  V.hub_mono_delta = huber_mono; V.hub_mono_dsqr = f32sq(huber_mono);
  V.hub_stereo_delta = huber_stereo; V.hub_stereo_dsqr = f32sq(huber_stereo);
  V.hub_prior_delta = huber_prior; V.hub_prior_dsqr = f32sq(huber_prior);
  V.n_aobs = n_aobs;
  V.o_u = d_o_u.p; V.o_v = d_o_v.p; V.o_ur = d_o_ur.p; V.o_w = d_o_w.p; V.o_rec = d_o_rec.p; V.o_lm = d_o_lm.p;
  V.o_flags = d_o_flags.p; V.o_orig = d_o_orig.p;
  V.n_lm = n_lm; V.lm_pt = d_lm_pt.p; V.lm_obs_begin = d_lm_obs_begin.p;
  V.n_tiles = n_tiles; V.tile_lm = d_tile_lm.p; V.tile_rlo = d_tile_rlo.p; V.tile_rcnt = d_tile_rcnt.p;
  V.rperm = d_rperm.p; V.n_rseg = n_rseg; V.rseg_rec = d_rseg_rec.p; V.rseg_begin = d_rseg_begin.p;
  V.n_pose = n_pose; V.n_hpp = n_hpp; V.n_hs = n_hs;
  V.rec_hpp11 = d_rec_hpp11.p; V.rec_hpp12 = d_rec_hpp12.p; V.rec_hpp22 = d_rec_hpp22.p;
  V.prior_hpp11 = d_prior_hpp11.p; V.prior_hpp12 = d_prior_hpp12.p; V.prior_hpp22 = d_prior_hpp22.p;
  V.pose_hpp_diag = d_pose_hpp_diag.p; V.hs_from_hpp = d_hs_from_hpp.p; V.hs_diag_pose = d_hs_diag_pose.p;
}

// device helpers for point gather / scatter between the full (original order) and sorted-landmark arrays
__global__ void k_gather_pts(int n_lm, const int* __restrict__ lm_pt, const double* __restrict__ full, double* __restrict__ s0, double* __restrict__ s1) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_lm * 3) return;
  const double v = full[3 * (size_t)lm_pt[i / 3] + i % 3];
  s0[i] = v; s1[i] = v;
}
__global__ void k_scatter_pts(int n_lm, const int* __restrict__ lm_pt, const double* __restrict__ s, double* __restrict__ full) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_lm * 3) return;
  full[3 * (size_t)lm_pt[i / 3] + i % 3] = s[i];
}

// multi-GPU: after the all-reduce (sum of zero-initialised staging buffers, every entry owned by exactly one rank)
__global__ void k_merge_pts(int n_pt, const int* __restrict__ pt_act, const double* __restrict__ g, double* __restrict__ full) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_pt * 3 && pt_act[i / 3]) full[i] = g[i];
}
__global__ void k_merge_chi2(int64_t n_obs, const uint8_t* __restrict__ flags, const double* __restrict__ g, double* __restrict__ chi2) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_obs; i += (int64_t)gridDim.x * blockDim.x)
    if (!(flags[i] & 0x2u)) chi2[i] = g[i];
}

// ---------------------------------------------------------------------------------------------------
// P0: SparseOptimizer::initializeOptimization + buildIndexMapping + BlockSolver::buildStructure
// (sparse_optimizer.cpp:199-267,166-190; block_solver.hpp:142-295), integer and bit-exact.  The host does the O(n_obs)
// bucket passes and the (small) pose-level patterns; the O(sum d^2) pairing runs on the device (gpba_structure.cuh).
int Solver::build_structure() {
  if (chol_future.valid()) chol_future.get();   // tables of an earlier, abandoned build
  CK(cudaSetDevice(device));
  g_alloc_stream = stream;
  const bool verbose = getenv("GPBA_VERBOSE") != nullptr;
  auto tp0 = std::chrono::steady_clock::now();
  auto lap = [&](const char* what) {
    if (!verbose) return;
    cudaStreamSynchronize(stream);
    auto now = std::chrono::steady_clock::now();
    fprintf(stderr, "[gpba] build_structure %-28s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(now - tp0).count());
    tp0 = now;
  };
  if (structure_ok && n_lm > 0) CKR(scatter_points(cur));  // keep d_pt_full current before re-sorting
  // exclusive / inclusive scans and sorts below all go through cub on `stream`
  auto excl_scan64 = [&](const int64_t* in, int64_t* out, int n) -> int {
    size_t need = 0;
    CK(cub::DeviceScan::ExclusiveSum(nullptr, need, in, out, n, stream));
    CK(cub_tmp.reserve(need, stream));
    CK(cub::DeviceScan::ExclusiveSum(cub_tmp.p, need, in, out, n, stream));
    return GPBA_OK;
  };
  const int gall = (int)std::min<int64_t>((n_obs + 255) / 256 + 1, 148 * 16);
  // --- active set (edge active iff level 0; vertex active iff it has an active edge): one device pass
  DBuf<unsigned char> d_rec_used;
  DBuf<int> d_first_kf, d_flag;
  CKR(d_rec_used.alloc((size_t)n_rec)); CKR(d_pt_act.alloc((size_t)n_pt)); CKR(d_first_kf.alloc((size_t)n_pt)); CKR(d_flag.alloc(1));
  CK(cudaMemsetAsync(d_rec_used.p, 0, (size_t)n_rec, stream));
  CK(cudaMemsetAsync(d_pt_act.p, 0, sizeof(int) * (size_t)n_pt, stream));
  CK(cudaMemsetAsync(d_first_kf.p, 0x7f, sizeof(int) * (size_t)n_pt, stream));
  CK(cudaMemsetAsync(d_flag.p, 0, sizeof(int), stream));
  CKR(d_all_flags.upload(obs_flags, stream));
  if (n_obs > 0) {
    k_scan_obs<<<gall, 256, 0, stream>>>(n_obs, d_all_flags.p, d_all_rec.p, d_all_pt.p, d_rec_kf2.p, d_rec_used.p, d_pt_act.p, d_first_kf.p, d_flag.p);
    CK(cudaGetLastError());
  }
  std::vector<unsigned char> rec_used((size_t)n_rec, 0);
  int h_any_level1 = 0;
  if (n_rec) CK(cudaMemcpyAsync(rec_used.data(), d_rec_used.p, (size_t)n_rec, cudaMemcpyDeviceToHost, stream));
  CK(cudaMemcpyAsync(&h_any_level1, d_flag.p, sizeof(int), cudaMemcpyDeviceToHost, stream));
  CK(cudaStreamSynchronize(stream));
  const bool any_level1 = h_any_level1 != 0;
  std::vector<char> kf_act(n_kf, 0);
  for (int r = 0; r < n_rec; ++r)
    if (rec_used[r]) { if (rec_kf1[r] >= 0) kf_act[rec_kf1[r]] = 1; kf_act[rec_kf2[r]] = 1; }
  for (int i = 0; i < n_prior; ++i)
    if (!(kf_fixed[prior_kf1[i]] && kf_fixed[prior_kf2[i]])) kf_act[prior_kf1[i]] = kf_act[prior_kf2[i]] = 1;
  for (int i = 0; i < n_velp; ++i)
    if (!kf_fixed[velp_kf[i]]) kf_act[velp_kf[i]] = 1;
  kf_h.assign(n_kf, -1);
  n_pose = 0;
  for (int k = 0; k < n_kf; ++k)
    if (kf_act[k] && !kf_fixed[k]) kf_h[k] = n_pose++;
  n_pose_kf = n_pose;
  // free extrinsic vertices follow the keyframes (their vertex ids are larger, Optimizer.cc:986); a vertex is active iff it
  // carries an active edge: its EdgeExtrinsicPrior (active iff the vertex is not fixed) or an active GP edge of its camera
  {
    std::vector<char> cam_act(n_cam, 0);
    for (int r = 0; r < n_rec; ++r) if (rec_used[r] && rec_kf1[r] >= 0) cam_act[rec_cam[r]] = 1;
    std::vector<unsigned char> prior_act(n_cam, 0);
    for (int c = 0; c < n_cam; ++c) {
      ext_h[c] = (ext_free[c] && (ext_prior_on[c] || cam_act[c])) ? n_pose++ : -1;
      prior_act[c] = ext_h[c] >= 0 && ext_prior_on[c];
    }
    CKR(d_ext_h.upload(ext_h, stream)); CKR(d_ext_prior_on.upload(prior_act, stream));
    CKR(d_ext_prior_qinv.upload(ext_prior_qinv, stream)); CKR(d_ext_prior_info.upload(ext_prior_info, stream));
    CK(cudaStreamSynchronize(stream));   // prior_act goes out of scope
  }
  CKR(d_kf_h.upload(kf_h, stream));
  // --- landmark order: ascending first keyframe, ties by point id (stable radix sort of the points);
  //     g2o landmark index = rank among active points in ascending id (exclusive scan of the activity flags)
  n_lm_all = 0;
  DBuf<int> d_pkey, d_pkey_s, d_pval, d_sorted_pt, d_pt_lm_all, d_rank_of_pt, d_cnt_all;
  const size_t npt1 = (size_t)std::max(n_pt, 1);
  CKR(d_pkey.alloc(npt1)); CKR(d_pkey_s.alloc(npt1)); CKR(d_pval.alloc(npt1)); CKR(d_sorted_pt.alloc(npt1)); CKR(d_pt_lm_all.alloc(npt1)); CKR(d_rank_of_pt.alloc(npt1));
  if (n_pt > 0) {
    k_point_keys<<<(n_pt + 255) / 256, 256, 0, stream>>>(n_pt, d_pt_act.p, d_first_kf.p, n_kf, d_pkey.p, d_pval.p);
    size_t need = 0;
    CK(cub::DeviceRadixSort::SortPairs(nullptr, need, d_pkey.p, d_pkey_s.p, d_pval.p, d_sorted_pt.p, n_pt, 0, bits_for((unsigned long long)n_kf + 1), stream));
    CK(cub_tmp.reserve(need, stream));
    CK(cub::DeviceRadixSort::SortPairs(cub_tmp.p, need, d_pkey.p, d_pkey_s.p, d_pval.p, d_sorted_pt.p, n_pt, 0, bits_for((unsigned long long)n_kf + 1), stream));
    need = 0;
    CK(cub::DeviceScan::ExclusiveSum(nullptr, need, d_pt_act.p, d_rank_of_pt.p, n_pt, stream));
    CK(cub_tmp.reserve(need, stream));
    CK(cub::DeviceScan::ExclusiveSum(cub_tmp.p, need, d_pt_act.p, d_rank_of_pt.p, n_pt, stream));
    int last_rank = 0, last_act = 0;
    CK(cudaMemcpyAsync(&last_rank, d_rank_of_pt.p + (n_pt - 1), sizeof(int), cudaMemcpyDeviceToHost, stream));
    CK(cudaMemcpyAsync(&last_act, d_pt_act.p + (n_pt - 1), sizeof(int), cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    n_lm_all = last_rank + last_act;
    k_point_index<<<(n_pt + 255) / 256, 256, 0, stream>>>(n_pt, n_lm_all, d_sorted_pt.p, d_pt_lm_all.p);
    CK(cudaGetLastError());
  }
  lap("active set + landmark order");
  // --- multi-GPU: this rank owns a contiguous range of the sorted landmarks, balanced by observation count;
  //     patterns (Hpp, Hschur) are built from ALL landmarks so that every rank packs the same block list
  CKR(d_cnt_all.alloc((size_t)n_lm_all + 1));
  CK(cudaMemsetAsync(d_cnt_all.p, 0, sizeof(int) * ((size_t)n_lm_all + 1), stream));
  if (n_obs > 0 && n_lm_all > 0) {
    k_count_lm_obs<<<gall, 256, 0, stream>>>(n_obs, d_all_flags.p, d_all_pt.p, d_pt_lm_all.p, 0, d_cnt_all.p);
    CK(cudaGetLastError());
  }
  int own_lo = 0, own_hi = n_lm_all;
  if (nranks > 1) {
    std::vector<int> c((size_t)n_lm_all + 1, 0);
    CK(cudaMemcpyAsync(c.data(), d_cnt_all.p, sizeof(int) * (size_t)n_lm_all, cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    std::vector<int64_t> cnt(n_lm_all + 1, 0);
    for (int l = 0; l < n_lm_all; ++l) cnt[l + 1] = cnt[l] + c[l];
    const int64_t total = cnt[n_lm_all];
    auto cut = [&](int r) { return (int)(std::lower_bound(cnt.begin(), cnt.end(), total * r / nranks) - cnt.begin()); };
    own_lo = rank == 0 ? 0 : std::min(cut(rank), n_lm_all);
    own_hi = rank == nranks - 1 ? n_lm_all : std::min(cut(rank + 1), n_lm_all);
    if (own_hi < own_lo) own_hi = own_lo;
  }
  n_lm = own_hi - own_lo;
  // --- compute list: active observations of the own landmarks, sorted by landmark (stable: insertion order inside)
  const size_t nl1 = (size_t)n_lm + 1;
  DBuf<int> d_okey, d_okey_s, d_rcnt;
  DBuf<int64_t> d_oval, d_cnt64, d_npair, d_lm_pair_begin;
  CKR(d_cnt64.alloc(nl1)); CKR(d_lm_obs_begin.alloc(nl1)); CKR(d_npair.alloc(nl1)); CKR(d_lm_pair_begin.alloc(nl1));
  CKR(d_lm_pt.alloc((size_t)std::max(n_lm, 1)));
  DBuf<int> d_lm_rank;
  CKR(d_lm_rank.alloc((size_t)std::max(n_lm, 1)));
  k_widen<<<(int)((nl1 + 255) / 256), 256, 0, stream>>>(n_lm + 1, d_cnt_all.p + own_lo, d_cnt64.p);
  if (n_lm > 0) {
    CK(cudaMemsetAsync(d_cnt64.p + n_lm, 0, sizeof(int64_t), stream));  // the entry after the range belongs to the next rank
    CK(cudaMemcpyAsync(d_lm_pt.p, d_sorted_pt.p + own_lo, sizeof(int) * (size_t)n_lm, cudaMemcpyDeviceToDevice, stream));
    k_lm_rank<<<(n_lm + 255) / 256, 256, 0, stream>>>(n_lm, d_lm_pt.p, d_rank_of_pt.p, d_lm_rank.p);
  }
  CKR(excl_scan64(d_cnt64.p, d_lm_obs_begin.p, n_lm + 1));
  k_pair_counts<<<(int)((nl1 + 255) / 256), 256, 0, stream>>>(n_lm, d_cnt_all.p + own_lo, d_npair.p);
  CKR(excl_scan64(d_npair.p, d_lm_pair_begin.p, n_lm + 1));
  CK(cudaGetLastError());
  lm_obs_begin.assign(nl1, 0);
  lm_pt.assign((size_t)n_lm, 0); lm_rank.assign((size_t)n_lm, 0);
  CK(cudaMemcpyAsync(lm_obs_begin.data(), d_lm_obs_begin.p, sizeof(int64_t) * nl1, cudaMemcpyDeviceToHost, stream));
  if (n_lm > 0) {
    CK(cudaMemcpyAsync(lm_pt.data(), d_lm_pt.p, sizeof(int) * (size_t)n_lm, cudaMemcpyDeviceToHost, stream));
    CK(cudaMemcpyAsync(lm_rank.data(), d_lm_rank.p, sizeof(int) * (size_t)n_lm, cudaMemcpyDeviceToHost, stream));
  }
  CK(cudaMemcpyAsync(&n_pairs, d_lm_pair_begin.p + n_lm, sizeof(int64_t), cudaMemcpyDeviceToHost, stream));
  CK(cudaStreamSynchronize(stream));
  n_aobs = lm_obs_begin[n_lm];
  const size_t na = (size_t)std::max<int64_t>(n_aobs, 1);
  CKR(d_o_orig.alloc(na)); CKR(d_o_lm.alloc(na));
  if (n_obs > 0) {
    const size_t no = (size_t)n_obs;
    CKR(d_okey.alloc(no)); CKR(d_okey_s.alloc(no)); CKR(d_oval.alloc(no));
    DBuf<int64_t> d_oval_s;
    CKR(d_oval_s.alloc(no));
    k_obs_keys<<<gall, 256, 0, stream>>>(n_obs, d_all_flags.p, d_all_pt.p, d_pt_lm_all.p, own_lo, n_lm, 0, d_okey.p, d_oval.p);
    size_t need = 0;
    CK(cub::DeviceRadixSort::SortPairs(nullptr, need, d_okey.p, d_okey_s.p, d_oval.p, d_oval_s.p, n_obs, 0, bits_for((unsigned long long)n_lm), stream));
    CK(cub_tmp.reserve(need, stream));
    CK(cub::DeviceRadixSort::SortPairs(cub_tmp.p, need, d_okey.p, d_okey_s.p, d_oval.p, d_oval_s.p, n_obs, 0, bits_for((unsigned long long)n_lm), stream));
    if (n_aobs > 0) {
      CK(cudaMemcpyAsync(d_o_orig.p, d_oval_s.p, sizeof(int64_t) * (size_t)n_aobs, cudaMemcpyDeviceToDevice, stream));
      CK(cudaMemcpyAsync(d_o_lm.p, d_okey_s.p, sizeof(int) * (size_t)n_aobs, cudaMemcpyDeviceToDevice, stream));
    }
    CK(cudaStreamSynchronize(stream));   // the temporaries above go back to the pool in stream order anyway
  }
  lap("observation sort");
  // --- device: sorted observation arrays
  CKR(d_o_u.alloc(na)); CKR(d_o_v.alloc(na)); CKR(d_o_w.alloc(na)); CKR(d_o_rec.alloc(na)); CKR(d_o_flags.alloc(na));
  if (stereo) CKR(d_o_ur.alloc(na));
  const int gobs = (int)std::min<int64_t>((n_aobs + 255) / 256 + 1, 148 * 16);
  std::vector<int64_t> rcount(n_rec + 1, 0);
  if (n_aobs > 0) {
    k_gather_obs_idx<<<gobs, 256, 0, stream>>>(n_aobs, d_o_orig.p, d_all_rec.p, d_all_flags.p, d_o_rec.p, d_o_flags.p);
    CKR(d_rcnt.alloc((size_t)n_rec));
    CK(cudaMemsetAsync(d_rcnt.p, 0, sizeof(int) * (size_t)n_rec, stream));
    k_hist_int<<<gobs, 256, 0, stream>>>(n_aobs, d_o_rec.p, d_rcnt.p);
    CK(cudaGetLastError());
    std::vector<int> rc((size_t)n_rec);
    CK(cudaMemcpyAsync(rc.data(), d_rcnt.p, sizeof(int) * (size_t)n_rec, cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    for (int r = 0; r < n_rec; ++r) rcount[r + 1] = rcount[r] + rc[r];
  }
  // --- landmark-aligned observation tiles (<= GPBA_TILE_OBS observations, a tile ends where a landmark ends) and the
  //     window of record rows each of them touches
  {
    std::vector<int> tile_lm;
    tile_lm.push_back(0);
    int64_t start = 0;
    for (int l = 0; l < n_lm; ++l)
      if (lm_obs_begin[l + 1] - start > GPBA_TILE_OBS && l > tile_lm.back()) { tile_lm.push_back(l); start = lm_obs_begin[l]; }
    if (n_lm > 0) tile_lm.push_back(n_lm);
    n_tiles = (int)tile_lm.size() - 1;
    CKR(d_tile_lm.upload(tile_lm, stream));
    CKR(d_tile_rlo.alloc((size_t)std::max(n_tiles, 1))); CKR(d_tile_rcnt.alloc((size_t)std::max(n_tiles, 1)));
    if (n_tiles > 0) {
      k_tile_windows<<<(n_tiles * 32 + 255) / 256, 256, 0, stream>>>(n_tiles, d_tile_lm.p, d_lm_obs_begin.p, d_o_rec.p, n_rec, d_tile_rlo.p, d_tile_rcnt.p);
      CK(cudaGetLastError());
    }
    CK(cudaStreamSynchronize(stream));   // tile_lm goes out of scope
  }
  lap("gather");
  // --- device: observation pairs grouped by record pair
  std::vector<unsigned long long> rp_key;   // unique record pairs of the compute list, ascending
  const unsigned long long nrec2 = (unsigned long long)n_rec * (unsigned long long)n_rec;
  static const int LM_CHUNK = getenv("GPBA_LM_CHUNK") ? std::max(256, atoi(getenv("GPBA_LM_CHUNK"))) : 16384;   // landmarks per chunk: ~31 MB of U rows at 10 observations per landmark (measured at C4: 2048 -> 8.2, 4096 -> 7.7, 8192 -> 7.4, 16384 -> 7.3 ms of K4b per optimize)
  // One pass = emit + sort + run-length encode.  with_items: runs of (chunk, record pair) become the work items of K4b
  // and the sorted observation pairs are kept; otherwise only the unique record pairs are wanted (pattern).
  auto pair_pass = [&](int nl, const int64_t* d_lob, const int64_t* d_lpb, int64_t np, const int* d_rec_sorted, bool with_items,
                       std::vector<unsigned long long>& keys_out) -> int {
    keys_out.clear();
    if (with_items) { n_items = 0; n_rp = 0; }
    if (np == 0) return GPBA_OK;
    const int chunk = with_items ? LM_CHUNK : 0;
    const unsigned long long n_chunks = chunk ? (unsigned long long)((nl + chunk - 1) / chunk) : 1ull;
    DBuf<unsigned long long> k0, k1, v0, v1, uq;
    DBuf<int> cnt, runs, dup;
    CKR(k0.alloc((size_t)np)); CKR(k1.alloc((size_t)np)); CKR(v0.alloc((size_t)np)); CKR(v1.alloc((size_t)np));
    CKR(uq.alloc((size_t)np)); CKR(cnt.alloc((size_t)np)); CKR(runs.alloc(1)); CKR(dup.alloc(1));
    CK(cudaMemsetAsync(dup.p, 0, sizeof(int), stream));
    k_emit_pairs<<<std::min((nl + 7) / 8, 148 * 16), 256, 0, stream>>>(nl, d_lob, d_lpb, d_rec_sorted, (unsigned long long)n_rec, chunk, k0.p, v0.p, dup.p);
    CK(cudaGetLastError());
    int h_runs = 0;
    unsigned long long *pk = k0.p, *pv = v0.p, *pka = k1.p, *pva = v1.p;
    CK(sort_and_encode(cub_tmp, pk, pv, pka, pva, np, bits_for(n_chunks * nrec2), uq.p, cnt.p, runs.p, &h_runs, stream));
    if (!with_items) {
      keys_out.resize(h_runs);
      CK(cudaMemcpyAsync(keys_out.data(), uq.p, sizeof(unsigned long long) * h_runs, cudaMemcpyDeviceToHost, stream));
      CK(cudaStreamSynchronize(stream));
      return GPBA_OK;
    }
    int h_dup = 0;
    CK(cudaMemcpyAsync(&h_dup, dup.p, sizeof(int), cudaMemcpyDeviceToHost, stream));
    CKR(d_pairs.alloc((size_t)np));
    CK(cudaMemcpyAsync(d_pairs.p, pv, sizeof(unsigned long long) * (size_t)np, cudaMemcpyDeviceToDevice, stream));
    // runs -> work items (long runs are cut); slot of a run = rank of its record pair among the unique record pairs
    const int ni = h_runs, gi = (ni + 255) / 256;
    DBuf<unsigned long long> low, low_s, rpk;
    DBuf<int> idx, idx_s, head, slot_p1, run_rp, nsub, sub_begin;
    DBuf<unsigned char> run_flags;
    DBuf<int64_t> beg;
    CKR(low.alloc(ni)); CKR(low_s.alloc(ni)); CKR(rpk.alloc(ni)); CKR(idx.alloc(ni)); CKR(idx_s.alloc(ni)); CKR(head.alloc(ni)); CKR(slot_p1.alloc(ni)); CKR(beg.alloc(ni));
    CKR(run_rp.alloc(ni)); CKR(run_flags.alloc(ni)); CKR(nsub.alloc((size_t)ni + 1)); CKR(sub_begin.alloc((size_t)ni + 1));
    k_item_low<<<gi, 256, 0, stream>>>(ni, uq.p, nrec2, low.p, idx.p);
    size_t need = 0;
    CK(cub::DeviceRadixSort::SortPairs(nullptr, need, low.p, low_s.p, idx.p, idx_s.p, ni, 0, bits_for(nrec2), stream));
    CK(cub_tmp.reserve(need, stream));
    CK(cub::DeviceRadixSort::SortPairs(cub_tmp.p, need, low.p, low_s.p, idx.p, idx_s.p, ni, 0, bits_for(nrec2), stream));
    k_slot_heads<<<gi, 256, 0, stream>>>(ni, low_s.p, head.p);
    need = 0;
    CK(cub::DeviceScan::InclusiveSum(nullptr, need, head.p, slot_p1.p, ni, stream));
    CK(cub_tmp.reserve(need, stream));
    CK(cub::DeviceScan::InclusiveSum(cub_tmp.p, need, head.p, slot_p1.p, ni, stream));
    k_slot_assign<<<gi, 256, 0, stream>>>(ni, low_s.p, idx_s.p, head.p, slot_p1.p, (unsigned long long)n_rec, run_rp.p, run_flags.p, rpk.p);
    need = 0;
    CK(cub::DeviceScan::ExclusiveSum(nullptr, need, cnt.p, beg.p, ni, stream));
    CK(cub_tmp.reserve(need, stream));
    CK(cub::DeviceScan::ExclusiveSum(cub_tmp.p, need, cnt.p, beg.p, ni, stream));
    CK(cudaMemsetAsync(nsub.p + ni, 0, sizeof(int), stream));
    k_run_items<<<gi, 256, 0, stream>>>(ni, cnt.p, nsub.p);
    need = 0;
    CK(cub::DeviceScan::ExclusiveSum(nullptr, need, nsub.p, sub_begin.p, ni + 1, stream));
    CK(cub_tmp.reserve(need, stream));
    CK(cub::DeviceScan::ExclusiveSum(cub_tmp.p, need, nsub.p, sub_begin.p, ni + 1, stream));
    int h_nrp = 0, h_items = 0;
    CK(cudaMemcpyAsync(&h_nrp, slot_p1.p + (ni - 1), sizeof(int), cudaMemcpyDeviceToHost, stream));
    CK(cudaMemcpyAsync(&h_items, sub_begin.p + ni, sizeof(int), cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    CKR(d_item_rp.alloc(h_items)); CKR(d_item_flags.alloc(h_items)); CKR(d_item_begin.alloc(h_items)); CKR(d_item_end.alloc(h_items));
    k_item_ranges<<<gi, 256, 0, stream>>>(ni, beg.p, cnt.p, sub_begin.p, run_rp.p, run_flags.p, d_item_begin.p, d_item_end.p, d_item_rp.p, d_item_flags.p);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(stream));
    if (h_dup) { g_err = "two observations of one landmark share a (keyframe pair, camera) record"; return GPBA_ERR_INVALID; }
    keys_out.resize(h_nrp);
    CKR(d_rp_key.alloc((size_t)h_nrp));
    CK(cudaMemcpyAsync(d_rp_key.p, rpk.p, sizeof(unsigned long long) * h_nrp, cudaMemcpyDeviceToDevice, stream));
    CK(cudaMemcpyAsync(keys_out.data(), rpk.p, sizeof(unsigned long long) * h_nrp, cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    n_items = h_items; n_rp = h_nrp;
    return GPBA_OK;
  };
  CKR(pair_pass(n_lm, d_lm_obs_begin.p, d_lm_pair_begin.p, n_pairs, d_o_rec.p, true, rp_key));
  // pattern keys: all edges (any level) of all active landmarks (block_solver.hpp:262-288) -- a second, key-only pass
  // when that set differs from the compute list (level-1 edges, or landmarks owned by other ranks)
  std::vector<unsigned long long> pat_key_store;
  const std::vector<unsigned long long>* pat_key = &rp_key;
  if ((any_level1 || nranks > 1) && n_lm_all > 0) {
    const size_t nla1 = (size_t)n_lm_all + 1, no = (size_t)n_obs;
    DBuf<int> cnt_any, key, key_s, recs;
    DBuf<int64_t> cnt64, lob, npair, lpb, val, val_s;
    CKR(cnt_any.alloc(nla1)); CKR(cnt64.alloc(nla1)); CKR(lob.alloc(nla1)); CKR(npair.alloc(nla1)); CKR(lpb.alloc(nla1));
    CKR(key.alloc(no)); CKR(key_s.alloc(no)); CKR(val.alloc(no)); CKR(val_s.alloc(no)); CKR(recs.alloc(no));
    CK(cudaMemsetAsync(cnt_any.p, 0, sizeof(int) * nla1, stream));
    k_count_lm_obs<<<gall, 256, 0, stream>>>(n_obs, d_all_flags.p, d_all_pt.p, d_pt_lm_all.p, 1, cnt_any.p);
    k_widen<<<(int)((nla1 + 255) / 256), 256, 0, stream>>>(n_lm_all + 1, cnt_any.p, cnt64.p);
    CKR(excl_scan64(cnt64.p, lob.p, n_lm_all + 1));
    k_pair_counts<<<(int)((nla1 + 255) / 256), 256, 0, stream>>>(n_lm_all, cnt_any.p, npair.p);
    CKR(excl_scan64(npair.p, lpb.p, n_lm_all + 1));
    k_obs_keys<<<gall, 256, 0, stream>>>(n_obs, d_all_flags.p, d_all_pt.p, d_pt_lm_all.p, 0, n_lm_all, 1, key.p, val.p);
    size_t need = 0;
    CK(cub::DeviceRadixSort::SortPairs(nullptr, need, key.p, key_s.p, val.p, val_s.p, n_obs, 0, bits_for((unsigned long long)n_lm_all), stream));
    CK(cub_tmp.reserve(need, stream));
    CK(cub::DeviceRadixSort::SortPairs(cub_tmp.p, need, key.p, key_s.p, val.p, val_s.p, n_obs, 0, bits_for((unsigned long long)n_lm_all), stream));
    int64_t n_any = 0, np_any = 0;
    CK(cudaMemcpyAsync(&n_any, lob.p + n_lm_all, sizeof(int64_t), cudaMemcpyDeviceToHost, stream));
    CK(cudaMemcpyAsync(&np_any, lpb.p + n_lm_all, sizeof(int64_t), cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    if (n_any > 0) { k_gather_int<<<gall, 256, 0, stream>>>(n_any, val_s.p, d_all_rec.p, recs.p); CK(cudaGetLastError()); }
    CKR(pair_pass(n_lm_all, lob.p, lpb.p, np_any, recs.p, false, pat_key_store));
    pat_key = &pat_key_store;
  }
  lap("pair sort");
  // --- #Hpl blocks: reported only (gpba_structure_info), counted on demand by count_hpl(): 0.6 ms and a host synchronisation
  // that a gpba_optimize on a fresh handle does not need
  fill_view();
  n_hpl = -1;
  // --- Hpp pattern (upper, host: a few thousand blocks): diagonals + priors + the keyframe pair of every used record
  std::vector<std::vector<int>> pp_rows(n_pose);
  auto add_pair = [](std::vector<std::vector<int>>& rows, int a, int b) {
    if (a < 0 || b < 0) return;
    if (a > b) std::swap(a, b);
    if (rows[a].empty() || rows[a].back() != b) rows[a].push_back(b);
  };
  for (int i = 0; i < n_pose; ++i) pp_rows[i].push_back(i);
  for (int i = 0; i < n_prior; ++i) add_pair(pp_rows, kf_h[prior_kf1[i]], kf_h[prior_kf2[i]]);
  for (int r = 0; r < n_rec; ++r)
    if (rec_used[r] && rec_kf1[r] >= 0) {
      add_pair(pp_rows, kf_h[rec_kf1[r]], kf_h[rec_kf2[r]]);
      const int h3 = ext_h[rec_cam[r]];   // EdgeMonoGPExtrinsic also links both keyframes to the extrinsic
      if (h3 >= 0) { add_pair(pp_rows, kf_h[rec_kf1[r]], h3); add_pair(pp_rows, kf_h[rec_kf2[r]], h3); }
    }
  for (auto& v : pp_rows) { std::sort(v.begin(), v.end()); v.erase(std::unique(v.begin(), v.end()), v.end()); }
  std::vector<unsigned long long> hpp_key;   // (col << 32 | row), sorted: the order of SparseBlockMatrix columns
  for (int r = 0; r < n_pose; ++r) for (int c : pp_rows[r]) hpp_key.push_back(((unsigned long long)c << 32) | (unsigned)r);
  std::sort(hpp_key.begin(), hpp_key.end());
  n_hpp = (int)hpp_key.size();
  hpp_row.resize(n_hpp); hpp_col.resize(n_hpp);
  for (int k = 0; k < n_hpp; ++k) { hpp_row[k] = (int)(unsigned)hpp_key[k]; hpp_col[k] = (int)(hpp_key[k] >> 32); }
  auto pp = [&](int a, int b) { return (int)(std::lower_bound(hpp_key.begin(), hpp_key.end(), ((unsigned long long)b << 32) | (unsigned)a) - hpp_key.begin()); };
  std::vector<int> rec11(n_rec, -1), rec12(n_rec, -1), rec22(n_rec, -1), pr11(n_prior, -1), pr12(n_prior, -1), pr22(n_prior, -1);
  auto pair_blocks = [&](int h1, int h2, int& b11, int& b12, int& b22) {
    if (h1 >= 0) b11 = pp(h1, h1);
    if (h2 >= 0) b22 = pp(h2, h2);
    if (h1 >= 0 && h2 >= 0) b12 = h1 <= h2 ? pp(h1, h2) : (pp(h2, h1) | 0x40000000);
  };
  std::vector<int> rec13(n_rec, -1), rec23(n_rec, -1), rec33(n_rec, -1);
  for (int r = 0; r < n_rec; ++r)
    if (rec_used[r]) {
      pair_blocks(rec_kf1[r] >= 0 ? kf_h[rec_kf1[r]] : -1, kf_h[rec_kf2[r]], rec11[r], rec12[r], rec22[r]);
      const int h3 = rec_kf1[r] >= 0 ? ext_h[rec_cam[r]] : -1;
      if (h3 >= 0) {
        rec33[r] = pp(h3, h3);
        if (kf_h[rec_kf1[r]] >= 0) rec13[r] = pp(kf_h[rec_kf1[r]], h3);
        if (kf_h[rec_kf2[r]] >= 0) rec23[r] = pp(kf_h[rec_kf2[r]], h3);
      }
    }
  for (int i = 0; i < n_prior; ++i) pair_blocks(kf_h[prior_kf1[i]], kf_h[prior_kf2[i]], pr11[i], pr12[i], pr22[i]);
  std::vector<int> pose_diag(n_pose);
  for (int i = 0; i < n_pose; ++i) pose_diag[i] = pp(i, i);
  // --- Hschur pattern (device): Hpp pattern U pose pairs of all edges of active landmarks = sorted unique block keys
  DBuf<unsigned long long> d_hs_key;
  {
    const int npat = (int)pat_key->size();
    const size_t nk = (size_t)npat * GPBA_KEYS_PER_RP + hpp_key.size();
    DBuf<unsigned long long> d_patk, kin, kout;
    DBuf<int> d_nsel;
    CKR(kin.alloc(nk)); CKR(kout.alloc(nk)); CKR(d_hs_key.alloc(nk)); CKR(d_nsel.alloc(1));
    if (npat) {
      const unsigned long long* src = d_rp_key.p;
      if (pat_key != &rp_key) { CKR(d_patk.upload(*pat_key, stream)); src = d_patk.p; }
      k_emit_block_keys<<<(npat + 255) / 256, 256, 0, stream>>>(V, npat, src, kin.p);
      CK(cudaGetLastError());
    }
    if (!hpp_key.empty()) CK(cudaMemcpyAsync(kin.p + (size_t)npat * GPBA_KEYS_PER_RP, hpp_key.data(), sizeof(unsigned long long) * hpp_key.size(), cudaMemcpyHostToDevice, stream));
    size_t need = 0;
    CK(cub::DeviceRadixSort::SortKeys(nullptr, need, kin.p, kout.p, (int)nk, 0, 64, stream));
    CK(cub_tmp.reserve(need, stream));
    CK(cub::DeviceRadixSort::SortKeys(cub_tmp.p, need, kin.p, kout.p, (int)nk, 0, 64, stream));
    need = 0;
    CK(cub::DeviceSelect::Unique(nullptr, need, kout.p, d_hs_key.p, d_nsel.p, (int)nk, stream));
    CK(cub_tmp.reserve(need, stream));
    CK(cub::DeviceSelect::Unique(cub_tmp.p, need, kout.p, d_hs_key.p, d_nsel.p, (int)nk, stream));
    int nsel = 0;
    CK(cudaMemcpyAsync(&nsel, d_nsel.p, sizeof(int), cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));   // also keeps hpp_key alive until its copy is done
    std::vector<unsigned long long> hs_key(nsel);
    if (nsel) CK(cudaMemcpyAsync(hs_key.data(), d_hs_key.p, sizeof(unsigned long long) * nsel, cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    if (nsel && hs_key.back() == ~0ull) --nsel;   // the "no block" marker sorts last
    n_hs = nsel;
    hs_row.resize(n_hs); hs_col.resize(n_hs);
    std::vector<int> hs_from(n_hs, -1), hs_diag(n_hs, -1);
    for (int k = 0; k < n_hs; ++k) { hs_row[k] = (int)(unsigned)hs_key[k]; hs_col[k] = (int)(hs_key[k] >> 32); if (hs_row[k] == hs_col[k]) hs_diag[k] = hs_row[k]; }
    // The host tables of the factorization (ordering, fill, level schedule, task list: 3.6 ms at C4) need nothing but this
    // block pattern: a helper thread builds them while this thread finishes the device-side structure.  hs_row / hs_col are
    // not touched again before build_cholesky_structure() has collected the result.
    static const bool chol_async = !getenv("GPBA_SYMBOLIC_SYNC");
    if (linear_solver != GPBA_SOLVER_PCG && chol_async) {
      const int np_ = n_pose, nh_ = n_hs;
      const int *hr = hs_row.data(), *hc = hs_col.data();
      chol_future = std::async(std::launch::async, [np_, nh_, hr, hc]() {
        std::unique_ptr<CholHost> H(new CholHost);
        H->rc = chol_host_phase(np_, nh_, hr, hc, *H);
        return H;
      });
    }
    for (int k = 0; k < n_hpp; ++k) hs_from[(int)(std::lower_bound(hs_key.begin(), hs_key.begin() + n_hs, hpp_key[k]) - hs_key.begin())] = k;
    CKR(d_hs_from_hpp.upload(hs_from, stream)); CKR(d_hs_diag_pose.upload(hs_diag, stream));
    CKR(d_hs_row.upload(hs_row, stream)); CKR(d_hs_col.upload(hs_col, stream));
    CK(cudaStreamSynchronize(stream));
  }
  lap("patterns");
  // --- K4c contribution lists (device): which record pairs feed which Hschur block, as (left record slice, right record
  //     slice, transpose flag), sorted by (block, left record slice) -- a stable radix sort keeps record-pair order inside
  n_con = 0;
  CKR(d_con_begin.alloc((size_t)n_hs + 1));
  if (n_rp > 0 && n_hs > 0) {
    const size_t cap = (size_t)n_rp * GPBA_MAX_CON_PER_RP;
    DBuf<unsigned long long> kin, kout, guq;
    DBuf<HsContrib> vin;
    DBuf<int> nvalid, gcnt, gstart, gruns;
    CKR(kin.alloc(cap)); CKR(kout.alloc(cap)); CKR(vin.alloc(cap)); CKR(d_con.alloc(cap)); CKR(nvalid.alloc(1));
    CK(cudaMemsetAsync(nvalid.p, 0, sizeof(int), stream));
    k_emit_contribs<<<(n_rp + 127) / 128, 128, 0, stream>>>(V, n_rp, d_rp_key.p, d_hs_key.p, n_hs, kin.p, vin.p, nvalid.p);
    CK(cudaGetLastError());
    const int kb = bits_for(((unsigned long long)n_hs * (unsigned long long)n_rec + (unsigned long long)n_rec) * 2ull);
    size_t need = 0;
    CK(cub::DeviceRadixSort::SortPairs(nullptr, need, kin.p, kout.p, vin.p, d_con.p, (int)cap, 0, 64, stream));
    CK(cub_tmp.reserve(need, stream));
    CK(cub::DeviceRadixSort::SortPairs(cub_tmp.p, need, kin.p, kout.p, vin.p, d_con.p, (int)cap, 0, 64, stream));
    (void)kb;
    int h_valid = 0;
    CK(cudaMemcpyAsync(&h_valid, nvalid.p, sizeof(int), cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    n_con = h_valid;
    CKR(guq.alloc((size_t)std::max(h_valid, 1))); CKR(gcnt.alloc((size_t)std::max(h_valid, 1))); CKR(gstart.alloc((size_t)std::max(h_valid, 1))); CKR(gruns.alloc(1));
    if (h_valid > 0) {
      need = 0;
      CK(cub::DeviceRunLengthEncode::Encode(nullptr, need, kout.p, guq.p, gcnt.p, gruns.p, h_valid, stream));
      CK(cub_tmp.reserve(need, stream));
      CK(cub::DeviceRunLengthEncode::Encode(cub_tmp.p, need, kout.p, guq.p, gcnt.p, gruns.p, h_valid, stream));
      int h_groups = 0;
      CK(cudaMemcpyAsync(&h_groups, gruns.p, sizeof(int), cudaMemcpyDeviceToHost, stream));
      CK(cudaStreamSynchronize(stream));
      need = 0;
      CK(cub::DeviceScan::ExclusiveSum(nullptr, need, gcnt.p, gstart.p, h_groups, stream));
      CK(cub_tmp.reserve(need, stream));
      CK(cub::DeviceScan::ExclusiveSum(cub_tmp.p, need, gcnt.p, gstart.p, h_groups, stream));
      k_mark_groups<<<(h_groups + 255) / 256, 256, 0, stream>>>(h_groups, gstart.p, gcnt.p, d_con.p);
    }
    k_con_begin<<<(n_hs + 256) / 256, 256, 0, stream>>>(n_hs, h_valid, 4ull * (unsigned long long)n_rec, kout.p, d_con_begin.p);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(stream));
  } else {
    CK(cudaMemsetAsync(d_con_begin.p, 0, sizeof(int) * ((size_t)n_hs + 1), stream));
  }
  lap("contribution lists");
  // --- record-major permutation (K2b): sorted-obs indices grouped by record (stable device sort), split into segments
  CKR(d_rperm.alloc(na));
  if (n_aobs > 0) {
    DBuf<int> kout; DBuf<int64_t> vin;
    CKR(kout.alloc(na)); CKR(vin.alloc(na));
    k_iota<<<gobs, 256, 0, stream>>>(n_aobs, vin.p);
    size_t need = 0;
    CK(cub::DeviceRadixSort::SortPairs(nullptr, need, d_o_rec.p, kout.p, vin.p, d_rperm.p, n_aobs, 0, bits_for((unsigned long long)n_rec), stream));
    CK(cub_tmp.reserve(need, stream));
    CK(cub::DeviceRadixSort::SortPairs(cub_tmp.p, need, d_o_rec.p, kout.p, vin.p, d_rperm.p, n_aobs, 0, bits_for((unsigned long long)n_rec), stream));
    CKR(d_r_u.alloc(na)); CKR(d_r_v.alloc(na)); CKR(d_r_w.alloc(na)); CKR(d_r_lm.alloc(na)); CKR(d_r_flags.alloc(na));
    if (stereo) CKR(d_r_ur.alloc(na));
    // the measurements are first needed here: everything above worked on indices and flags only
    if (meas_pending) { CK(cudaStreamWaitEvent(stream, ev_meas, 0)); }
    k_gather_obs_meas<<<gobs, 256, 0, stream>>>(n_aobs, d_o_orig.p, d_all_u.p, d_all_v.p, stereo ? d_all_ur.p : nullptr, d_all_w.p,
                                                d_o_u.p, d_o_v.p, stereo ? d_o_ur.p : nullptr, d_o_w.p);
    k_gather_recmajor<<<gobs, 256, 0, stream>>>(n_aobs, d_rperm.p, d_o_u.p, d_o_v.p, stereo ? d_o_ur.p : nullptr, d_o_w.p, d_o_lm.p, d_o_flags.p,
                                                d_r_u.p, d_r_v.p, stereo ? d_r_ur.p : nullptr, d_r_w.p, d_r_lm.p, d_r_flags.p);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(stream));
  }
  if (meas_pending) { CK(cudaStreamSynchronize(copy_holder.s)); meas_pending = false; }   // the caller's arrays are free again
  const int SEG = 512;
  std::vector<int> rseg_rec; std::vector<int64_t> rseg_begin;
  for (int r = 0; r < n_rec; ++r)
    for (int64_t b = rcount[r]; b < rcount[r + 1]; b += SEG) { rseg_rec.push_back(r); rseg_begin.push_back(b); }
  rseg_begin.push_back(n_aobs);
  n_rseg = (int)rseg_rec.size();
  lap("record-major permutation");
  // --- upload
  CKR(d_rseg_rec.upload(rseg_rec, stream)); CKR(d_rseg_begin.upload(rseg_begin, stream));
  CKR(d_rec_hpp11.upload(rec11, stream)); CKR(d_rec_hpp12.upload(rec12, stream)); CKR(d_rec_hpp22.upload(rec22, stream));
  CKR(d_rec_hpp13.upload(rec13, stream)); CKR(d_rec_hpp23.upload(rec23, stream)); CKR(d_rec_hpp33.upload(rec33, stream));
  CKR(d_prior_hpp11.upload(pr11, stream)); CKR(d_prior_hpp12.upload(pr12, stream)); CKR(d_prior_hpp22.upload(pr22, stream));
  CKR(d_pose_hpp_diag.upload(pose_diag, stream));
  // --- storage
  CKR(d_ptS[0].alloc((size_t)n_lm * 3)); CKR(d_ptS[1].alloc((size_t)n_lm * 3));
  CKR(d_hll.alloc((size_t)n_lm * 9)); CKR(d_bl.alloc((size_t)n_lm * 3)); CKR(d_ptL.alloc((size_t)n_lm * 9)); CKR(d_xl.alloc((size_t)n_lm * 3));
  CKR(d_W.alloc(na * 18)); CKR(d_U.alloc(na * GPBA_U_STRIDE)); CKR(d_C.alloc((size_t)std::max(n_rp, 1) * GPBA_RP_STRIDE));
  CKR(d_hpp.alloc((size_t)n_hpp * 144)); CKR(d_bp.alloc((size_t)n_pose * 12));
  CKR(d_hs.alloc((size_t)n_hs * 144 + (size_t)n_pose * 12 + 8)); CKR(d_x.alloc((size_t)n_pose * 12)); CKR(d_pose_scale.alloc((size_t)n_pose));
  d_bs.release();
  grid_obs = (int)std::min<int64_t>((n_aobs + 255) / 256, 148 * 8);
  if (grid_obs < 1) grid_obs = 1;
  CKR(d_partial.alloc((size_t)std::max(grid_obs, 148 * 16) + 16));
  fill_view();
  if (n_lm > 0) {
    k_gather_pts<<<(n_lm * 3 + 255) / 256, 256, 0, stream>>>(n_lm, d_lm_pt.p, d_pt_full.p, d_ptS[0].p, d_ptS[1].p);
    CK(cudaGetLastError());
  }
  // both state buffers must agree on fixed / inactive vertices
  CK(cudaMemcpyAsync(d_pose[1 - cur].p, d_pose[cur].p, sizeof(double) * 7 * (size_t)n_kf, cudaMemcpyDeviceToDevice, stream));
  CK(cudaMemcpyAsync(d_vel[1 - cur].p, d_vel[cur].p, sizeof(double) * 6 * (size_t)n_kf, cudaMemcpyDeviceToDevice, stream));
  lap("upload + alloc");
  if (linear_solver != GPBA_SOLVER_PCG) CKR(build_cholesky_structure());
  else CKR(pcg.setup(n_pose, n_hs, hs_row, hs_col, stream));
  CK(cudaStreamSynchronize(stream));
  lap("linear solver structure");
  info.n_free_kf = n_pose; info.n_active_pt = n_lm; info.n_active_obs = n_aobs; info.n_hpl = n_hpl; info.n_hpp = n_hpp; info.n_hschur = n_hs;
  structure_ok = true; system_ok = false; lambda_applied = false; structure_dirty = false;
  last_eval = cur;
  return GPBA_OK;
}

// #(free pose, landmark) blocks of Hpl = _Hpl->nonZeroBlocks() (block_solver.hpp:206-254)
int Solver::count_hpl() {
  if (n_hpl >= 0) return GPBA_OK;
  fill_view();
  DBuf<unsigned long long> d_cnt;
  CKR(d_cnt.alloc(1));
  CK(cudaMemsetAsync(d_cnt.p, 0, sizeof(unsigned long long), stream));
  if (n_lm > 0) { k_count_hpl<<<std::min((n_lm + 7) / 8, 148 * 16), 256, 0, stream>>>(V, d_o_rec.p, d_cnt.p); CK(cudaGetLastError()); }
  unsigned long long h = 0;
  CK(cudaMemcpyAsync(&h, d_cnt.p, sizeof(h), cudaMemcpyDeviceToHost, stream));
  CK(cudaStreamSynchronize(stream));
  n_hpl = (int64_t)h;
  info.n_hpl = n_hpl;
  return GPBA_OK;
}


// tile-level symbolic factorization (the analyzePattern of the sparse path): host tables (computed here, or already by the
// helper thread build_structure started), then allocation and upload
int Solver::build_cholesky_structure() {
  std::unique_ptr<CholHost> Hp;
  if (chol_future.valid()) Hp = chol_future.get();
  else { Hp.reset(new CholHost); Hp->rc = chol_host_phase(n_pose, n_hs, hs_row.data(), hs_col.data(), *Hp); }
  CholHost& H = *Hp;
  if (H.rc != GPBA_OK) { g_err = H.err; return H.rc; }
  const CholSymbolic& sym = H.sym;
  NT = sym.NT; chol_parts = sym.n_parts; chol_doubles = sym.doubles;
  chol_col_begin = sym.col_begin; chol_row_begin = sym.row_begin;
  lvl_pan_begin = H.lvl_pan_begin; lvl_upd_begin = H.lvl_upd_begin; lvl_back_begin = H.lvl_back_begin; lvl_ncols = H.lvl_ncols;
  chol_products = H.products; chol_update_ctas = H.update_ctas;
  const int n_levels = sym.n_levels;
  CKR(d_chol_pos_used.upload(sym.pos_used, stream));
  CKR(d_tile_off.upload(sym.tile_off, stream)); CKR(d_col_begin.upload(chol_col_begin, stream)); CKR(d_col_rows.upload(sym.col_rows, stream));
  CKR(d_chol_perm.upload(sym.perm, stream));
  CKR(d_tiles.alloc((size_t)chol_doubles)); CKR(d_chol_work.alloc((size_t)NT * GPBA_NB));
  CKR(d_chol_dinv.alloc((size_t)NT * GPBA_NB * GPBA_NB)); CKR(d_chol_x.alloc((size_t)NT * GPBA_NB));
  CKR(d_row_begin.upload(chol_row_begin, stream)); CKR(d_row_cols.upload(sym.row_cols, stream));
  CKR(d_pan_tab.upload(H.pan_tab, stream)); CKR(d_upd_tab.upload(H.upd_tab, stream)); CKR(d_back_tab.upload(H.back_tab, stream));
  CKR(d_klist.upload(H.klist, stream));
  CKR(d_lu_counter.alloc(2 * (size_t)std::max(n_levels, 1)));
  cf_tasks = (int)H.cf_tab.size();
  CKR(d_cf_prod.upload(H.prod, stream));
  CKR(d_cf_tab.upload(H.cf_tab, stream)); CKR(d_cf_need.upload(H.cf_need, stream)); CKR(d_cf_cnt.alloc(3 * (size_t)NT + 1)); CKR(d_cf_d8.alloc((size_t)NT * (GPBA_CF_D8_BYTES / 8)));
  if (getenv("GPBA_VERBOSE")) fprintf(stderr, "[gpba] cholesky: %d partitions, %d levels for %d tile columns, %zu panel CTAs, %zu update CTAs for %zu tile products\n", chol_parts, n_levels, NT, H.pan_tab.size(), H.upd_tab.size(), H.klist.size());
  CK(cudaStreamSynchronize(stream));  // the host tables go out of scope
  if (chol_graph) { cudaGraphExecDestroy(chol_graph); chol_graph = nullptr; }
  if (chol_back_graph) { cudaGraphExecDestroy(chol_back_graph); chol_back_graph = nullptr; }
  return GPBA_OK;
}

// The factorization of one structure is a fixed launch sequence (2 kernels per tile column); replaying it as a
// CUDA graph removes the per-launch host cost from the critical path of every LM trial.
// launch with the programmatic-stream-serialization attribute (PDL edge when captured into a graph)
template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl_smem(void (*kernel)(KArgs...), int grid, int block, size_t smem, cudaStream_t s, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(void (*kernel)(KArgs...), int grid, int block, cudaStream_t s, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = 0; cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

int Solver::capture_cholesky_graph() {
  static const bool use_pdl = !getenv("GPBA_NO_PDL");
  double* bs = d_hs.p + (size_t)n_hs * 144;
  const CholView C = chol_view();
  auto finish = [&](cudaError_t e, cudaGraphExec_t* exec) -> int {
    cudaGraph_t graph = nullptr;
    cudaError_t e2 = cudaStreamEndCapture(stream, &graph);
    if (e != cudaSuccess || e2 != cudaSuccess) {
      if (graph) cudaGraphDestroy(graph);
      g_err = std::string("cholesky graph capture: ") + cudaGetErrorString(e != cudaSuccess ? e : e2);
      return GPBA_ERR_CUDA;
    }
    e = cudaGraphInstantiate(exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) { *exec = nullptr; g_err = std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e); return GPBA_ERR_CUDA; }
    return GPBA_OK;
  };
  // ---- graph A: load + factorization + forward substitution
  CK(cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal));
  int launches = 0;
  cudaError_t e = cudaMemsetAsync(d_tiles.p, 0, sizeof(double) * (size_t)chol_doubles, stream);
  // chunk counters of the left-looking update: reset by every replay of the graph (a memset node in front of the kernels)
  if (e == cudaSuccess) e = cudaMemsetAsync(d_lu_counter.p, 0, sizeof(int) * 2 * (size_t)std::max<size_t>(lvl_ncols.size(), 1), stream);
  if (e == cudaSuccess) {
    const int64_t work = std::max((int64_t)n_hs * 144, (int64_t)NT * GPBA_NB);
    k_chol_load<<<(int)std::min((work + 255) / 256, (int64_t)148 * 8), 256, 0, stream>>>(C, n_hs, d_hs_row.p, d_hs_col.p, d_hs.p, bs);
    ++launches;
    const int n_levels = (int)lvl_ncols.size();
    static const bool per_level = getenv("GPBA_CHOL_LEVELS") != nullptr;   // the level-by-level launch sequence (A/B comparisons)
    if (!per_level) {
      // one persistent dataflow kernel: as many CTAs as are resident at once (the waits inside rely on that)
      void (*kern)(CholView, ChFactorArgs) = k_chol_factor;
      const size_t cf_smem = (size_t)GPBA_CF_STAGES * GPBA_CF_STAGE_BYTES + 4 * (((size_t)NT + 31) / 32);
      e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cf_smem);
      int per_sm = 0, sms = 148;
      if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, GPBA_CF_THREADS, cf_smem);
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
      if (e == cudaSuccess && per_sm < 1) e = cudaErrorLaunchOutOfResources;
      if (e == cudaSuccess) e = cudaMemsetAsync(d_cf_cnt.p, 0, sizeof(int) * (3 * (size_t)NT + 1), stream);
      if (e == cudaSuccess) {
        static const int cap = getenv("GPBA_CF_CTAS_PER_SM") ? atoi(getenv("GPBA_CF_CTAS_PER_SM")) : 8;
        const int grid = std::max(1, std::min(std::min(per_sm, cap) * sms, cf_tasks));
        if (getenv("GPBA_VERBOSE")) fprintf(stderr, "[gpba] cholesky: persistent factorization, %d tasks, %d CTAs (%d per SM), %zu B smem\n", cf_tasks, grid, per_sm, cf_smem);
        ChFactorArgs F;
        F.tab = d_cf_tab.p; F.n_tasks = cf_tasks; F.prod = d_cf_prod.p;
        F.upd_need = d_cf_need.p; F.pan_need = d_cf_need.p + NT;
        F.upd_cnt = d_cf_cnt.p; F.pan_cnt = d_cf_cnt.p + NT; F.diag_cnt = d_cf_cnt.p + 2 * (size_t)NT; F.task_counter = d_cf_cnt.p + 3 * (size_t)NT;
        F.d8 = d_cf_d8.p;
        F.fail = d_fail.p;
        F.trace = nullptr;
        if (getenv("GPBA_CF_TRACE")) { if (d_cf_trace.alloc(8 * (size_t)cf_tasks) == GPBA_OK) { F.trace = d_cf_trace.p; cudaMemsetAsync(d_cf_trace.p, 0, 64 * (size_t)cf_tasks, stream); } }
        // cooperative launch: the runtime refuses a grid that is not resident in full instead of letting the waits spin
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid); cfg.blockDim = dim3(GPBA_CF_THREADS); cfg.dynamicSmemBytes = cf_smem; cfg.stream = stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeCooperative;
        at[0].val.cooperative = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        e = cudaLaunchKernelEx(&cfg, kern, C, F);
        ++launches;
        if (e == cudaSuccess) e = cudaGetLastError();
      }
    }
    const size_t lu_smem = (size_t)GPBA_LU_STAGES * (2 * GPBA_TILE_BYTES + GPBA_NB * 8);
    e = cudaFuncSetAttribute(k_chol_lupdate, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lu_smem);
    int lu_ctas = 148;
    cudaDeviceGetAttribute(&lu_ctas, cudaDevAttrMultiProcessorCount, device);
    for (int l = 0; per_level && l < n_levels && e == cudaSuccess; ++l) {
      const int npan = lvl_pan_begin[l + 1] - lvl_pan_begin[l], nupd = lvl_upd_begin[l + 1] - lvl_upd_begin[l];
      const int2* pt = d_pan_tab.p + lvl_pan_begin[l];
      const int4* ut = d_upd_tab.p + lvl_upd_begin[l];
      // left-looking: the level's tiles first receive the products of all finished columns, then the panel step
      if (nupd > 0) {
        const int grid = std::min(nupd, lu_ctas);   // persistent CTAs, one per SM, drawing chunks from the level's counter
        // completion counter of the previous level's panel grid and the number of its CTAs (l >= 1 here: level 0 has no sources)
        int* prev_done = d_lu_counter.p + n_levels + (l - 1);
        const int prev_count = lvl_pan_begin[l] - lvl_pan_begin[l - 1];
        if (use_pdl) e = launch_pdl_smem(k_chol_lupdate, grid, GPBA_LU_THREADS, lu_smem, stream, C, ut, nupd, (const unsigned*)d_klist.p, d_lu_counter.p + l, (const int*)prev_done, prev_count, d_fail.p);
        else k_chol_lupdate<<<grid, GPBA_LU_THREADS, lu_smem, stream>>>(C, ut, nupd, d_klist.p, d_lu_counter.p + l, prev_done, prev_count, d_fail.p);
        ++launches;
        if (e != cudaSuccess) break;
      }
      if (use_pdl && l > 0) e = launch_pdl(k_chol_panel, npan, GPBA_PANEL_THREADS, stream, C, pt, d_fail.p, d_lu_counter.p + n_levels + l);
      else k_chol_panel<<<npan, GPBA_PANEL_THREADS, 0, stream>>>(C, pt, d_fail.p, d_lu_counter.p + n_levels + l);
      ++launches;
    }
    if (e == cudaSuccess) e = cudaGetLastError();
  }
  CKR(finish(e, &chol_graph));
  chol_graph_launches = launches;
  // ---- graph B: backward substitution
  CK(cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal));
  launches = 0;
  e = cudaSuccess;
  for (int l = (int)lvl_ncols.size() - 1; l >= 0 && e == cudaSuccess; --l) {
    const int nb = lvl_back_begin[l + 1] - lvl_back_begin[l];
    const int2* bt = d_back_tab.p + lvl_back_begin[l];
    const bool first = l == (int)lvl_ncols.size() - 1;
    if (lvl_ncols[l] > 1) {
      if (use_pdl && !first) e = launch_pdl(k_chol_back<true>, nb, 192, stream, C, bt);
      else k_chol_back<true><<<nb, 192, 0, stream>>>(C, bt);
    } else {
      if (use_pdl && !first) e = launch_pdl(k_chol_back<false>, nb, 192, stream, C, bt);
      else k_chol_back<false><<<nb, 192, 0, stream>>>(C, bt);
    }
    ++launches;
  }
  if (e == cudaSuccess) {
    if (use_pdl) e = launch_pdl(k_chol_unpermute, (n_pose * 12 + 255) / 256, 256, stream, C, d_x.p);
    else k_chol_unpermute<<<(n_pose * 12 + 255) / 256, 256, 0, stream>>>(C, d_x.p);
    ++launches;
  }
  if (e == cudaSuccess) e = cudaGetLastError();
  CKR(finish(e, &chol_back_graph));
  chol_back_launches = launches;
  return GPBA_OK;
}

// per-camera constants of a state buffer from its extrinsics (after a state reload: pop, reset)
__global__ void k_refresh_cams(int n_cam, const double* __restrict__ ext, CamConst* __restrict__ cam) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_cam) return;
  CamConst cc = cam[c];
  cam_from_tbc(load_se3(ext + 7 * c), cc);
  cam[c] = cc;
}
int Solver::refresh_cams() {
  for (int b = 0; b < 2; ++b) { k_refresh_cams<<<(n_cam + 31) / 32, 32, 0, stream>>>(n_cam, d_ext[b].p, d_cam[b].p); CK(cudaGetLastError()); }
  return GPBA_OK;
}

// computeLambdaInit helpers: diagonal of the pose blocks, max |diagonal| of the landmark blocks
__global__ void k_hpp_diag(int n_pose, const int* __restrict__ pose_diag, const double* __restrict__ hpp, double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_pose * 12) out[i] = hpp[(size_t)pose_diag[i / 12] * 144 + (i % 12) * 13];
}
__global__ void __launch_bounds__(256) k_hll_absmax(int n_lm, const double* __restrict__ hll, double* __restrict__ partial) {
  __shared__ double red[256];
  double m = 0.0;
  for (int l = blockIdx.x * blockDim.x + threadIdx.x; l < n_lm; l += gridDim.x * blockDim.x)
    m = fmax(m, fmax(fabs(hll[9 * (size_t)l]), fmax(fabs(hll[9 * (size_t)l + 4]), fabs(hll[9 * (size_t)l + 8]))));
  red[threadIdx.x] = m;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) { if ((int)threadIdx.x < o) red[threadIdx.x] = fmax(red[threadIdx.x], red[threadIdx.x + o]); __syncthreads(); }
  if (threadIdx.x == 0) partial[blockIdx.x] = red[0];
}

int Solver::compute_records(int buf, bool full) {
  t0();
  if (full) k_records<true><<<(n_rec + 63) / 64, 64, 0, stream>>>(Vb(buf), d_pose[buf].p, d_vel[buf].p, d_rec.p, d_rec_lite.p);
  else k_records<false><<<(n_rec + 63) / 64, 64, 0, stream>>>(Vb(buf), d_pose[buf].p, d_vel[buf].p, d_rec_lite.p);
  CK(cudaGetLastError());
  t1(0, 1);
  return GPBA_OK;
}

// [chi2, scale, fail, stop] of one LM trial leave the device in ONE copy (and, on several GPUs, ONE all-reduce): the
// failure flag of the solve and the caller's stop flag ride along as doubles, so every rank takes the same accept /
// reject / stop decision from the same four sums (a rank-local flag would let the replicated LM loops diverge).
__global__ void k_pack_trial(double* __restrict__ scal, const int* __restrict__ fail, int trial, double stop) {
  if (threadIdx.x == 0) {
    if (!trial) scal[1] = 0.0;
    scal[2] = trial && fail[0] ? 1.0 : 0.0;
    scal[3] = stop;
  }
}

// SparseOptimizer::computeActiveErrors + activeRobustChi2 on state buffer `buf`.  trial: the evaluation closes an LM
// trial (solve + apply_update were enqueued before): the same host read also brings computeScale and the failure flag.
int Solver::compute_errors(int buf, bool store, double* chi2, bool trial, const volatile unsigned char* stop) {
  CKR(compute_records(buf, false));
  t0();
  double* out = store ? (chi2_store_override ? chi2_store_override : d_chi2.p) : nullptr;
  if (stereo) k_residual<true><<<grid_obs, 256, 0, stream>>>(Vb(buf), d_rec_lite.p, d_ptS[buf].p, d_partial.p, out);
  else k_residual<false><<<grid_obs, 256, 0, stream>>>(Vb(buf), d_rec_lite.p, d_ptS[buf].p, d_partial.p, out);
  CK(cudaGetLastError());
  const int np = n_prior + (n_velp + 63) / 64;
  const bool priors_here = rank == 0;  // priors are replicated: counted once
  if (np > 0 && priors_here) {
    k_priors<<<np, 64, 0, stream>>>(V, d_pose[buf].p, d_vel[buf].p, 0, d_prior_rho.p, nullptr, nullptr);
    CK(cudaGetLastError());
  }
  if (priors_here) {   // EdgeExtrinsicPrior of the free extrinsics (zero for the others)
    k_ext_prior<<<(n_cam + 31) / 32, 32, 0, stream>>>(V, d_ext[buf].p, 0, d_prior_rho.p + n_prior + n_velp, nullptr, nullptr);
    CK(cudaGetLastError());
  }
  k_reduce<<<1, 256, 0, stream>>>(d_partial.p, n_aobs > 0 ? grid_obs : 0, d_prior_rho.p, priors_here ? n_prior + n_velp + n_cam : 0, d_scal.p);
  CK(cudaGetLastError());
  k_pack_trial<<<1, 32, 0, stream>>>(d_scal.p, d_fail.p, trial ? 1 : 0, (stop && *stop) ? 1.0 : 0.0);
  CK(cudaGetLastError());
  t1(1, 5);
  if (nranks > 1) CKR(allreduce_scalar(d_scal.p, 4));
  CK(cudaMemcpyAsync(h_scal, d_scal.p, 4 * sizeof(double), cudaMemcpyDeviceToHost, stream));
  CK(cudaStreamSynchronize(stream));
  *chi2 = h_scal[0];
  trial_scale = h_scal[1]; trial_failed = h_scal[2] != 0.0; stop_seen = h_scal[3] != 0.0;
  last_eval = buf;
  return GPBA_OK;
}

// BlockSolver::buildSystem at the current state.
int Solver::build_system() {
  CKR(compute_records(cur, true));
  t0();
  CK(cudaMemsetAsync(d_recS.p, 0, sizeof(double) * 27 * (size_t)n_rec, stream));
  CK(cudaMemsetAsync(d_hpp.p, 0, sizeof(double) * 144 * (size_t)n_hpp, stream));
  CK(cudaMemsetAsync(d_bp.p, 0, sizeof(double) * 12 * (size_t)std::max(n_pose, 1), stream));
  int launches = 0;
  if (n_lm > 0) {
    CK(cudaMemsetAsync(d_hll.p, 0, sizeof(double) * 9 * (size_t)n_lm, stream));
    CK(cudaMemsetAsync(d_bl.p, 0, sizeof(double) * 3 * (size_t)n_lm, stream));
    const int g = std::max(1, std::min(n_tiles, 148 * 5));
    if (stereo) k_lin_points<true><<<g, GPBA_K2_THREADS, 0, stream>>>(Vb(cur), d_rec_lite.p, d_ptS[cur].p, d_hll.p, d_bl.p, d_W.p);
    else k_lin_points<false><<<g, GPBA_K2_THREADS, 0, stream>>>(Vb(cur), d_rec_lite.p, d_ptS[cur].p, d_hll.p, d_bl.p, d_W.p);
    CK(cudaGetLastError());
    t1(2, 1);
    t0();
    const int g2 = std::min((n_rseg + 3) / 4, 148 * 16);
    if (stereo) k_lin_records<true><<<g2, 128, 0, stream>>>(Vb(cur), d_rec.p, d_ptS[cur].p, d_recS.p, d_r_u.p, d_r_v.p, d_r_ur.p, d_r_w.p, d_r_lm.p, d_r_flags.p);
    else k_lin_records<false><<<g2, 128, 0, stream>>>(Vb(cur), d_rec.p, d_ptS[cur].p, d_recS.p, d_r_u.p, d_r_v.p, nullptr, d_r_w.p, d_r_lm.p, d_r_flags.p);
    CK(cudaGetLastError());
    k_rec_to_hpp<<<n_rec, 128, 0, stream>>>(V, d_rec.p, d_recS.p, d_hpp.p, d_bp.p);
    CK(cudaGetLastError());
    launches += 2;
  }
  const int np = n_prior + (n_velp + 63) / 64;
  if (np > 0 && rank == 0) {
    k_priors<<<np, 64, 0, stream>>>(V, d_pose[cur].p, d_vel[cur].p, 1, nullptr, d_hpp.p, d_bp.p);
    CK(cudaGetLastError());
    launches++;
  }
  if (rank == 0 && n_pose > n_pose_kf) {
    k_ext_prior<<<(n_cam + 31) / 32, 32, 0, stream>>>(V, d_ext[cur].p, 1, nullptr, d_hpp.p, d_bp.p);
    CK(cudaGetLastError());
    launches++;
  }
  t1(3, launches);
  system_ok = true;
  return GPBA_OK;
}

int Solver::allreduce_system() {
  if (nranks <= 1) return GPBA_OK;
  t0();
  const size_t count = (size_t)n_hs * 144 + (size_t)n_pose * 12;
  int rc = g_nccl.AllReduce(d_hs.p, d_hs.p, count, /*ncclDouble*/ 8, /*ncclSum*/ 0, comm, stream);
  if (rc != 0) { g_err = std::string("ncclAllReduce: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?"); return GPBA_ERR_NCCL; }
  t1(9, 1);
  return GPBA_OK;
}
int Solver::allreduce_scalar(double* v, int count, int op) {
  t0();
  int rc = g_nccl.AllReduce(v, v, (size_t)count, kNcclDouble, op, comm, stream);
  if (rc != 0) { g_err = "ncclAllReduce(scalar) failed"; return GPBA_ERR_NCCL; }
  t1(9, 1);
  return GPBA_OK;
}

// BlockSolver::solve with lambda on the diagonals of Hpp and Hll (setLambda folded in: block_solver.hpp:563-589,353-486).
int Solver::solve(double lambda) {
  double* bs = d_hs.p + (size_t)n_hs * 144;  // bschur lives right behind the Hschur values (one allreduce)
  CK(cudaMemsetAsync(d_fail.p, 0, sizeof(int), stream));
  t0();
  // K4a + K4b: per-landmark factor and the record-pair products
  if (n_lm > 0) {
    k_schur_prep<<<std::min((n_lm + 15) / 16, 148 * 16), 128, 0, stream>>>(V, lambda, d_hll.p, d_bl.p, d_W.p, d_U.p, d_ptL.p, d_fail.p);
    CK(cudaGetLastError());
  }
  t1(4, n_lm > 0 ? 1 : 0);
  t0();
  int launches = 0;
  if (n_items > 0) {
    CK(cudaMemsetAsync(d_C.p, 0, sizeof(double) * (size_t)n_rp * GPBA_RP_STRIDE, stream));
    CK(cudaMemsetAsync(d_fail.p + 1, 0, sizeof(int), stream));
    static const int pairs_ctas_per_sm = getenv("GPBA_PAIRS_CTAS") ? atoi(getenv("GPBA_PAIRS_CTAS")) : 16;
    static const int pairs_batch = getenv("GPBA_PAIRS_BATCH") ? std::max(1, atoi(getenv("GPBA_PAIRS_BATCH"))) : 16;   // measured at C4: 8 -> 7.9, 16 -> 7.4, 32 -> 8.4, 64 -> 11.2 ms per optimize
    // Measured alternatives at C4 (profiles/r02_k4b_experiments.txt): CTAs of 512 / 1024 threads sharing 64..256 items (first-side rows
    // kept in one SM's L1) 8.1 .. 12.7 ms per optimize; first-side rows of a (chunk, first record) group staged in shared memory with
    // cp.async.bulk, second side streamed, 7.8 .. 14.4 ms; this kernel 7.3 .. 7.5 ms.
    k_schur_pairs<<<std::min((n_items + 3) / 4, 148 * pairs_ctas_per_sm), 128, 0, stream>>>(n_items, d_item_rp.p, d_item_begin.p, d_item_end.p, d_item_flags.p,
                                                                             d_pairs.p, d_o_lm.p, d_U.p, d_ptL.p, d_C.p, d_fail.p + 1, pairs_batch);
    CK(cudaGetLastError());
    ++launches;
  }
  t1(5, launches);
  t0();
  launches = 0;
  // K4c: rank 0 carries lambda (pose priors / damping must enter the sum exactly once); the GP-edge part of
  // Hpp is a per-rank partial, so every rank adds its own Hpp but only rank 0 adds lambda.
  if (n_hs > 0) {
    k_schur_expand<<<(n_hs + GPBA_K4C_WARPS - 1) / GPBA_K4C_WARPS, 32 * GPBA_K4C_WARPS, 0, stream>>>(V, rank == 0 ? lambda : 0.0, d_rec.p, d_hpp.p, d_bp.p, d_con_begin.p, d_con.p, d_C.p, d_hs.p, bs);
    CK(cudaGetLastError());
    ++launches;
  }
  t1(10, launches);
  CKR(allreduce_system());
  t0();
  launches = 0;
  if (linear_solver != GPBA_SOLVER_PCG) {
    if (!chol_graph) CKR(capture_cholesky_graph());
    CK(cudaGraphLaunch(chol_graph, stream));
    t1(6, chol_graph_launches);
    if (d_cf_trace.p && getenv("GPBA_CF_TRACE")) {   // tools only: dump the task timeline of this factorization
      std::vector<long long> tr(8 * (size_t)cf_tasks);
      CK(cudaStreamSynchronize(stream));
      CK(cudaMemcpy(tr.data(), d_cf_trace.p, tr.size() * sizeof(long long), cudaMemcpyDeviceToHost));
      std::vector<int4> tab((size_t)cf_tasks);
      CK(cudaMemcpy(tab.data(), d_cf_tab.p, tab.size() * sizeof(int4), cudaMemcpyDeviceToHost));
      if (FILE* f = fopen(getenv("GPBA_CF_TRACE"), "wb")) {
        fwrite(&cf_tasks, sizeof(int), 1, f); fwrite(tab.data(), sizeof(int4), tab.size(), f); fwrite(tr.data(), sizeof(long long), tr.size(), f);
        fclose(f);
      }
    }
    t0();
    CK(cudaGraphLaunch(chol_back_graph, stream));
    t1(7, chol_back_launches);
  } else {
    int it = 0;
    CKR(pcg.solve(n_pose, n_hs, d_hs.p, bs, d_x.p, stream, &it, d_fail.p));
    launches += 1;
    t1(6, launches);
  }
  // Levels that hold several tile columns accumulate with floating-point atomics, whose order is not fixed: the ranks
  // of a multi-GPU run would keep replicated poses that differ in the last bits.  Rank 0's solution is everybody's.
  if (nranks > 1 && n_pose > 0) {
    t0();
    if (g_nccl.Broadcast(d_x.p, d_x.p, (size_t)n_pose * 12, kNcclDouble, 0, comm, stream) != 0) { g_err = "ncclBroadcast(x) failed"; return GPBA_ERR_NCCL; }
    t1(9, 1);
  }
  return GPBA_OK;
}

// solve()'s failure flag (a non-positive landmark block or reduced system), agreed on by all ranks: used by the L1 path;
// the LM loop gets the same flag with the trial's chi2 (compute_errors).
int Solver::read_fail(bool* ok) {
  if (nranks > 1) {
    k_pack_trial<<<1, 32, 0, stream>>>(d_scal.p + 4, d_fail.p, 1, 0.0);
    CK(cudaGetLastError());
    CKR(allreduce_scalar(d_scal.p + 6, 1));
    CK(cudaMemcpyAsync(h_scal + 6, d_scal.p + 6, sizeof(double), cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    *ok = h_scal[6] == 0.0;
    return GPBA_OK;
  }
  CK(cudaMemcpyAsync(h_fail, d_fail.p, sizeof(int), cudaMemcpyDeviceToHost, stream));
  CK(cudaStreamSynchronize(stream));
  *ok = (*h_fail == 0);
  return GPBA_OK;
}

// SparseOptimizer::update: cur (+) x -> the other buffer; also computeScale's sum x (lambda x + b).
int Solver::apply_update(double lambda, double* scale) {
  const int nb = 1 - cur;
  t0();
  int gp = 0;
  if (n_lm > 0) {
    gp = std::min((n_lm + 15) / 16, 148 * 16);   // four landmarks per warp, four warps per CTA
    k_rec_y<<<(n_rec * 6 + 127) / 128, 128, 0, stream>>>(V, d_rec.p, d_x.p, d_Y.p);
    k_backsub<<<gp, 128, 0, stream>>>(V, lambda, d_U.p, d_ptL.p, d_bl.p, d_Y.p, d_ptS[cur].p, d_ptS[nb].p, d_xl.p, d_partial.p);
    CK(cudaGetLastError());
  }
  // computeScale = sum x (lambda x + b) over poses and landmarks (optimization_algorithm_levenberg.cpp:187-194).  Under
  // sharding b_p is a per-rank partial (the GP-edge part of the rank's own landmarks; the priors live on rank 0), so every
  // rank contributes x_p . b_p^(rank) and rank 0 alone adds lambda |x_p|^2: the all-reduced sum is x_p . (lambda x_p + b_p).
  k_update_poses<<<(n_kf + 63) / 64, 64, 0, stream>>>(V, rank == 0 ? lambda : 0.0, d_x.p, d_bp.p, d_pose[cur].p, d_vel[cur].p, d_pose[nb].p, d_vel[nb].p, d_pose_scale.p);
  CK(cudaGetLastError());
  if (n_pose > n_pose_kf) {   // free extrinsics: Tbc (+) x and the per-camera constants of the trial state
    k_update_ext<<<(n_cam + 31) / 32, 32, 0, stream>>>(V, rank == 0 ? lambda : 0.0, d_x.p, d_bp.p, d_ext[cur].p, d_ext[nb].p, d_cam[cur].p, d_cam[nb].p, d_pose_scale.p);
    CK(cudaGetLastError());
  }
  k_reduce<<<1, 256, 0, stream>>>(d_partial.p, gp, d_pose_scale.p, n_pose, d_scal.p + 1);
  CK(cudaGetLastError());
  t1(8, 4);
  (void)scale;   // read back together with the trial's chi2 (compute_errors(trial = true))
  return GPBA_OK;
}

// OptimizationAlgorithmLevenberg::solve (optimization_algorithm_levenberg.cpp:61-169)
int Solver::lm_solve(int iteration, const gpba_lm_params& P, const volatile unsigned char* stop, gpba_lm_trace* tr, int* result) {
  if (iteration == 0 && structure_dirty) CKR(build_structure());
  double currentChi = 0;
  // computeActiveErrors + activeRobustChi2 at the estimate (:72-73).  After an accepted trial the estimate IS the state
  // that trial evaluated, so the value is already known and the pass over the observations is skipped.
  if (chi_cache_valid && last_eval == cur) currentChi = chi_cache;
  else CKR(compute_errors(cur, false, &currentChi));
  chi_cache_valid = false;
  double tempChi = currentChi;
  const double iniChi = currentChi;
  CKR(build_system());
  if (iteration == 0) {
    if (lambda_init > 0) lambda_cur = lambda_init;
    else {  // computeLambdaInit: tau * max |H_jj| over poses and landmarks (:171-185)
      // the pose diagonal of a multi-GPU run is the SUM of the ranks' partial Hpp, the landmark diagonal is rank-local:
      // sum-reduce the first, max-reduce the result, so every rank starts from the same lambda
      std::vector<double> hd((size_t)std::max(n_pose, 1) * 12, 0.0);
      DBuf<double> d_diag;
      CKR(d_diag.alloc(hd.size() + 1));
      const int gl = n_lm > 0 ? std::min((n_lm + 255) / 256, 148 * 8) : 0;
      if (n_pose > 0) { k_hpp_diag<<<(n_pose * 12 + 255) / 256, 256, 0, stream>>>(n_pose, d_pose_hpp_diag.p, d_hpp.p, d_diag.p); CK(cudaGetLastError()); }
      if (gl > 0) { k_hll_absmax<<<gl, 256, 0, stream>>>(n_lm, d_hll.p, d_partial.p); CK(cudaGetLastError()); }
      if (nranks > 1 && n_pose > 0) CKR(allreduce_scalar(d_diag.p, n_pose * 12, kNcclSum));
      std::vector<double> part((size_t)std::max(gl, 1), 0.0);
      if (n_pose > 0) CK(cudaMemcpyAsync(hd.data(), d_diag.p, sizeof(double) * 12 * (size_t)n_pose, cudaMemcpyDeviceToHost, stream));
      if (gl > 0) CK(cudaMemcpyAsync(part.data(), d_partial.p, sizeof(double) * gl, cudaMemcpyDeviceToHost, stream));
      CK(cudaStreamSynchronize(stream));
      double mx = 0;
      for (double v : hd) mx = std::max(mx, std::fabs(v));
      for (double v : part) mx = std::max(mx, v);
      if (nranks > 1) {
        h_scal[8] = mx;
        CK(cudaMemcpyAsync(d_scal.p + 7, h_scal + 8, sizeof(double), cudaMemcpyHostToDevice, stream));
        CKR(allreduce_scalar(d_scal.p + 7, 1, kNcclMax));
        CK(cudaMemcpyAsync(h_scal + 8, d_scal.p + 7, sizeof(double), cudaMemcpyDeviceToHost, stream));
        CK(cudaStreamSynchronize(stream));
        mx = h_scal[8];
      }
      lambda_cur = P.tau * mx;
    }
    ni = 2; nBad = 0;
  }
  double rho = 0;
  int qmax = 0;
  bool accepted = false;
  do {
    // one trial = solve + update + evaluation enqueued back to back, ONE host synchronisation at its end
    CKR(solve(lambda_cur));
    CKR(apply_update(lambda_cur, nullptr));
    CKR(compute_errors(1 - cur, false, &tempChi, true, stop));
    if (trial_failed) tempChi = std::numeric_limits<double>::max();   // solve() == false (:110-113)
    double scale = trial_scale;
    rho = (currentChi - tempChi);
    scale += 1e-3;
    rho /= scale;
    accepted = rho > 0 && std::isfinite(tempChi);
    if (accepted) {
      double alpha = 1. - std::pow((2 * rho - 1), 3);
      alpha = (std::min)(alpha, P.good_step_upper);
      const double scaleFactor = (std::max)(P.good_step_lower, alpha);
      lambda_cur *= scaleFactor;
      ni = 2;
      currentChi = tempChi;
      cur = 1 - cur;  // discardTop(): the trial buffer becomes the estimate
    } else {
      lambda_cur *= ni;
      ni *= 2;       // pop(): the estimate buffer is untouched
    }
    qmax++;
  } while (rho < 0 && qmax < P.max_trials_after_failure && !stop_seen);
  if (accepted) { chi_cache_valid = true; chi_cache = currentChi; }
  if (tr && iteration < GPBA_MAX_ITERS) {
    tr->levenberg_iterations[iteration] = qmax;
    tr->chi2_before[iteration] = iniChi;
    tr->chi2_after[iteration] = currentChi;
    tr->lambda[iteration] = lambda_cur;
    tr->total_trials += qmax;
    tr->last_trial_chi2 = tempChi;
  }
  *result = GPBA_RESULT_OK;
  if (qmax == P.max_trials_after_failure || rho == 0) { *result = GPBA_TERMINATE; return GPBA_OK; }
  if ((iniChi - currentChi) * 1e3 < iniChi) nBad++; else nBad = 0;
  if (nBad >= 3) *result = GPBA_TERMINATE;
  return GPBA_OK;
}

int Solver::scatter_points(int buf) {
  if (n_lm > 0) {
    k_scatter_pts<<<(n_lm * 3 + 255) / 256, 256, 0, stream>>>(n_lm, d_lm_pt.p, d_ptS[buf].p, d_pt_full.p);
    CK(cudaGetLastError());
  }
  return GPBA_OK;
}

int Solver::optimize(int iters, const volatile unsigned char* stop, const gpba_lm_params& P, gpba_lm_trace* tr) {
  CK(cudaSetDevice(device));
  if (tr) { std::memset(tr, 0, sizeof(*tr)); tr->result = GPBA_RESULT_OK; }
  pcg.total_iterations = 0;
  int cj = 0, result = GPBA_RESULT_OK;
  bool ok = true;
  chi_cache_valid = false;
  // terminate() (sparse_optimizer.h:188) is polled where the reference polls it -- between outer iterations and in the
  // trial loop -- but on several GPUs every rank must see the same answer: the flag is sampled when a trial's scalars are
  // packed and summed over the ranks with them (stop_seen); before the first trial a one-scalar all-reduce does the same.
  stop_seen = stop && *stop;
  if (nranks > 1 && stop) {
    h_scal[8] = stop_seen ? 1.0 : 0.0;
    CK(cudaMemcpyAsync(d_scal.p + 7, h_scal + 8, sizeof(double), cudaMemcpyHostToDevice, stream));
    CKR(allreduce_scalar(d_scal.p + 7, 1));
    CK(cudaMemcpyAsync(h_scal + 8, d_scal.p + 7, sizeof(double), cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    stop_seen = h_scal[8] != 0.0;
  }
  for (int i = 0; i < iters && !stop_seen && ok; ++i) {
    CKR(lm_solve(i, P, stop, tr, &result));
    ok = (result == GPBA_RESULT_OK);
    ++cj;
  }
  if (tr) { tr->n_iters = cj; tr->result = result; tr->cg_iterations = pcg.total_iterations; }
  if (structure_ok) {
    // stale-error quirk: edge errors are those of the LAST EVALUATED trial, even if it was rejected
    double dummy;
    const int ev = last_eval;
    if (nranks > 1) {
      // every rank ends with the full state: stored edge errors and landmark positions of the other ranks' shards
      // arrive through two all-reduces of zero-initialised staging buffers (each entry is owned by exactly one rank)
      const size_t ng = (size_t)std::max<int64_t>(std::max<int64_t>(n_obs, (int64_t)n_pt * 3), 1);
      CKR(d_gather.alloc(ng));
      CK(cudaMemsetAsync(d_gather.p, 0, sizeof(double) * (size_t)std::max<int64_t>(n_obs, 1), stream));
      chi2_store_override = d_gather.p;
      const int rc = compute_errors(ev, true, &dummy);
      chi2_store_override = nullptr;
      CKR(rc);
      if (n_obs > 0) {
        if (g_nccl.AllReduce(d_gather.p, d_gather.p, (size_t)n_obs, 8, 0, comm, stream) != 0) { g_err = "ncclAllReduce(edge chi2) failed"; return GPBA_ERR_NCCL; }
        k_merge_chi2<<<(int)std::min<int64_t>((n_obs + 255) / 256, 148 * 8), 256, 0, stream>>>(n_obs, d_all_flags.p, d_gather.p, d_chi2.p);
      }
      CK(cudaMemsetAsync(d_gather.p, 0, sizeof(double) * 3 * (size_t)std::max(n_pt, 1), stream));
      if (n_lm > 0) k_scatter_pts<<<(n_lm * 3 + 255) / 256, 256, 0, stream>>>(n_lm, d_lm_pt.p, d_ptS[cur].p, d_gather.p);
      if (n_pt > 0) {
        if (g_nccl.AllReduce(d_gather.p, d_gather.p, (size_t)n_pt * 3, 8, 0, comm, stream) != 0) { g_err = "ncclAllReduce(points) failed"; return GPBA_ERR_NCCL; }
        k_merge_pts<<<(n_pt * 3 + 255) / 256, 256, 0, stream>>>(n_pt, d_pt_act.p, d_gather.p, d_pt_full.p);
      }
      CK(cudaGetLastError());
      pts_gathered = true;
    } else {
      CKR(compute_errors(ev, true, &dummy));
      CKR(scatter_points(cur));
    }
    CK(cudaStreamSynchronize(stream));
  }
  return GPBA_OK;
}

int Solver::download_state(double* kf_pose, double* kf_vel, double* pt_xyz) {
  CK(cudaSetDevice(device));
  if (structure_ok) CKR(scatter_points(cur));
  if (kf_pose) CK(cudaMemcpyAsync(kf_pose, d_pose[cur].p, sizeof(double) * 7 * (size_t)n_kf, cudaMemcpyDeviceToHost, stream));
  if (kf_vel) CK(cudaMemcpyAsync(kf_vel, d_vel[cur].p, sizeof(double) * 6 * (size_t)n_kf, cudaMemcpyDeviceToHost, stream));
  if (pt_xyz) CK(cudaMemcpyAsync(pt_xyz, d_pt_full.p, sizeof(double) * 3 * (size_t)n_pt, cudaMemcpyDeviceToHost, stream));
  CK(cudaStreamSynchronize(stream));
  return GPBA_OK;
}

}  // namespace

// ===================================================================================== C ABI
struct gpba_handle { Solver s; };
#define S(h) ((h)->s)
#define NEED(h) do { if (!(h)) { g_err = "null handle"; return GPBA_ERR_INVALID; } g_alloc_stream = S(h).stream; } while (0)
#define NEED_STRUCT(h) do { NEED(h); if (!S(h).structure_ok) { g_err = "call gpba_build_structure first"; return GPBA_ERR_STATE; } if (cudaSetDevice(S(h).device) != cudaSuccess) return GPBA_ERR_CUDA; } while (0)

extern "C" {

const char* gpba_last_error(void) { return g_err.c_str(); }

void gpba_default_lm_params(gpba_lm_params* p) {
  p->max_trials_after_failure = 10; p->tau = 1e-5; p->good_step_lower = 1. / 3.; p->good_step_upper = 2. / 3.;
  p->pcg_tolerance = 1e-12; p->pcg_max_iterations = 4000;
}

static int create_impl(const gpba_problem* prob, int device, bool async_upload, gpba_handle** out) {
  if (!prob || !out) { g_err = "null argument"; return GPBA_ERR_INVALID; }
  gpba_handle* h = new (std::nothrow) gpba_handle();
  if (!h) return GPBA_ERR_INVALID;
  int rc = h->s.init(prob, device, async_upload);
  if (rc != GPBA_OK) { if (h->s.copy_holder.s) cudaStreamSynchronize(h->s.copy_holder.s); delete h; return rc; }
  *out = h;
  return GPBA_OK;
}
int gpba_create(const gpba_problem* prob, int device, gpba_handle** out) { return create_impl(prob, device, false, out); }

int gpba_destroy(gpba_handle* h) {
  if (h) {
    cudaSetDevice(h->s.device); g_alloc_stream = h->s.stream;
    if (h->s.copy_holder.s) cudaStreamSynchronize(h->s.copy_holder.s);   // no copy may be in flight when the buffers go
    if (h->s.ev_meas) { cudaEventDestroy(h->s.ev_meas); h->s.ev_meas = nullptr; }
    const bool verbose = getenv("GPBA_VERBOSE") != nullptr;
    auto t0 = std::chrono::steady_clock::now();
    if (h->s.chol_graph) { cudaGraphExecDestroy(h->s.chol_graph); h->s.chol_graph = nullptr; }
    if (h->s.chol_back_graph) { cudaGraphExecDestroy(h->s.chol_back_graph); h->s.chol_back_graph = nullptr; }
    auto t1 = std::chrono::steady_clock::now();
    delete h;
    auto t2 = std::chrono::steady_clock::now();
    if (verbose) fprintf(stderr, "[gpba] destroy: graphs %.2f ms, rest %.2f ms\n", std::chrono::duration<double, std::milli>(t1 - t0).count(), std::chrono::duration<double, std::milli>(t2 - t1).count());
  }
  return GPBA_OK;
}

int gpba_nccl_unique_id(unsigned char id_out[128]) {
  if (!g_nccl.load()) { g_err = "libnccl.so.2 not found"; return GPBA_ERR_NCCL; }
  return g_nccl.GetUniqueId(id_out) == 0 ? GPBA_OK : GPBA_ERR_NCCL;
}

// One communicator per (NCCL id, rank) and process: an ncclUniqueId can initialise a communicator only once, while a
// SLAM process runs many BA calls (one handle each, like one g2o::SparseOptimizer each) over the same set of GPUs.
static std::mutex g_comm_mutex;
static std::map<std::string, ncclComm_t> g_comms;

int gpba_create_dist(const gpba_problem* prob, int device, int rank, int nranks, const unsigned char id[128], gpba_handle** out) {
  gpba_create_options o;
  o.device = device; o.rank = rank; o.nranks = nranks; o.nccl_id = id; o.flags = 0;
  return gpba_create_ex(prob, &o, out);
}

int gpba_create_ex(const gpba_problem* prob, const gpba_create_options* opt, gpba_handle** out) {
  if (!opt) { g_err = "null argument"; return GPBA_ERR_INVALID; }
  const int device = opt->device, rank = opt->rank, nranks = opt->nranks;
  const unsigned char* id = opt->nccl_id;
  if (nranks > 1 && (!id || rank < 0 || rank >= nranks)) { g_err = "bad rank / NCCL id"; return GPBA_ERR_INVALID; }
  int rc = create_impl(prob, device, (opt->flags & GPBA_CREATE_ASYNC_UPLOAD) != 0, out);
  if (rc != GPBA_OK) return rc;
  Solver& s = (*out)->s;
  s.rank = nranks > 1 ? rank : 0; s.nranks = nranks > 1 ? nranks : 1;
  if (nranks > 1) {
    if (!g_nccl.load()) { g_err = "libnccl.so.2 not found"; gpba_destroy(*out); *out = nullptr; return GPBA_ERR_NCCL; }
    std::lock_guard<std::mutex> lock(g_comm_mutex);
    const std::string key = std::string((const char*)id, 128) + "#" + std::to_string(rank) + "/" + std::to_string(nranks);
    auto it = g_comms.find(key);
    if (it == g_comms.end()) {
      NcclId nid;
      std::memcpy(nid.internal, id, 128);
      ncclCommInitRank_t init = (ncclCommInitRank_t)dlsym(g_nccl.lib, "ncclCommInitRank");
      ncclComm_t c = nullptr;
      int nrc = init(&c, nranks, nid, rank);
      if (nrc != 0) {
        g_err = std::string("ncclCommInitRank: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(nrc) : "?");
        gpba_destroy(*out); *out = nullptr;
        return GPBA_ERR_NCCL;
      }
      it = g_comms.emplace(key, c).first;
    }
    s.comm = it->second;
  }
  return GPBA_OK;
}

int gpba_build_structure(gpba_handle* h, gpba_structure_info* info) {
  NEED(h);
  CKR(S(h).build_structure());
  if (info) { CKR(S(h).count_hpl()); *info = S(h).info; }
  return GPBA_OK;
}
int gpba_get_hpp_pattern(gpba_handle* h, int32_t* rows, int32_t* cols) {
  NEED_STRUCT(h);
  std::memcpy(rows, S(h).hpp_row.data(), S(h).hpp_row.size() * 4); std::memcpy(cols, S(h).hpp_col.data(), S(h).hpp_col.size() * 4);
  return GPBA_OK;
}
int gpba_get_hschur_pattern(gpba_handle* h, int32_t* rows, int32_t* cols) {
  NEED_STRUCT(h);
  std::memcpy(rows, S(h).hs_row.data(), S(h).hs_row.size() * 4); std::memcpy(cols, S(h).hs_col.data(), S(h).hs_col.size() * 4);
  return GPBA_OK;
}
int gpba_compute_errors(gpba_handle* h, double* robust_chi2) {
  NEED_STRUCT(h);
  double c = 0;
  CKR(S(h).compute_errors(S(h).cur, true, &c));
  if (robust_chi2) *robust_chi2 = c;
  return GPBA_OK;
}
int gpba_build_system(gpba_handle* h) { NEED_STRUCT(h); return S(h).build_system(); }
int gpba_set_lambda(gpba_handle* h, double lambda, int) {
  NEED_STRUCT(h);
  S(h).lambda_set = lambda; S(h).lambda_applied = true;  // damping is applied inside the Schur kernels, Hpp/Hll stay undamped
  return GPBA_OK;
}
int gpba_restore_diagonal(gpba_handle* h) { NEED_STRUCT(h); S(h).lambda_applied = false; return GPBA_OK; }
int gpba_solve(gpba_handle* h, int* ok) {
  NEED_STRUCT(h);
  if (!S(h).system_ok) { g_err = "call gpba_build_system first"; return GPBA_ERR_STATE; }
  bool k = false;
  Solver& s = S(h);
  const double lam = s.lambda_applied ? s.lambda_set : 0.0;
  CKR(s.solve(lam));
  // landmark part of x (no state change): run the back-substitution into the scratch buffer
  CKR(s.apply_update(lam, nullptr));
  CKR(s.read_fail(&k));
  if (ok) *ok = k ? 1 : 0;
  return GPBA_OK;
}
// poses (12 each) then ALL active landmarks in g2o order (3 each); a rank of a multi-GPU run fills only its own landmarks
int gpba_vector_size(gpba_handle* h, int64_t* n) { NEED_STRUCT(h); *n = (int64_t)S(h).n_pose * 12 + (int64_t)S(h).n_lm_all * 3; return GPBA_OK; }
int gpba_get_x(gpba_handle* h, double* x) {
  NEED_STRUCT(h);
  Solver& s = S(h);
  CK(cudaMemcpyAsync(x, s.d_x.p, sizeof(double) * 12 * (size_t)s.n_pose, cudaMemcpyDeviceToHost, s.stream));
  std::vector<double> xl((size_t)s.n_lm * 3);
  if (s.n_lm) CK(cudaMemcpyAsync(xl.data(), s.d_xl.p, xl.size() * 8, cudaMemcpyDeviceToHost, s.stream));
  CK(cudaStreamSynchronize(s.stream));
  std::fill(x + (size_t)s.n_pose * 12, x + (size_t)s.n_pose * 12 + (size_t)s.n_lm_all * 3, 0.0);
  for (int l = 0; l < s.n_lm; ++l) for (int c = 0; c < 3; ++c) x[(size_t)s.n_pose * 12 + (size_t)s.lm_rank[l] * 3 + c] = xl[(size_t)l * 3 + c];
  return GPBA_OK;
}
int gpba_get_b(gpba_handle* h, double* b) {
  NEED_STRUCT(h);
  Solver& s = S(h);
  CK(cudaStreamSynchronize(s.stream));
  CK(cudaMemcpyAsync(b, s.d_bp.p, sizeof(double) * 12 * (size_t)s.n_pose, cudaMemcpyDeviceToHost, s.stream));
  std::vector<double> bl((size_t)s.n_lm * 3);
  if (s.n_lm) CK(cudaMemcpyAsync(bl.data(), s.d_bl.p, bl.size() * 8, cudaMemcpyDeviceToHost, s.stream));
  CK(cudaStreamSynchronize(s.stream));
  std::fill(b + (size_t)s.n_pose * 12, b + (size_t)s.n_pose * 12 + (size_t)s.n_lm_all * 3, 0.0);
  for (int l = 0; l < s.n_lm; ++l) for (int c = 0; c < 3; ++c) b[(size_t)s.n_pose * 12 + (size_t)s.lm_rank[l] * 3 + c] = bl[(size_t)l * 3 + c];
  return GPBA_OK;
}
int gpba_get_hpp(gpba_handle* h, double* blocks) {
  NEED_STRUCT(h);
  Solver& s = S(h);
  CK(cudaStreamSynchronize(s.stream));
  CK(cudaMemcpyAsync(blocks, s.d_hpp.p, sizeof(double) * 144 * (size_t)s.n_hpp, cudaMemcpyDeviceToHost, s.stream));
  CK(cudaStreamSynchronize(s.stream));
  if (s.lambda_applied)
    for (int k = 0; k < s.n_hpp; ++k) if (s.hpp_row[k] == s.hpp_col[k]) for (int j = 0; j < 12; ++j) blocks[(size_t)k * 144 + j * 13] += s.lambda_set;
  return GPBA_OK;
}
int gpba_get_hschur(gpba_handle* h, double* blocks, double* bschur) {
  NEED_STRUCT(h);
  Solver& s = S(h);
  CK(cudaStreamSynchronize(s.stream));
  if (blocks) CK(cudaMemcpyAsync(blocks, s.d_hs.p, sizeof(double) * 144 * (size_t)s.n_hs, cudaMemcpyDeviceToHost, s.stream));
  if (bschur) CK(cudaMemcpyAsync(bschur, s.d_hs.p + (size_t)s.n_hs * 144, sizeof(double) * 12 * (size_t)s.n_pose, cudaMemcpyDeviceToHost, s.stream));
  CK(cudaStreamSynchronize(s.stream));
  return GPBA_OK;
}
int gpba_get_hll(gpba_handle* h, double* blocks) {
  NEED_STRUCT(h);
  Solver& s = S(h);
  if (s.nranks > 1) { g_err = "gpba_get_hll is a single-GPU parity accessor (landmarks are sharded)"; return GPBA_ERR_STATE; }
  CK(cudaStreamSynchronize(s.stream));
  std::vector<double> hl((size_t)s.n_lm * 9);
  if (s.n_lm) CK(cudaMemcpyAsync(hl.data(), s.d_hll.p, hl.size() * 8, cudaMemcpyDeviceToHost, s.stream));
  CK(cudaStreamSynchronize(s.stream));
  for (int l = 0; l < s.n_lm; ++l) {
    for (int c = 0; c < 9; ++c) blocks[(size_t)s.lm_rank[l] * 9 + c] = hl[(size_t)l * 9 + c];
    if (s.lambda_applied) for (int j = 0; j < 3; ++j) blocks[(size_t)s.lm_rank[l] * 9 + j * 4] += s.lambda_set;
  }
  return GPBA_OK;
}
int gpba_get_hpl(gpba_handle* h, int64_t* lm_begin, int32_t* pose, double* blocks) {  // (landmark, pose) order of the g2o landmark numbering
  // Debug / parity accessor: the product path never forms Hpl.  Hpl_(pose,l) = sum over the landmark's observations whose
  // record touches the pose of (M_r^(pose))^T W_o, assembled here on the host from W and the record table.
  NEED_STRUCT(h);
  Solver& s = S(h);
  if (s.nranks > 1) { g_err = "gpba_get_hpl is a single-GPU parity accessor (landmarks are sharded)"; return GPBA_ERR_STATE; }
  CK(cudaStreamSynchronize(s.stream));
  std::vector<double> W((size_t)s.n_aobs * 18), R((size_t)s.n_rec * GPBA_REC_STRIDE);
  std::vector<int> orec((size_t)s.n_aobs);
  if (s.n_aobs) {
    CK(cudaMemcpyAsync(W.data(), s.d_W.p, W.size() * 8, cudaMemcpyDeviceToHost, s.stream));
    CK(cudaMemcpyAsync(orec.data(), s.d_o_rec.p, orec.size() * 4, cudaMemcpyDeviceToHost, s.stream));
  }
  if (s.n_rec) CK(cudaMemcpyAsync(R.data(), s.d_rec.p, R.size() * 8, cudaMemcpyDeviceToHost, s.stream));
  CK(cudaStreamSynchronize(s.stream));
  std::vector<int> inv(s.n_lm);
  for (int l = 0; l < s.n_lm; ++l) inv[s.lm_rank[l]] = l;
  int64_t cursor = 0;
  std::vector<int> poses;
  for (int g = 0; g < s.n_lm; ++g) {
    const int l = inv[g];
    if (lm_begin) lm_begin[g] = cursor;
    poses.clear();
    for (int64_t j = s.lm_obs_begin[l]; j < s.lm_obs_begin[l + 1]; ++j) {
      const int r = orec[j];
      if (s.rec_kf1[r] >= 0 && s.kf_h[s.rec_kf1[r]] >= 0) poses.push_back(s.kf_h[s.rec_kf1[r]]);
      if (s.kf_h[s.rec_kf2[r]] >= 0) poses.push_back(s.kf_h[s.rec_kf2[r]]);
      if (s.rec_kf1[r] >= 0 && s.ext_h[s.rec_cam[r]] >= 0) poses.push_back(s.ext_h[s.rec_cam[r]]);
    }
    std::sort(poses.begin(), poses.end());
    poses.erase(std::unique(poses.begin(), poses.end()), poses.end());
    for (size_t a = 0; a < poses.size(); ++a) {
      if (pose) pose[cursor + a] = poses[a];
      if (blocks) std::fill(blocks + (cursor + a) * 36, blocks + (cursor + a + 1) * 36, 0.0);
    }
    if (blocks)
      for (int64_t j = s.lm_obs_begin[l]; j < s.lm_obs_begin[l + 1]; ++j) {
        const int r = orec[j];
        const double* M = R.data() + (size_t)r * GPBA_REC_STRIDE + GPBA_REC_M;
        for (int which = 0; which < 3; ++which) {
          const int k = which == 1 ? s.rec_kf2[r] : s.rec_kf1[r];
          const int hh = which == 2 ? (k >= 0 ? s.ext_h[s.rec_cam[r]] : -1) : (k >= 0 ? s.kf_h[k] : -1);
          if (hh < 0) continue;
          const size_t slot = cursor + (std::lower_bound(poses.begin(), poses.end(), hh) - poses.begin());
          for (int c12 = 0; c12 < 12; ++c12)
            for (int c = 0; c < 3; ++c) {
              double acc = 0.0;
              for (int m = 0; m < 6; ++m) acc += M[m * GPBA_REC_MS + 12 * which + c12] * W[(size_t)j * 18 + m * 3 + c];
              blocks[slot * 36 + c12 * 3 + c] += acc;
            }
        }
      }
    cursor += (int64_t)poses.size();
  }
  if (lm_begin) lm_begin[s.n_lm] = cursor;
  return GPBA_OK;
}
int gpba_oplus(gpba_handle* h, const double* x) {
  NEED_STRUCT(h);
  Solver& s = S(h);
  if (x) {  // caller-provided update in g2o order
    CK(cudaMemcpyAsync(s.d_x.p, x, sizeof(double) * 12 * (size_t)s.n_pose, cudaMemcpyHostToDevice, s.stream));
    std::vector<double> xl((size_t)s.n_lm * 3);
    for (int l = 0; l < s.n_lm; ++l) for (int c = 0; c < 3; ++c) xl[(size_t)l * 3 + c] = x[(size_t)s.n_pose * 12 + (size_t)s.lm_rank[l] * 3 + c];
    if (s.n_lm) CK(cudaMemcpyAsync(s.d_xl.p, xl.data(), xl.size() * 8, cudaMemcpyHostToDevice, s.stream));
    // landmarks: plain addition of the provided update
    std::vector<double> pt((size_t)s.n_lm * 3);
    if (s.n_lm) {
      CK(cudaMemcpyAsync(pt.data(), s.d_ptS[s.cur].p, pt.size() * 8, cudaMemcpyDeviceToHost, s.stream));
      CK(cudaStreamSynchronize(s.stream));
      for (size_t i = 0; i < pt.size(); ++i) pt[i] += xl[i];
      CK(cudaMemcpyAsync(s.d_ptS[1 - s.cur].p, pt.data(), pt.size() * 8, cudaMemcpyHostToDevice, s.stream));
    }
    k_update_poses<<<(s.n_kf + 63) / 64, 64, 0, s.stream>>>(s.V, 0.0, s.d_x.p, s.d_bp.p, s.d_pose[s.cur].p, s.d_vel[s.cur].p,
                                                           s.d_pose[1 - s.cur].p, s.d_vel[1 - s.cur].p, s.d_pose_scale.p);
    CK(cudaGetLastError());
    if (s.n_pose > s.n_pose_kf) {
      k_update_ext<<<(s.n_cam + 31) / 32, 32, 0, s.stream>>>(s.V, 0.0, s.d_x.p, s.d_bp.p, s.d_ext[s.cur].p, s.d_ext[1 - s.cur].p,
                                                            s.d_cam[s.cur].p, s.d_cam[1 - s.cur].p, s.d_pose_scale.p);
      CK(cudaGetLastError());
    }
    CK(cudaStreamSynchronize(s.stream));
  }
  // x == NULL: gpba_solve already wrote state (+) x into the other buffer
  s.cur = 1 - s.cur;
  return GPBA_OK;
}
int gpba_push(gpba_handle* h) {
  NEED_STRUCT(h);
  Solver& s = S(h);
  std::vector<double> p((size_t)s.n_kf * 7), v((size_t)s.n_kf * 6), q((size_t)s.n_lm * 3), x((size_t)s.n_cam * 7);
  CK(cudaStreamSynchronize(s.stream));
  CK(cudaMemcpyAsync(x.data(), s.d_ext[s.cur].p, x.size() * 8, cudaMemcpyDeviceToHost, s.stream));
  CK(cudaMemcpyAsync(p.data(), s.d_pose[s.cur].p, p.size() * 8, cudaMemcpyDeviceToHost, s.stream));
  CK(cudaMemcpyAsync(v.data(), s.d_vel[s.cur].p, v.size() * 8, cudaMemcpyDeviceToHost, s.stream));
  if (s.n_lm) CK(cudaMemcpyAsync(q.data(), s.d_ptS[s.cur].p, q.size() * 8, cudaMemcpyDeviceToHost, s.stream));
  CK(cudaStreamSynchronize(s.stream));
  s.stack_pose.push_back(p); s.stack_vel.push_back(v); s.stack_pt.push_back(q); s.stack_ext.push_back(x);
  return GPBA_OK;
}
int gpba_pop(gpba_handle* h) {
  NEED_STRUCT(h);
  Solver& s = S(h);
  if (s.stack_pose.empty()) { g_err = "pop on empty stack"; return GPBA_ERR_STATE; }
  for (int b = 0; b < 2; ++b) {
    CK(cudaMemcpyAsync(s.d_pose[b].p, s.stack_pose.back().data(), s.stack_pose.back().size() * 8, cudaMemcpyHostToDevice, s.stream));
    CK(cudaMemcpyAsync(s.d_vel[b].p, s.stack_vel.back().data(), s.stack_vel.back().size() * 8, cudaMemcpyHostToDevice, s.stream));
    if (s.n_lm) CK(cudaMemcpyAsync(s.d_ptS[b].p, s.stack_pt.back().data(), s.stack_pt.back().size() * 8, cudaMemcpyHostToDevice, s.stream));
    CK(cudaMemcpyAsync(s.d_ext[b].p, s.stack_ext.back().data(), s.stack_ext.back().size() * 8, cudaMemcpyHostToDevice, s.stream));
  }
  CKR(s.refresh_cams());
  CK(cudaStreamSynchronize(s.stream));
  s.stack_pose.pop_back(); s.stack_vel.pop_back(); s.stack_pt.pop_back(); s.stack_ext.pop_back();
  return GPBA_OK;
}
int gpba_discard_top(gpba_handle* h) {
  NEED_STRUCT(h);
  Solver& s = S(h);
  if (s.stack_pose.empty()) { g_err = "discardTop on empty stack"; return GPBA_ERR_STATE; }
  s.stack_pose.pop_back(); s.stack_vel.pop_back(); s.stack_pt.pop_back(); s.stack_ext.pop_back();
  return GPBA_OK;
}

int gpba_optimize(gpba_handle* h, int iters, const volatile unsigned char* stop_flag, const gpba_lm_params* params, gpba_lm_trace* trace) {
  NEED(h);
  gpba_lm_params d;
  gpba_default_lm_params(&d);
  if (params) d = *params;
  S(h).pcg.tolerance = d.pcg_tolerance; S(h).pcg.max_iterations = d.pcg_max_iterations;
  return S(h).optimize(iters, stop_flag, d, trace);
}
int gpba_download_state(gpba_handle* h, double* kf_pose, double* kf_vel, double* pt_xyz) { NEED(h); return S(h).download_state(kf_pose, kf_vel, pt_xyz); }
int gpba_edge_chi2(gpba_handle* h, double* chi2) {
  NEED(h);
  Solver& s = S(h);
  CK(cudaSetDevice(s.device));
  CK(cudaStreamSynchronize(s.stream));
  if (s.n_obs) CK(cudaMemcpyAsync(chi2, s.d_chi2.p, sizeof(double) * (size_t)s.n_obs, cudaMemcpyDeviceToHost, s.stream));
  CK(cudaStreamSynchronize(s.stream));
  return GPBA_OK;
}

// BaseEdge::_error of every ACTIVE reprojection edge as the last evaluation left it (the stale-error quirk: after a
// rejected last trial these are the rejected state's errors, which is what the reference's e->chi2() reads afterwards);
// entries of inactive (level 1) edges are NaN = "unchanged".  Recomputed on demand from the state buffer of that evaluation.
int gpba_edge_errors(gpba_handle* h, double* err3) {
  NEED_STRUCT(h);
  Solver& s = S(h);
  if (!err3) { g_err = "null argument"; return GPBA_ERR_INVALID; }
  if (s.nranks > 1) { g_err = "gpba_edge_errors is a single-GPU accessor"; return GPBA_ERR_STATE; }
  if (s.n_obs == 0) return GPBA_OK;
  DBuf<double> d_err;
  CKR(d_err.alloc(3 * (size_t)s.n_obs));
  CK(cudaMemsetAsync(d_err.p, 0xFF, sizeof(double) * 3 * (size_t)s.n_obs, s.stream));   // all-ones = NaN
  const int ev = s.last_eval;
  CKR(s.compute_records(ev, false));
  if (s.n_aobs > 0) {
    if (s.stereo) k_residual<true><<<s.grid_obs, 256, 0, s.stream>>>(s.Vb(ev), s.d_rec_lite.p, s.d_ptS[ev].p, s.d_partial.p, nullptr, d_err.p);
    else k_residual<false><<<s.grid_obs, 256, 0, s.stream>>>(s.Vb(ev), s.d_rec_lite.p, s.d_ptS[ev].p, s.d_partial.p, nullptr, d_err.p);
    CK(cudaGetLastError());
  }
  CK(cudaMemcpyAsync(err3, d_err.p, sizeof(double) * 3 * (size_t)s.n_obs, cudaMemcpyDeviceToHost, s.stream));
  CK(cudaStreamSynchronize(s.stream));
  return GPBA_OK;
}
// Keyframe states and extrinsics of the last EVALUATED state (equal to gpba_download_state's unless the last trial was
// rejected): lets the caller recompute the errors of its few prior edges exactly where the reference left them.
int gpba_download_evaluated_state(gpba_handle* h, double* kf_pose, double* kf_vel, double* cam_Tbc) {
  NEED(h);
  Solver& s = S(h);
  CK(cudaSetDevice(s.device));
  const int ev = s.last_eval;
  if (kf_pose) CK(cudaMemcpyAsync(kf_pose, s.d_pose[ev].p, sizeof(double) * 7 * (size_t)s.n_kf, cudaMemcpyDeviceToHost, s.stream));
  if (kf_vel) CK(cudaMemcpyAsync(kf_vel, s.d_vel[ev].p, sizeof(double) * 6 * (size_t)s.n_kf, cudaMemcpyDeviceToHost, s.stream));
  if (cam_Tbc) CK(cudaMemcpyAsync(cam_Tbc, s.d_ext[ev].p, sizeof(double) * 7 * (size_t)s.n_cam, cudaMemcpyDeviceToHost, s.stream));
  CK(cudaStreamSynchronize(s.stream));
  return GPBA_OK;
}

__global__ void k_robust_sum(DevView V, int64_t n_obs, const double* __restrict__ chi2, const uint8_t* __restrict__ flags,
                             const double* __restrict__ ur, double* __restrict__ partial) {
  __shared__ double red[32];
  double acc = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_obs; i += (int64_t)gridDim.x * blockDim.x) {
    if (flags[i] & 0x2u) continue;
    const bool st = ur && ur[i] >= 0.0;
    const double delta = st ? V.hub_stereo_delta : V.hub_mono_delta, dsqr = st ? V.hub_stereo_dsqr : V.hub_mono_dsqr;
    double r1;
    acc += (delta > 0.0 && !(flags[i] & 0x4u)) ? huber(chi2[i], delta, dsqr, &r1) : chi2[i];
  }
  const double s = block_sum(acc, red);
  if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

int gpba_active_robust_chi2(gpba_handle* h, double* chi2) {
  NEED_STRUCT(h);
  Solver& s = S(h);
  const int g = (int)std::min<int64_t>((s.n_obs + 255) / 256 + 1, 148 * 8);
  k_robust_sum<<<g, 256, 0, s.stream>>>(s.V, s.n_obs, s.d_chi2.p, s.d_all_flags.p, s.stereo ? s.d_all_ur.p : nullptr, s.d_partial.p);
  CK(cudaGetLastError());
  const int np = s.n_prior + (s.n_velp + 63) / 64;
  if (np > 0) { k_priors<<<np, 64, 0, s.stream>>>(s.V, s.d_pose[s.last_eval].p, s.d_vel[s.last_eval].p, 0, s.d_prior_rho.p, nullptr, nullptr); CK(cudaGetLastError()); }
  k_ext_prior<<<(s.n_cam + 31) / 32, 32, 0, s.stream>>>(s.V, s.d_ext[s.last_eval].p, 0, s.d_prior_rho.p + s.n_prior + s.n_velp, nullptr, nullptr);
  CK(cudaGetLastError());
  k_reduce<<<1, 256, 0, s.stream>>>(s.d_partial.p, g, s.d_prior_rho.p, s.n_prior + s.n_velp + s.n_cam, s.d_scal.p + 2);
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(s.h_scal + 2, s.d_scal.p + 2, 8, cudaMemcpyDeviceToHost, s.stream));
  CK(cudaStreamSynchronize(s.stream));
  *chi2 = s.h_scal[2];
  return GPBA_OK;
}

int gpba_outlier_flags(gpba_handle* h, const gpba_thresholds* th, uint8_t* flags) {
  NEED_STRUCT(h);
  Solver& s = S(h);
  if (!th || !flags) { g_err = "null argument"; return GPBA_ERR_INVALID; }
  CKR(s.scatter_points(s.cur));
  DBuf<uint8_t> d_flags;
  CKR(d_flags.alloc((size_t)s.n_obs));
  const int g = (int)std::min<int64_t>((s.n_obs + 255) / 256 + 1, 148 * 8);
  k_flags<<<g, 256, 0, s.stream>>>(s.Vb(s.cur), s.n_obs, s.d_chi2.p, s.stereo ? s.d_all_ur.p : nullptr, s.d_all_rec.p, s.d_all_flags.p,
                                   s.d_all_pt.p, s.d_pt_full.p, s.d_pose[s.cur].p, th->chi2_mono, th->chi2_mono_close, th->chi2_stereo, d_flags.p);
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(flags, d_flags.p, (size_t)s.n_obs, cudaMemcpyDeviceToHost, s.stream));
  CK(cudaStreamSynchronize(s.stream));
  return GPBA_OK;
}

int gpba_set_levels(gpba_handle* h, const uint8_t* level) {
  NEED(h);
  Solver& s = S(h);
  for (int64_t i = 0; i < s.n_obs; ++i) { if (level[i]) s.obs_flags[i] |= GPBA_OBS_LEVEL1; else s.obs_flags[i] &= ~GPBA_OBS_LEVEL1; }
  s.structure_dirty = true;
  return GPBA_OK;  // takes effect at the next build_structure / optimize (initializeOptimization)
}
int gpba_set_robust_kernel(gpba_handle* h, int enabled) {
  NEED(h);
  Solver& s = S(h);
  for (int64_t i = 0; i < s.n_obs; ++i) { if (!enabled) s.obs_flags[i] |= GPBA_OBS_NO_KERNEL; else s.obs_flags[i] &= ~GPBA_OBS_NO_KERNEL; }
  s.structure_dirty = true;
  return GPBA_OK;
}

// chi2 of the currently inactive (level 1) edges at the current estimate: "if (mvbOutlier[idx]) e->computeError()" (Optimizer.cc:591-592)
__global__ void k_chi2_inactive(DevView V, int64_t n_obs, const uint8_t* __restrict__ flags, const int* __restrict__ obs_rec,
                                const int* __restrict__ obs_pt, const double* __restrict__ u, const double* __restrict__ v,
                                const double* __restrict__ ur, const double* __restrict__ w, const double* __restrict__ rec_lite,
                                const double* __restrict__ pt_full, double* __restrict__ chi2) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_obs; i += (int64_t)gridDim.x * blockDim.x) {
    if (!(flags[i] & 0x2u)) continue;
    const int r = obs_rec[i];
    const size_t p = (size_t)obs_pt[i];
    ObsEval<true> E;
    eval_obs<true, false>(V, rec_lite + (size_t)r * GPBA_REC_LITE_STRIDE, V.cam[V.rec_cam[r]], pt_full[3 * p], pt_full[3 * p + 1],
                          pt_full[3 * p + 2], u[i], v[i], ur ? ur[i] : -1.0, w[i], flags[i], E, nullptr, nullptr);
    chi2[i] = E.chi2;
  }
}

int gpba_compute_errors_inactive(gpba_handle* h) {
  NEED_STRUCT(h);
  Solver& s = S(h);
  CKR(s.scatter_points(s.cur));
  CKR(s.compute_records(s.cur, false));
  const int g = (int)std::min<int64_t>((s.n_obs + 255) / 256 + 1, 148 * 8);
  k_chi2_inactive<<<g, 256, 0, s.stream>>>(s.Vb(s.cur), s.n_obs, s.d_all_flags.p, s.d_all_rec.p, s.d_all_pt.p, s.d_all_u.p, s.d_all_v.p,
                                           s.stereo ? s.d_all_ur.p : nullptr, s.d_all_w.p, s.d_rec_lite.p, s.d_pt_full.p, s.d_chi2.p);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(s.stream));
  return GPBA_OK;
}

int gpba_rejection_rounds(gpba_handle* h, int n_rounds, int iters, const gpba_thresholds* th, const gpba_lm_params* params,
                          uint8_t* flags_out, gpba_lm_trace* traces) {
  NEED(h);
  Solver& s = S(h);
  std::vector<uint8_t> fl((size_t)s.n_obs, 0);
  for (int it = 0; it < n_rounds; ++it) {
    CKR(gpba_optimize(h, iters, nullptr, params, traces ? &traces[it] : nullptr));
    CKR(gpba_compute_errors_inactive(h));
    CKR(gpba_outlier_flags(h, th, fl.data()));
    CKR(gpba_set_levels(h, fl.data()));
    if (it == 2) CKR(gpba_set_robust_kernel(h, 0));
    CKR(s.d_all_flags.upload(s.obs_flags, s.stream));
  }
  if (flags_out) std::memcpy(flags_out, fl.data(), fl.size());
  return GPBA_OK;
}

// ---- extrinsic self-calibration (LocalGPBA's second stage, src/Optimizer.cc:983-995, 1228-1240, 1419-1428)
int gpba_set_extrinsics(gpba_handle* h, const gpba_extrinsics* e) {
  NEED(h);
  Solver& s = S(h);
  if (!e || !e->free_mask) { g_err = "null argument"; return GPBA_ERR_INVALID; }
  if ((e->prior_R == nullptr) != (e->prior_info == nullptr)) { g_err = "prior_R and prior_info go together"; return GPBA_ERR_INVALID; }
  bool any = false;
  for (int c = 0; c < s.n_cam; ++c) any = any || e->free_mask[c];
  if (any && s.nranks > 1) { g_err = "extrinsic self-calibration belongs to the local BA: single GPU only"; return GPBA_ERR_INVALID; }
  for (int c = 0; c < s.n_cam; ++c) {
    s.ext_free[c] = e->free_mask[c] ? 1 : 0;
    s.ext_prior_on[c] = e->prior_R ? 1 : 0;
    if (e->prior_R) {
      Quat q; q.x = e->prior_R[4 * c]; q.y = e->prior_R[4 * c + 1]; q.z = e->prior_R[4 * c + 2]; q.w = e->prior_R[4 * c + 3];
      const Quat qi = quat_inv(q);   // EdgeExtrinsicPrior keeps R_ini^-1 (G2oTypes.h:474)
      s.ext_prior_qinv[4 * c] = qi.x; s.ext_prior_qinv[4 * c + 1] = qi.y; s.ext_prior_qinv[4 * c + 2] = qi.z; s.ext_prior_qinv[4 * c + 3] = qi.w;
      for (int k = 0; k < 9; ++k) s.ext_prior_info[9 * c + k] = e->prior_info[9 * c + k];
    }
  }
  s.structure_dirty = true;   // takes effect at the next build_structure / optimize (initializeOptimization, :1238)
  return GPBA_OK;
}
int gpba_get_extrinsics(gpba_handle* h, double* cam_Tbc) {
  NEED(h);
  Solver& s = S(h);
  if (!cam_Tbc) { g_err = "null argument"; return GPBA_ERR_INVALID; }
  CK(cudaSetDevice(s.device));
  CK(cudaMemcpyAsync(cam_Tbc, s.d_ext[s.cur].p, sizeof(double) * 7 * (size_t)s.n_cam, cudaMemcpyDeviceToHost, s.stream));
  CK(cudaStreamSynchronize(s.stream));
  return GPBA_OK;
}
__global__ void k_count_cam_obs(int64_t n_obs, const int* __restrict__ obs_rec, const int* __restrict__ rec_kf1, const int* __restrict__ rec_cam,
                                unsigned long long* __restrict__ cnt) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_obs; i += (int64_t)gridDim.x * blockDim.x) {
    const int r = obs_rec[i];
    if (rec_kf1[r] >= 0) atomicAdd(&cnt[rec_cam[r]], 1ull);
  }
}
int gpba_count_camera_observations(gpba_handle* h, int64_t* cam_obs) {
  NEED(h);
  Solver& s = S(h);
  if (!cam_obs) { g_err = "null argument"; return GPBA_ERR_INVALID; }
  CK(cudaSetDevice(s.device));
  DBuf<unsigned long long> d_cnt;
  CKR(d_cnt.alloc((size_t)s.n_cam));
  CK(cudaMemsetAsync(d_cnt.p, 0, sizeof(unsigned long long) * (size_t)s.n_cam, s.stream));
  if (s.n_obs > 0) {
    k_count_cam_obs<<<(int)std::min<int64_t>((s.n_obs + 255) / 256, 148 * 8), 256, 0, s.stream>>>(s.n_obs, s.d_all_rec.p, s.d_rec_kf1.p, s.d_rec_cam.p, d_cnt.p);
    CK(cudaGetLastError());
  }
  std::vector<unsigned long long> c((size_t)s.n_cam);
  CK(cudaMemcpyAsync(c.data(), d_cnt.p, sizeof(unsigned long long) * (size_t)s.n_cam, cudaMemcpyDeviceToHost, s.stream));
  CK(cudaStreamSynchronize(s.stream));
  for (int i = 0; i < s.n_cam; ++i) cam_obs[i] = (int64_t)c[i];
  return GPBA_OK;
}
int gpba_calibrate_extrinsics(gpba_handle* h, const gpba_extrinsics* candidates, int min_obs, int iters, const gpba_lm_params* params,
                              gpba_lm_trace* trace, uint8_t* freed_out) {
  NEED(h);
  Solver& s = S(h);
  if (!candidates || !candidates->free_mask) { g_err = "null argument"; return GPBA_ERR_INVALID; }
  std::vector<int64_t> cam_obs((size_t)s.n_cam);
  CKR(gpba_count_camera_observations(h, cam_obs.data()));
  std::vector<uint8_t> fr((size_t)s.n_cam, 0);
  for (int c = 0; c < s.n_cam; ++c) fr[c] = candidates->free_mask[c] && cam_obs[c] >= min_obs;   // "if (cam_obs[i] < extrin_thresh) continue" (:1232)
  gpba_extrinsics e = *candidates;
  e.free_mask = fr.data();
  CKR(gpba_set_extrinsics(h, &e));
  if (freed_out) std::memcpy(freed_out, fr.data(), fr.size());
  return gpba_optimize(h, iters, nullptr, params, trace);   // initializeOptimization + computeActiveErrors + optimize(opt_it2) (:1238-1240)
}

int gpba_stage_stats(gpba_handle* h, double ms_total[GPBA_N_STAGES], int64_t launches[GPBA_N_STAGES], int reset) {
  NEED(h);
  S(h).collect_events();
  for (int i = 0; i < GPBA_N_STAGES; ++i) {
    if (ms_total) ms_total[i] = S(h).stage_ms[i];
    if (launches) launches[i] = S(h).stage_launches[i];
    if (reset) { S(h).stage_ms[i] = 0; S(h).stage_launches[i] = 0; }
  }
  return GPBA_OK;
}
int gpba_schur_stats(gpba_handle* h, int64_t out[4]) {
  NEED_STRUCT(h);
  out[0] = S(h).n_pairs; out[1] = S(h).n_rp; out[2] = S(h).n_items; out[3] = S(h).n_con;
  return GPBA_OK;
}
int gpba_solver_stats(gpba_handle* h, int64_t out[6]) {
  NEED_STRUCT(h);
  out[0] = S(h).NT; out[1] = (int64_t)S(h).lvl_ncols.size(); out[2] = S(h).chol_parts; out[3] = S(h).chol_doubles / GPBA_TILE;
  out[4] = S(h).chol_products; out[5] = S(h).chol_update_ctas;
  return GPBA_OK;
}
int gpba_symbolic_analyze(int32_t n_pose, int32_t n_hs, const int32_t* hs_row, const int32_t* hs_col, int32_t nd_depth,
                          int32_t* perm_out, int64_t out[5]) {
  if (n_pose < 0 || n_hs < 0 || (n_hs > 0 && (!hs_row || !hs_col)) || !out) { g_err = "invalid argument"; return GPBA_ERR_INVALID; }
  for (int k = 0; k < n_hs; ++k)
    if (hs_row[k] < 0 || hs_col[k] < hs_row[k] || hs_col[k] >= n_pose) { g_err = "block index out of range (upper pattern expected)"; return GPBA_ERR_INVALID; }
  CholSymbolic sym;
  chol_symbolic(n_pose, n_hs, hs_row, hs_col, GPBA_NB / 12, GPBA_TILE, nd_depth, false, sym);
  int64_t pairs = 0;
  for (int k = 0; k < sym.NT; ++k) { const int64_t nr = sym.col_begin[k + 1] - sym.col_begin[k]; pairs += nr * (nr + 1) / 2; }
  out[0] = sym.NT; out[1] = sym.n_levels; out[2] = sym.n_parts; out[3] = sym.doubles / GPBA_TILE; out[4] = pairs;
  if (perm_out) for (int i = 0; i < n_pose; ++i) perm_out[i] = sym.perm[i];
  return GPBA_OK;
}
// Host only: build the task list of the persistent factorization for a block pattern and check the invariants its
// in-kernel waits rely on.  out[0] tasks, out[1] update chunks, out[2] tile products, out[3] panel tasks,
// out[4] VIOLATIONS (must be 0):
//   * every product of a chunk reads a source column whose panel tasks ALL sit in front of the chunk in the list;
//   * every panel task sits behind all chunks that target its column, a solve-only task behind its column's publishing task;
//   * the completion counts the kernel waits for (chunks per column, panel tasks per column) equal what the list holds;
//   * every tile referenced by a task or a product exists, every below-diagonal tile has exactly one panel task.
int gpba_factor_schedule_check(int32_t n_pose, int32_t n_hs, const int32_t* hs_row, const int32_t* hs_col, int64_t out[5]) {
  if (n_pose < 0 || n_hs < 0 || (n_hs > 0 && (!hs_row || !hs_col)) || !out) { g_err = "invalid argument"; return GPBA_ERR_INVALID; }
  for (int k = 0; k < n_hs; ++k)
    if (hs_row[k] < 0 || hs_col[k] < hs_row[k] || hs_col[k] >= n_pose) { g_err = "block index out of range (upper pattern expected)"; return GPBA_ERR_INVALID; }
  CholHost H;
  const int rc = chol_host_phase(n_pose, n_hs, hs_row, hs_col, H);
  if (rc != GPBA_OK) { g_err = H.err; return rc; }
  const int NT = H.sym.NT;
  const int64_t n_tiles = H.sym.doubles / GPBA_TILE;
  int64_t bad = 0, chunks = 0, products = 0, panels = 0;
  std::vector<int> pan_seen(NT, 0), upd_seen(NT, 0), diag_pos(NT, -1), publishes(NT, 0);
  std::vector<char> tile_has_task((size_t)std::max<int64_t>(n_tiles, 1), 0);
  // first pass: position of the last panel task of every column, number of chunks per column
  std::vector<int> last_panel(NT, -1), last_chunk(NT, -1);
  for (size_t t = 0; t < H.cf_tab.size(); ++t) {
    const int4 e = H.cf_tab[t];
    const int j = e.x & ~GPBA_CF_DIAG;
    if (j < 0 || j >= NT || e.y < 0 || e.y >= n_tiles) { ++bad; continue; }
    if (e.z < 0) last_panel[j] = (int)t; else last_chunk[j] = (int)t;
  }
  for (size_t t = 0; t < H.cf_tab.size(); ++t) {
    const int4 e = H.cf_tab[t];
    const int j = e.x & ~GPBA_CF_DIAG;
    if (j < 0 || j >= NT || e.y < 0 || e.y >= n_tiles) continue;
    const bool diag = (e.x & GPBA_CF_DIAG) != 0;
    if (e.z >= 0) {
      ++chunks; ++upd_seen[j];
      if (e.w <= e.z || e.w > (int)H.prod.size()) { ++bad; continue; }
      if (diag != (H.sym.tile_off[(size_t)j * NT + j] / GPBA_TILE == e.y)) ++bad;
      for (int p = e.z; p < e.w; ++p) {
        ++products;
        const int4 r = H.prod[p];
        if (r.z < 0 || r.z >= j || r.x < 0 || r.x >= n_tiles || r.y < 0 || r.y >= n_tiles) { ++bad; continue; }
        if (last_panel[r.z] < 0 || last_panel[r.z] > (int)t) ++bad;                       // the source column is complete in front of the chunk
        if (H.sym.tile_off[(size_t)j * NT + r.z] / GPBA_TILE != r.y) ++bad;                // second operand: tile (j, k)
      }
    } else {
      ++panels; ++pan_seen[j];
      if (last_chunk[j] > (int)t) ++bad;                                                   // all updates of the column are in front of its panel tasks
      if (e.w != H.sym.tile_off[(size_t)j * NT + j] / GPBA_TILE) ++bad;                    // the diagonal tile the task factorizes / reads
      if (diag) { if (diag_pos[j] >= 0) ++bad; diag_pos[j] = (int)t; if (e.z == -2) publishes[j] = 1; if (e.z == -3) ++bad; }
      else {
        if (tile_has_task[e.y]) ++bad;
        tile_has_task[e.y] = 1;
        if (e.z == -3 && !(publishes[j] && diag_pos[j] >= 0 && diag_pos[j] < (int)t)) ++bad;   // solve-only: behind the publishing task
        if (e.z == -1 && publishes[j]) ++bad;                                               // one mode per column
      }
    }
  }
  for (int j = 0; j < NT; ++j) {
    const int nr = H.sym.col_begin[j + 1] - H.sym.col_begin[j];
    if (pan_seen[j] != nr + 1 || H.cf_need[(size_t)NT + j] != nr + 1) ++bad;
    if (upd_seen[j] != H.cf_need[j]) ++bad;
    if (diag_pos[j] < 0) ++bad;
  }
  if (products != H.products) ++bad;
  out[0] = (int64_t)H.cf_tab.size(); out[1] = chunks; out[2] = products; out[3] = panels; out[4] = bad;
  return GPBA_OK;
}
int gpba_set_profiling(gpba_handle* h, int enabled) { NEED(h); S(h).collect_events(); S(h).profiling = enabled != 0; return GPBA_OK; }
void* gpba_get_stream(gpba_handle* h) { return h ? (void*)S(h).stream : nullptr; }

int gpba_reset_state(gpba_handle* h, const double* kf_pose, const double* kf_vel, const double* pt_xyz) {
  NEED(h);
  Solver& s = S(h);
  CK(cudaSetDevice(s.device));
  if (kf_pose) for (int b = 0; b < 2; ++b) CK(cudaMemcpyAsync(s.d_pose[b].p, kf_pose, sizeof(double) * 7 * (size_t)s.n_kf, cudaMemcpyHostToDevice, s.stream));
  if (kf_vel) for (int b = 0; b < 2; ++b) CK(cudaMemcpyAsync(s.d_vel[b].p, kf_vel, sizeof(double) * 6 * (size_t)s.n_kf, cudaMemcpyHostToDevice, s.stream));
  if (s.any_ext_free()) {   // the extrinsics are part of the state: back to the ones the handle was created with
    for (int b = 0; b < 2; ++b) {
      CK(cudaMemcpyAsync(s.d_ext[b].p, s.h_ext0.data(), sizeof(double) * 7 * (size_t)s.n_cam, cudaMemcpyHostToDevice, s.stream));
      CK(cudaMemcpyAsync(s.d_cam[b].p, s.h_cam0.data(), sizeof(CamConst) * (size_t)s.n_cam, cudaMemcpyHostToDevice, s.stream));
    }
  }
  if (pt_xyz) {
    CK(cudaMemcpyAsync(s.d_pt_full.p, pt_xyz, sizeof(double) * 3 * (size_t)s.n_pt, cudaMemcpyHostToDevice, s.stream));
    if (s.structure_ok && s.n_lm > 0) {
      k_gather_pts<<<(s.n_lm * 3 + 255) / 256, 256, 0, s.stream>>>(s.n_lm, s.d_lm_pt.p, s.d_pt_full.p, s.d_ptS[0].p, s.d_ptS[1].p);
      CK(cudaGetLastError());
    }
  }
  CK(cudaStreamSynchronize(s.stream));
  return GPBA_OK;
}

// ---- pose-only GP optimisation (Optimizer::PoseGPOptimizationFromeLastFrame, src/Optimizer.cc:369-686)
int gpba_pose_optimize(const gpba_pose_batch* B, int device, double* cur_pose_out, double* cur_vel_out, double* prev_pose_out,
                       double* prev_vel_out, uint8_t* outlier_out, int32_t* n_inliers_out, gpba_lm_trace* traces) {
  if (!B || B->n_cam < 1 || B->n_cam > GPBA_POSE_MAX_CAM || B->n_frames < 0 || !B->cam_intr || !B->cam_Tbc) { g_err = "invalid pose batch"; return GPBA_ERR_INVALID; }
  const int nf = B->n_frames;
  if (nf == 0) return GPBA_OK;
  if (!B->prev_pose || !B->prev_vel || !B->prev_time || !B->prev_fixed || !B->cur_pose || !B->cur_vel || !B->cur_time || !B->cam_time || !B->obs_begin) { g_err = "invalid pose batch"; return GPBA_ERR_INVALID; }
  const int64_t n_obs = B->obs_begin[nf];
  for (int f = 0; f < nf; ++f) if (B->obs_begin[f + 1] < B->obs_begin[f] || B->obs_begin[f] < 0) { g_err = "obs_begin not monotone"; return GPBA_ERR_INVALID; }
  if (n_obs > 0 && (!B->obs_u || !B->obs_v || !B->obs_inv_sigma2 || !B->obs_xw || !B->obs_cam)) { g_err = "invalid pose batch"; return GPBA_ERR_INVALID; }
  for (int64_t i = 0; i < n_obs; ++i) if (B->obs_cam[i] < 0 || B->obs_cam[i] >= B->n_cam) { g_err = "camera index out of range"; return GPBA_ERR_INVALID; }
  for (int f = 0; f < nf; ++f) if (!(B->cur_time[f] > B->prev_time[f])) { g_err = "frame times not increasing"; return GPBA_ERR_INVALID; }
  {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { g_err = "no CUDA device (libgpba has no CPU fallback)"; return GPBA_ERR_NO_DEVICE; }
  }
  if (device < 0) CK(cudaGetDevice(&device));
  CK(cudaSetDevice(device));
  cudaStream_t st = small_call_stream(device);
  if (!st) { g_err = "cudaStreamCreate failed"; return GPBA_ERR_CUDA; }
  g_alloc_stream = st;
  {
    cudaMemPool_t pool; uint64_t thr = UINT64_MAX;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
  }
  std::vector<CamConst> cams(B->n_cam);
  for (int c = 0; c < B->n_cam; ++c) {
    CamConst& cc = cams[c];
    cc.fx = B->cam_intr[4 * c]; cc.fy = B->cam_intr[4 * c + 1]; cc.cx = B->cam_intr[4 * c + 2]; cc.cy = B->cam_intr[4 * c + 3];
    SE3 Tbc = load_se3(B->cam_Tbc + 7 * c);
    SE3 Tcb = se3_inv(Tbc);
    M3 Rcb = quat_to_R(Tcb.q), Rbc = quat_to_R(Tbc.q);
    for (int i = 0; i < 9; ++i) { cc.Rcb[i] = Rcb.a[i]; cc.Rbc[i] = Rbc.a[i]; }
    for (int i = 0; i < 3; ++i) { cc.tcb[i] = Tcb.t[i]; cc.tbc[i] = Tbc.t[i]; }
    cc.qbc[0] = Tbc.q.x; cc.qbc[1] = Tbc.q.y; cc.qbc[2] = Tbc.q.z; cc.qbc[3] = Tbc.q.w;
  }
  DBuf<CamConst> d_cam;
  DBuf<double> d_pp, d_pv, d_pt, d_cp, d_cv, d_ct, d_camt, d_u, d_v, d_ur, d_w, d_xw, d_chi2, d_ocp, d_ocv, d_opp, d_opv;
  DBuf<uint8_t> d_fix, d_fl, d_level, d_koff;
  DBuf<int64_t> d_ob;
  DBuf<int> d_cam_of, d_inl;
  DBuf<gpba_lm_trace> d_tr;
  CKR(d_cam.upload(cams, st));
  CKR(d_pp.upload(B->prev_pose, (size_t)7 * nf, st)); CKR(d_pv.upload(B->prev_vel, (size_t)6 * nf, st)); CKR(d_pt.upload(B->prev_time, nf, st));
  CKR(d_cp.upload(B->cur_pose, (size_t)7 * nf, st)); CKR(d_cv.upload(B->cur_vel, (size_t)6 * nf, st)); CKR(d_ct.upload(B->cur_time, nf, st));
  CKR(d_camt.upload(B->cam_time, (size_t)nf * B->n_cam, st)); CKR(d_fix.upload(B->prev_fixed, nf, st));
  CKR(d_ob.upload(B->obs_begin, (size_t)nf + 1, st));
  CKR(d_u.upload(B->obs_u, (size_t)n_obs, st)); CKR(d_v.upload(B->obs_v, (size_t)n_obs, st)); CKR(d_w.upload(B->obs_inv_sigma2, (size_t)n_obs, st));
  if (B->obs_ur) CKR(d_ur.upload(B->obs_ur, (size_t)n_obs, st));
  CKR(d_xw.upload(B->obs_xw, (size_t)3 * n_obs, st)); CKR(d_cam_of.upload(B->obs_cam, (size_t)n_obs, st));
  if (B->obs_flags) CKR(d_fl.upload(B->obs_flags, (size_t)n_obs, st));
  else { CKR(d_fl.alloc((size_t)n_obs)); CK(cudaMemsetAsync(d_fl.p, 0, (size_t)std::max<int64_t>(n_obs, 1), st)); }
  CKR(d_level.alloc((size_t)n_obs)); CKR(d_koff.alloc((size_t)n_obs)); CKR(d_chi2.alloc((size_t)n_obs));
  CKR(d_ocp.alloc((size_t)7 * nf)); CKR(d_ocv.alloc((size_t)6 * nf)); CKR(d_opp.alloc((size_t)7 * nf)); CKR(d_opv.alloc((size_t)6 * nf));
  CKR(d_inl.alloc(nf));
  if (traces) {   // rounds that are never run (edges().size() < 10) report an empty trace
    CKR(d_tr.alloc((size_t)nf * GPBA_POSE_ROUNDS));
    CK(cudaMemsetAsync(d_tr.p, 0, sizeof(gpba_lm_trace) * (size_t)nf * GPBA_POSE_ROUNDS, st));
  }

  DevView V;
  std::memset(&V, 0, sizeof(V));
  for (int i = 0; i < 6; ++i) V.qc_inv[i] = 1.0 / B->qc[i];
  V.bf = B->bf;
  V.hub_mono_delta = B->huber_mono; V.hub_mono_dsqr = f32sq(B->huber_mono);
  V.hub_stereo_delta = B->huber_stereo; V.hub_stereo_dsqr = f32sq(B->huber_stereo);
  PoseBatchView P;
  P.n_cam = B->n_cam; P.n_frames = nf; P.cam = d_cam.p;
  P.prev_pose = d_pp.p; P.prev_vel = d_pv.p; P.prev_time = d_pt.p; P.cur_pose = d_cp.p; P.cur_vel = d_cv.p; P.cur_time = d_ct.p;
  P.cam_time = d_camt.p; P.prev_fixed = d_fix.p; P.obs_begin = d_ob.p;
  P.obs_u = d_u.p; P.obs_v = d_v.p; P.obs_ur = B->obs_ur ? d_ur.p : nullptr; P.obs_w = d_w.p; P.obs_xw = d_xw.p; P.obs_cam = d_cam_of.p;
  P.obs_flags = d_fl.p; P.level = d_level.p; P.kernel_off = d_koff.p; P.chi2 = d_chi2.p;
  P.out_cur_pose = d_ocp.p; P.out_cur_vel = d_ocv.p; P.out_prev_pose = d_opp.p; P.out_prev_vel = d_opv.p; P.out_inliers = d_inl.p;
  P.traces = traces ? d_tr.p : nullptr;
  const bool verbose = getenv("GPBA_VERBOSE") != nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (verbose) { CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1)); CK(cudaEventRecord(e0, st)); }
  k_pose_only<<<nf, GPBA_POSE_THREADS, 0, st>>>(P, V);
  CK(cudaGetLastError());
  if (verbose) CK(cudaEventRecord(e1, st));
  if (cur_pose_out) CK(cudaMemcpyAsync(cur_pose_out, d_ocp.p, sizeof(double) * 7 * nf, cudaMemcpyDeviceToHost, st));
  if (cur_vel_out) CK(cudaMemcpyAsync(cur_vel_out, d_ocv.p, sizeof(double) * 6 * nf, cudaMemcpyDeviceToHost, st));
  if (prev_pose_out) CK(cudaMemcpyAsync(prev_pose_out, d_opp.p, sizeof(double) * 7 * nf, cudaMemcpyDeviceToHost, st));
  if (prev_vel_out) CK(cudaMemcpyAsync(prev_vel_out, d_opv.p, sizeof(double) * 6 * nf, cudaMemcpyDeviceToHost, st));
  if (outlier_out && n_obs) CK(cudaMemcpyAsync(outlier_out, d_level.p, (size_t)n_obs, cudaMemcpyDeviceToHost, st));
  if (n_inliers_out) CK(cudaMemcpyAsync(n_inliers_out, d_inl.p, sizeof(int) * nf, cudaMemcpyDeviceToHost, st));
  if (traces) CK(cudaMemcpyAsync(traces, d_tr.p, sizeof(gpba_lm_trace) * (size_t)nf * GPBA_POSE_ROUNDS, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  if (verbose) {
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    fprintf(stderr, "[gpba] pose-only: %d frames, %lld matches, kernel %.3f ms\n", nf, (long long)n_obs, ms);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
  }
  return GPBA_OK;
}

// ---- essential-graph optimisation (the solve inside Optimizer::OptimizeEssentialGraph, src/Optimizer.cc:1434-1717)
int gpba_pose_graph_optimize(const gpba_pose_graph* G, int device, int iters, const gpba_lm_params* params, double* sim3_out, gpba_lm_trace* trace) {
  if (!G || G->n_kf < 0 || G->n_edge < 0 || iters < 0 || (G->n_kf > 0 && (!G->sim3 || !G->fixed)) ||
      (G->n_edge > 0 && (!G->edge_i || !G->edge_j || !G->edge_meas))) { g_err = "invalid pose graph"; return GPBA_ERR_INVALID; }
  const int n = G->n_kf;
  const int64_t ne = G->n_edge;
  for (int64_t k = 0; k < ne; ++k)
    if (G->edge_i[k] < 0 || G->edge_i[k] >= n || G->edge_j[k] < 0 || G->edge_j[k] >= n || G->edge_i[k] == G->edge_j[k]) { g_err = "edge vertex index out of range"; return GPBA_ERR_INVALID; }
  if (trace) { std::memset(trace, 0, sizeof(*trace)); trace->result = GPBA_RESULT_OK; }
  if (n == 0) return GPBA_OK;
  {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { g_err = "no CUDA device (libgpba has no CPU fallback)"; return GPBA_ERR_NO_DEVICE; }
  }
  if (device < 0) CK(cudaGetDevice(&device));
  CK(cudaSetDevice(device));
  gpba_lm_params P;
  gpba_default_lm_params(&P);
  if (params) P = *params;
  // The reduced-system machinery (symbolic phase, tile Cholesky, its CUDA graphs) is the BA solver's: an otherwise empty
  // Solver serves as the factorization engine of this call.
  std::unique_ptr<gpba_handle> eng(new (std::nothrow) gpba_handle());
  if (!eng) return GPBA_ERR_INVALID;
  Solver& c = eng->s;
  c.device = device;
  CK(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
  c.stream_holder.s = c.stream;
  cudaStream_t st = c.stream;
  g_alloc_stream = st;
  // ---- structure (host: the graph is tiny next to a BA): active vertices, Hessian order, block pattern
  std::vector<char> act(n, 0);
  for (int64_t k = 0; k < ne; ++k) if (!(G->fixed[G->edge_i[k]] && G->fixed[G->edge_j[k]])) { act[G->edge_i[k]] = 1; act[G->edge_j[k]] = 1; }
  std::vector<int> h(n, -1);
  int np = 0;
  for (int i = 0; i < n; ++i) if (act[i] && !G->fixed[i]) h[i] = np++;
  if (np == 0 || iters == 0) {   // nothing to optimise
    if (sim3_out) std::memcpy(sim3_out, G->sim3, sizeof(double) * 8 * (size_t)n);
    return GPBA_OK;
  }
  std::vector<unsigned long long> keys;
  keys.reserve((size_t)np + (size_t)ne);
  for (int i = 0; i < np; ++i) keys.push_back(((unsigned long long)i << 32) | (unsigned)i);
  for (int64_t k = 0; k < ne; ++k) {
    const int a = h[G->edge_i[k]], b = h[G->edge_j[k]];
    if (a < 0 || b < 0) continue;
    keys.push_back(((unsigned long long)std::max(a, b) << 32) | (unsigned)std::min(a, b));   // (col, row), row <= col
  }
  std::sort(keys.begin(), keys.end());
  keys.erase(std::unique(keys.begin(), keys.end()), keys.end());
  const int nb = (int)keys.size();
  c.n_pose = np; c.n_pose_kf = np; c.n_hs = nb; c.linear_solver = GPBA_SOLVER_SPARSE_CHOL;
  c.hs_row.resize(nb); c.hs_col.resize(nb);
  std::vector<int> diag_of(nb, -1);
  for (int q = 0; q < nb; ++q) { c.hs_row[q] = (int)(unsigned)keys[q]; c.hs_col[q] = (int)(keys[q] >> 32); if (c.hs_row[q] == c.hs_col[q]) diag_of[q] = c.hs_row[q]; }
  auto blk = [&](int a, int b) { return (int)(std::lower_bound(keys.begin(), keys.end(), ((unsigned long long)b << 32) | (unsigned)a) - keys.begin()); };
  std::vector<int> bii((size_t)ne, -1), bjj((size_t)ne, -1), bij((size_t)ne, -1);
  for (int64_t k = 0; k < ne; ++k) {
    const int a = h[G->edge_i[k]], b = h[G->edge_j[k]];
    if (a >= 0) bii[k] = blk(a, a);
    if (b >= 0) bjj[k] = blk(b, b);
    if (a >= 0 && b >= 0) bij[k] = a < b ? blk(a, b) : (blk(b, a) | 0x40000000);
  }
  DBuf<double> d_S[2], d_meas, d_hpp, d_bp, d_partial, d_pscale, d_scal;
  DBuf<int> d_ei, d_ej, d_h, d_bii, d_bjj, d_bij, d_diag;
  DBuf<unsigned char> d_fixed;
  for (int b = 0; b < 2; ++b) CKR(d_S[b].upload(G->sim3, (size_t)8 * n, st));
  CKR(d_meas.upload(G->edge_meas, (size_t)8 * ne, st)); CKR(d_ei.upload(G->edge_i, (size_t)ne, st)); CKR(d_ej.upload(G->edge_j, (size_t)ne, st));
  CKR(d_fixed.upload(G->fixed, (size_t)n, st)); CKR(d_h.upload(h, st));
  CKR(d_bii.upload(bii, st)); CKR(d_bjj.upload(bjj, st)); CKR(d_bij.upload(bij, st)); CKR(d_diag.upload(diag_of, st));
  CKR(d_hpp.alloc((size_t)nb * 144)); CKR(d_bp.alloc((size_t)np * 12)); CKR(d_pscale.alloc((size_t)np)); CKR(d_scal.alloc(8));
  CKR(c.d_hs.alloc((size_t)nb * 144 + (size_t)np * 12 + 8)); CKR(c.d_x.alloc((size_t)np * 12)); CKR(c.d_fail.alloc(2));
  CKR(c.d_hs_row.upload(c.hs_row, st)); CKR(c.d_hs_col.upload(c.hs_col, st));
  CKR(c.build_cholesky_structure());
  CKR(c.capture_cholesky_graph());
  const int ge = (int)std::min<int64_t>((ne + 127) / 128, 148 * 8);
  CKR(d_partial.alloc((size_t)std::max(ge, 1)));
  PgView V;
  V.n_kf = n; V.fix_scale = G->fix_scale; V.n_edge = ne; V.ei = d_ei.p; V.ej = d_ej.p; V.meas = d_meas.p; V.fixed = d_fixed.p; V.h = d_h.p;
  V.blk_ii = d_bii.p; V.blk_jj = d_bjj.p; V.blk_ij = d_bij.p;
  double* h_scal = (double*)g_pinned.take();
  if (!h_scal) { g_err = "out of pinned scratch slots"; return GPBA_ERR_CUDA; }
  struct Give { double* p; ~Give() { g_pinned.give(p); } } give{h_scal};
  int cur = 0;
  auto errors = [&](int buf, double* chi) -> int {   // computeActiveErrors + activeRobustChi2
    k_pg_errors<<<std::max(ge, 1), 128, 0, st>>>(V, d_S[buf].p, d_partial.p);
    k_reduce<<<1, 256, 0, st>>>(d_partial.p, ne > 0 ? ge : 0, nullptr, 0, d_scal.p);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(h_scal, d_scal.p, 3 * sizeof(double), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(h_scal + 4, c.d_fail.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    *chi = h_scal[0];
    return GPBA_OK;
  };
  double lambda = -1, ni = 2;
  int nBad = 0, cj = 0, result = GPBA_RESULT_OK;
  bool ok = true;
  double* bs = c.d_hs.p + (size_t)nb * 144;
  for (int it = 0; it < iters && ok; ++it) {
    double currentChi = 0;
    CKR(errors(cur, &currentChi));
    double tempChi = currentChi;
    const double iniChi = currentChi;
    // buildSystem
    CK(cudaMemsetAsync(d_hpp.p, 0, sizeof(double) * 144 * (size_t)nb, st));
    CK(cudaMemsetAsync(d_bp.p, 0, sizeof(double) * 12 * (size_t)np, st));
    if (ne > 0) { k_pg_linearize<<<(int)((ne + 63) / 64), 64, 0, st>>>(V, d_S[cur].p, d_hpp.p, d_bp.p); CK(cudaGetLastError()); }
    if (it == 0) {
      if (G->lambda_init > 0) lambda = G->lambda_init;
      else {   // computeLambdaInit: tau * max |H_jj|
        std::vector<double> hd((size_t)np * 12);
        DBuf<double> d_diagv;
        CKR(d_diagv.alloc(hd.size()));
        DBuf<int> d_pd;
        std::vector<int> pd(np);
        for (int q = 0; q < nb; ++q) if (diag_of[q] >= 0) pd[diag_of[q]] = q;
        CKR(d_pd.upload(pd, st));
        k_hpp_diag<<<(np * 12 + 255) / 256, 256, 0, st>>>(np, d_pd.p, d_hpp.p, d_diagv.p);
        CK(cudaMemcpyAsync(hd.data(), d_diagv.p, sizeof(double) * hd.size(), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        double mx = 0;
        for (double v : hd) mx = std::max(mx, std::fabs(v));
        lambda = P.tau * mx;
      }
      ni = 2; nBad = 0;
    }
    double rho = 0;
    int qmax = 0;
    do {
      const int nbuf = 1 - cur;
      CK(cudaMemsetAsync(c.d_fail.p, 0, sizeof(int), st));
      k_pg_damp<<<(std::max(nb * 144, np * 12) + 255) / 256, 256, 0, st>>>(nb, d_diag.p, d_hpp.p, lambda, c.d_hs.p, np, d_bp.p, bs);
      CK(cudaGetLastError());
      CK(cudaGraphLaunch(c.chol_graph, st));
      CK(cudaGraphLaunch(c.chol_back_graph, st));
      k_pg_update<<<(n + 63) / 64, 64, 0, st>>>(V, lambda, c.d_x.p, d_bp.p, d_S[cur].p, d_S[nbuf].p, d_pscale.p);
      k_reduce<<<1, 256, 0, st>>>(d_pscale.p, np, nullptr, 0, d_scal.p + 1);
      CK(cudaGetLastError());
      CKR(errors(nbuf, &tempChi));
      const bool ok2 = *(int*)(h_scal + 4) == 0;
      if (!ok2) tempChi = std::numeric_limits<double>::max();
      double scale = h_scal[1];
      rho = currentChi - tempChi;
      scale += 1e-3;
      rho /= scale;
      if (rho > 0 && std::isfinite(tempChi)) {
        double alpha = 1. - std::pow((2 * rho - 1), 3);
        alpha = (std::min)(alpha, P.good_step_upper);
        lambda *= (std::max)(P.good_step_lower, alpha);
        ni = 2;
        currentChi = tempChi;
        cur = nbuf;
      } else {
        lambda *= ni;
        ni *= 2;
      }
      qmax++;
    } while (rho < 0 && qmax < P.max_trials_after_failure);
    if (trace && it < GPBA_MAX_ITERS) {
      trace->levenberg_iterations[it] = qmax; trace->chi2_before[it] = iniChi; trace->chi2_after[it] = currentChi;
      trace->lambda[it] = lambda; trace->total_trials += qmax; trace->last_trial_chi2 = tempChi;
    }
    ++cj;
    if (qmax == P.max_trials_after_failure || rho == 0) { result = GPBA_TERMINATE; ok = false; }
    else {
      if ((iniChi - currentChi) * 1e3 < iniChi) nBad++; else nBad = 0;
      if (nBad >= 3) { result = GPBA_TERMINATE; ok = false; }
    }
  }
  if (trace) { trace->n_iters = cj; trace->result = result; }
  if (sim3_out) CK(cudaMemcpyAsync(sim3_out, d_S[cur].p, sizeof(double) * 8 * (size_t)n, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return GPBA_OK;
}

int gpba_correct_points(int device, int64_t n_pt, const double* xyz, const int32_t* ref_kf, int32_t n_kf, const double* sim3_before,
                        const double* sim3_after, double* xyz_out) {
  if (n_pt < 0 || n_kf < 0 || (n_pt > 0 && (!xyz || !ref_kf || !sim3_before || !sim3_after || !xyz_out))) { g_err = "invalid argument"; return GPBA_ERR_INVALID; }
  for (int64_t i = 0; i < n_pt; ++i) if (ref_kf[i] < 0 || ref_kf[i] >= n_kf) { g_err = "reference keyframe index out of range"; return GPBA_ERR_INVALID; }
  if (n_pt == 0) return GPBA_OK;
  {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { g_err = "no CUDA device (libgpba has no CPU fallback)"; return GPBA_ERR_NO_DEVICE; }
  }
  if (device < 0) CK(cudaGetDevice(&device));
  CK(cudaSetDevice(device));
  cudaStream_t st = small_call_stream(device);
  if (!st) { g_err = "cudaStreamCreate failed"; return GPBA_ERR_CUDA; }
  g_alloc_stream = st;
  DBuf<double> d_x, d_a, d_b, d_o;
  DBuf<int> d_r;
  CKR(d_x.upload(xyz, (size_t)3 * n_pt, st)); CKR(d_r.upload(ref_kf, (size_t)n_pt, st));
  CKR(d_a.upload(sim3_before, (size_t)8 * n_kf, st)); CKR(d_b.upload(sim3_after, (size_t)8 * n_kf, st));
  CKR(d_o.alloc((size_t)3 * n_pt));
  k_correct_points<<<(int)std::min<int64_t>((n_pt + 255) / 256, 148 * 8), 256, 0, st>>>(n_pt, d_x.p, d_r.p, d_a.p, d_b.p, d_o.p);
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(xyz_out, d_o.p, sizeof(double) * 3 * (size_t)n_pt, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return GPBA_OK;
}

// ---- velocity RANSAC (Tracking::MCRansac / Optimizer::OptimizeVel, src/Tracking.cc:1939-2002, src/Optimizer.cc:2364-2447)
int gpba_vel_ransac(const gpba_vel_batch* B, int device, double* vel_out, int32_t* inliers_out, uint8_t* inlier_mask_out,
                    int32_t* best_out, gpba_lm_trace* traces) {
  if (!B || B->n_cam < 1 || B->n_cam > GPBA_VEL_MAX_CAM || !B->cam_intr || !B->cam_Tbc || !B->cam_dt || B->n_match < 0 || B->n_hyp < 0 ||
      B->set_size < 1 || B->set_size > GPBA_VEL_MAX_SET || B->iterations < 0 || B->iterations > GPBA_MAX_ITERS) { g_err = "invalid velocity batch"; return GPBA_ERR_INVALID; }
  if (best_out) *best_out = -1;
  if (B->n_hyp == 0) return GPBA_OK;
  if (!B->samples || (B->n_match > 0 && (!B->obs_u || !B->obs_v || !B->obs_inv_sigma2 || !B->obs_xw || !B->obs_cam))) { g_err = "invalid velocity batch"; return GPBA_ERR_INVALID; }
  for (int i = 0; i < B->n_match; ++i) if (B->obs_cam[i] < 0 || B->obs_cam[i] >= B->n_cam) { g_err = "camera index out of range"; return GPBA_ERR_INVALID; }
  for (int64_t i = 0; i < (int64_t)B->n_hyp * B->set_size; ++i) if (B->samples[i] < 0 || B->samples[i] >= B->n_match) { g_err = "sample index out of range"; return GPBA_ERR_INVALID; }
  {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { g_err = "no CUDA device (libgpba has no CPU fallback)"; return GPBA_ERR_NO_DEVICE; }
  }
  if (device < 0) CK(cudaGetDevice(&device));
  CK(cudaSetDevice(device));
  cudaStream_t st = small_call_stream(device);
  if (!st) { g_err = "cudaStreamCreate failed"; return GPBA_ERR_CUDA; }
  g_alloc_stream = st;
  std::vector<CamConst> cams(B->n_cam);
  for (int c = 0; c < B->n_cam; ++c) {
    CamConst& cc = cams[c];
    cc.fx = B->cam_intr[4 * c]; cc.fy = B->cam_intr[4 * c + 1]; cc.cx = B->cam_intr[4 * c + 2]; cc.cy = B->cam_intr[4 * c + 3];
    SE3 Tbc = load_se3(B->cam_Tbc + 7 * c);
    SE3 Tcb = se3_inv(Tbc);
    M3 Rcb = quat_to_R(Tcb.q), Rbc = quat_to_R(Tbc.q);
    for (int i = 0; i < 9; ++i) { cc.Rcb[i] = Rcb.a[i]; cc.Rbc[i] = Rbc.a[i]; }
    for (int i = 0; i < 3; ++i) { cc.tcb[i] = Tcb.t[i]; cc.tbc[i] = Tbc.t[i]; }
    cc.qbc[0] = Tbc.q.x; cc.qbc[1] = Tbc.q.y; cc.qbc[2] = Tbc.q.z; cc.qbc[3] = Tbc.q.w;
  }
  const size_t nm = (size_t)B->n_match, nh = (size_t)B->n_hyp;
  DBuf<CamConst> d_cam;
  DBuf<double> d_dt, d_u, d_v, d_w, d_xw, d_vel;
  DBuf<int> d_camof, d_samples, d_inl;
  DBuf<uint8_t> d_mask;
  DBuf<gpba_lm_trace> d_tr;
  CKR(d_cam.upload(cams, st)); CKR(d_dt.upload(B->cam_dt, (size_t)B->n_cam, st));
  CKR(d_u.upload(B->obs_u, nm, st)); CKR(d_v.upload(B->obs_v, nm, st)); CKR(d_w.upload(B->obs_inv_sigma2, nm, st));
  CKR(d_xw.upload(B->obs_xw, 3 * nm, st)); CKR(d_camof.upload(B->obs_cam, nm, st));
  CKR(d_samples.upload(B->samples, nh * B->set_size, st));
  CKR(d_vel.alloc(6 * nh)); CKR(d_inl.alloc(nh));
  if (inlier_mask_out) CKR(d_mask.alloc(nh * std::max<size_t>(nm, 1)));
  if (traces) { CKR(d_tr.alloc(nh)); CK(cudaMemsetAsync(d_tr.p, 0, sizeof(gpba_lm_trace) * nh, st)); }
  VelView V;
  V.n_cam = B->n_cam; V.n_match = B->n_match; V.n_hyp = B->n_hyp; V.set_size = B->set_size; V.iterations = B->iterations;
  V.cam = d_cam.p; V.cam_dt = d_dt.p;
  store_se3(se3_inv(load_se3(B->last_pose)), V.Tinv);
  for (int i = 0; i < 6; ++i) V.vel_init[i] = B->vel_init[i];
  V.obs_u = d_u.p; V.obs_v = d_v.p; V.obs_w = d_w.p; V.obs_xw = d_xw.p; V.obs_cam = d_camof.p; V.samples = d_samples.p;
  V.hub_delta = B->huber_delta; V.hub_dsqr = f32sq(B->huber_delta); V.threshold = B->threshold;
  V.vel_out = d_vel.p; V.inliers_out = d_inl.p; V.mask_out = inlier_mask_out ? d_mask.p : nullptr; V.traces = traces ? d_tr.p : nullptr;
  const bool verbose = getenv("GPBA_VERBOSE") != nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (verbose) { CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1)); CK(cudaEventRecord(e0, st)); }
  k_vel_ransac<<<B->n_hyp, GPBA_VEL_THREADS, 0, st>>>(V);
  CK(cudaGetLastError());
  if (verbose) CK(cudaEventRecord(e1, st));
  std::vector<int> inl(nh);
  CK(cudaMemcpyAsync(inl.data(), d_inl.p, sizeof(int) * nh, cudaMemcpyDeviceToHost, st));
  if (vel_out) CK(cudaMemcpyAsync(vel_out, d_vel.p, sizeof(double) * 6 * nh, cudaMemcpyDeviceToHost, st));
  if (inlier_mask_out && nm) CK(cudaMemcpyAsync(inlier_mask_out, d_mask.p, nh * nm, cudaMemcpyDeviceToHost, st));
  if (traces) CK(cudaMemcpyAsync(traces, d_tr.p, sizeof(gpba_lm_trace) * nh, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  if (verbose) {
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    fprintf(stderr, "[gpba] velocity ransac: %d hypotheses, %d matches, kernel %.3f ms\n", B->n_hyp, B->n_match, ms);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
  }
  int best = -1, best_inl = 0;
  for (size_t h = 0; h < nh; ++h) {
    if (inliers_out) inliers_out[h] = inl[h];
    if (inl[h] > best_inl) { best_inl = inl[h]; best = (int)h; }   // `inliers > bestInliers` (Tracking.cc:1973)
  }
  if (best_out) *best_out = best;
  return GPBA_OK;
}

}  // extern "C"

#ifdef GPBA_CHOL_TIMING
extern "C" int gpba_debug_chol_clocks(long long* out) {
  return cudaMemcpyFromSymbol(out, gpba::g_chol_clk, sizeof(long long) * 64) == cudaSuccess ? 0 : -2;
}
#endif
