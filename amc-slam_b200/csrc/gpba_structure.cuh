// gpba_structure.cuh -- P0 on the device: the index structures of one optimize() call.
//
// BlockSolver::buildStructure (Thirdparty/g2o/g2o/core/block_solver.hpp:142-295) walks every landmark's edge list
// through std::map / unordered_map on one core; here the O(sum d^2) part -- which pairs of observations of a
// landmark meet in which block -- is one radix sort.  Everything is integer work and deterministic (cub's radix sort
// is stable), so the resulting pattern is bit-exact and the summation order of every block is fixed.
//
//   observations sorted by landmark  ->  k_emit_pairs      one (record pair key, observation pair) per unordered
//                                                         pair of observations of a landmark (self pairs included)
//                                    ->  cub radix sort    by record pair key
//                                    ->  cub run-length    unique record pairs + their list offsets
// The host then turns the (few) unique record pairs into the Hschur block pattern (pose pairs) and the per-block
// contribution lists; see Solver::build_structure.
#pragma once
#include <cub/cub.cuh>
#include "gpba_kernels.cuh"

namespace gpba {

// ---- active set and orderings (SparseOptimizer::initializeOptimization, sparse_optimizer.cpp:199-267)
// One pass over the observations: which records / points carry an active (level 0) edge, first keyframe of every
// point, whether any level-1 edge exists.  Plain stores of the same value and atomicMin: the result is order independent.
__global__ void k_scan_obs(int64_t n_obs, const uint8_t* __restrict__ flags, const int* __restrict__ rec, const int* __restrict__ pt,
                           const int* __restrict__ rec_kf2, unsigned char* __restrict__ rec_used, int* __restrict__ pt_act,
                           int* __restrict__ first_kf, int* __restrict__ any_level1) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_obs; i += (int64_t)gridDim.x * blockDim.x) {
    if (flags[i] & 0x2u) { *any_level1 = 1; continue; }
    const int r = rec[i], p = pt[i];
    rec_used[r] = 1;
    pt_act[p] = 1;
    atomicMin(first_kf + p, rec_kf2[r]);
  }
}
// landmark order key of a point: first keyframe (locality of record / pose accesses); inactive points sort last
__global__ void k_point_keys(int n_pt, const int* __restrict__ pt_act, const int* __restrict__ first_kf, int n_kf,
                             int* __restrict__ key, int* __restrict__ val) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p < n_pt) { key[p] = pt_act[p] ? first_kf[p] : n_kf + 1; val[p] = p; }
}
__global__ void k_point_index(int n_pt, int n_lm_all, const int* __restrict__ sorted_pt, int* __restrict__ pt_lm_all) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l < n_pt) pt_lm_all[sorted_pt[l]] = l < n_lm_all ? l : -1;
}
// observations per landmark (all ranks' landmarks, level 0 only / any level)
__global__ void k_count_lm_obs(int64_t n_obs, const uint8_t* __restrict__ flags, const int* __restrict__ pt,
                               const int* __restrict__ pt_lm_all, int any_level, int* __restrict__ cnt) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_obs; i += (int64_t)gridDim.x * blockDim.x) {
    if (!any_level && (flags[i] & 0x2u)) continue;
    const int l = pt_lm_all[pt[i]];
    if (l >= 0) atomicAdd(cnt + l, 1);
  }
}
// sort key of an observation: its landmark relative to the owned range, or the sentinel `n_own` (dropped)
__global__ void k_obs_keys(int64_t n_obs, const uint8_t* __restrict__ flags, const int* __restrict__ pt, const int* __restrict__ pt_lm_all,
                           int own_lo, int n_own, int any_level, int* __restrict__ key, int64_t* __restrict__ val) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_obs; i += (int64_t)gridDim.x * blockDim.x) {
    int l = -1;
    if (any_level || !(flags[i] & 0x2u)) l = pt_lm_all[pt[i]];
    l = l >= 0 ? l - own_lo : -1;
    key[i] = (l >= 0 && l < n_own) ? l : n_own;
    val[i] = i;
  }
}
__global__ void k_gather_int(int64_t n, const int64_t* __restrict__ idx, const int* __restrict__ src, int* __restrict__ dst) {
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x) dst[j] = src[idx[j]];
}
__global__ void k_hist_int(int64_t n, const int* __restrict__ key, int* __restrict__ cnt) {
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x) atomicAdd(cnt + key[j], 1);
}
// n(n+1)/2 observation pairs per landmark
__global__ void k_pair_counts(int n_lm, const int* __restrict__ cnt, int64_t* __restrict__ np) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l < n_lm) { const int64_t n = cnt[l]; np[l] = n * (n + 1) / 2; }
  else if (l == n_lm) np[l] = 0;
}
__global__ void k_widen(int n, const int* __restrict__ in, int64_t* __restrict__ out) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l < n) out[l] = in[l];
}
__global__ void k_lm_rank(int n_lm, const int* __restrict__ lm_pt, const int* __restrict__ rank_of_pt, int* __restrict__ lm_rank) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l < n_lm) lm_rank[l] = rank_of_pt[lm_pt[l]];
}
__global__ void k_check_indices(int64_t n_obs, const int* __restrict__ rec, const int* __restrict__ pt, int n_rec, int n_pt, int* __restrict__ bad) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_obs; i += (int64_t)gridDim.x * blockDim.x)
    if (rec[i] < 0 || rec[i] >= n_rec || pt[i] < 0 || pt[i] >= n_pt) *bad = 1;
}

// gather of the landmark-sorted observation arrays, split in two for the asynchronous upload: indices first, measurements
// when their copy has landed
__global__ void k_gather_obs_idx(int64_t n, const int64_t* __restrict__ o_orig, const int* __restrict__ rec,
                                 const uint8_t* __restrict__ flags, int* __restrict__ srec, uint8_t* __restrict__ sfl) {
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = o_orig[j];
    srec[j] = rec[i]; sfl[j] = flags[i];
  }
}
__global__ void k_gather_obs_meas(int64_t n, const int64_t* __restrict__ o_orig, const double* __restrict__ u,
                                  const double* __restrict__ v, const double* __restrict__ ur, const double* __restrict__ w,
                                  double* __restrict__ su, double* __restrict__ sv, double* __restrict__ sur, double* __restrict__ sw) {
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = o_orig[j];
    su[j] = u[i]; sv[j] = v[i]; sw[j] = w[i];
    if (ur) sur[j] = ur[i];
  }
}

// record-major copies of the landmark-sorted observation inputs (K2b streams them)
__global__ void k_gather_recmajor(int64_t n, const int64_t* __restrict__ rperm, const double* __restrict__ u, const double* __restrict__ v,
                                  const double* __restrict__ ur, const double* __restrict__ w, const int* __restrict__ lm,
                                  const uint8_t* __restrict__ flags, double* __restrict__ ru, double* __restrict__ rv,
                                  double* __restrict__ rur, double* __restrict__ rw, int* __restrict__ rlm, uint8_t* __restrict__ rfl) {
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = rperm[j];
    ru[j] = u[i]; rv[j] = v[i]; rw[j] = w[i]; rlm[j] = lm[i]; rfl[j] = flags[i];
    if (ur) rur[j] = ur[i];
  }
}

__global__ void k_iota(int64_t n, int64_t* __restrict__ out) {
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x) out[j] = j;
}

__global__ void k_fill_lm(int n_lm, const int64_t* __restrict__ lm_obs_begin, int* __restrict__ o_lm) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31, nw = (gridDim.x * blockDim.x) >> 5;
  for (int l = warp; l < n_lm; l += nw)
    for (int64_t j = lm_obs_begin[l] + lane; j < lm_obs_begin[l + 1]; j += 32) o_lm[j] = l;
}

// One warp per landmark.  Observation pairs (a <= b, positions inside the landmark) are enumerated row by row;
// the key orders the two records, the value keeps the observation of the smaller record first.
// dup[0] is raised when two different observations of one landmark share a record (never produced by the reference:
// a MapPoint holds one observation per keyframe and camera).
// The key is (landmark chunk, record pair): work items of K4b are then (chunk, record pair) runs, so that the U rows
// of one chunk of landmarks (a few MB) serve all their ~11 uses from L2 before the next chunk is touched.
__global__ void k_emit_pairs(int n_lm, const int64_t* __restrict__ lm_obs_begin, const int64_t* __restrict__ lm_pair_begin,
                             const int* __restrict__ o_rec, unsigned long long n_rec, int lm_chunk, unsigned long long* __restrict__ keys,
                             unsigned long long* __restrict__ vals, int* __restrict__ dup) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31, nw = (gridDim.x * blockDim.x) >> 5;
  for (int l = warp; l < n_lm; l += nw) {
    const int64_t ob = lm_obs_begin[l];
    const int n = (int)(lm_obs_begin[l + 1] - ob);
    const int64_t pb = lm_pair_begin[l];
    const int np = n * (n + 1) / 2;
    for (int idx = lane; idx < np; idx += 32) {
      // idx = b*(b+1)/2 + a with a <= b
      int b = (int)((sqrtf(8.0f * (float)idx + 1.0f) - 1.0f) * 0.5f);
      while ((b + 1) * (b + 2) / 2 <= idx) ++b;
      while (b * (b + 1) / 2 > idx) --b;
      const int a = idx - b * (b + 1) / 2;
      const unsigned long long oa = (unsigned long long)(ob + a), obb = (unsigned long long)(ob + b);
      const unsigned long long ra = (unsigned long long)o_rec[ob + a], rb = (unsigned long long)o_rec[ob + b];
      if (a != b && ra == rb) atomicExch(dup, 1);
      const bool swap = ra > rb;
      const unsigned long long chunk = lm_chunk > 0 ? (unsigned long long)(l / lm_chunk) : 0ull;
      keys[pb + idx] = chunk * n_rec * n_rec + (swap ? rb * n_rec + ra : ra * n_rec + rb);
      vals[pb + idx] = swap ? (obb << 32) | oa : (oa << 32) | obb;
    }
  }
}

// #(free pose, landmark) blocks of Hpl = _Hpl->nonZeroBlocks() (block_solver.hpp:206-254), only reported through
// gpba_structure_info: per landmark the number of distinct free poses among its active observations.
__global__ void k_count_hpl(DevView V, const int* __restrict__ o_rec, unsigned long long* __restrict__ total) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31, nw = (gridDim.x * blockDim.x) >> 5;
  unsigned long long cnt = 0;
  auto pose_of = [&](int r, int which) -> int {   // 0 previous keyframe, 1 current keyframe, 2 extrinsic
    if (which == 2) return V.rec_kf1[r] >= 0 ? V.ext_h[V.rec_cam[r]] : -1;
    const int k = which ? V.rec_kf2[r] : V.rec_kf1[r];
    return k >= 0 ? V.kf_h[k] : -1;
  };
  for (int l = warp; l < V.n_lm; l += nw) {
    const int64_t ob = V.lm_obs_begin[l];
    const int m = 3 * (int)(V.lm_obs_begin[l + 1] - ob);
    for (int j = lane; j < m; j += 32) {
      const int h = pose_of(o_rec[ob + j / 3], j % 3);
      if (h < 0) continue;
      bool first = true;
      for (int q = 0; q < j && first; ++q)
        if (pose_of(o_rec[ob + q / 3], q % 3) == h) first = false;
      if (first) ++cnt;
    }
  }
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if (lane == 0 && cnt) atomicAdd(total, cnt);
}

// Record window of every observation tile: [min record, min record + GPBA_WIN_ROWS) clipped to what the tile touches.
// One warp per tile.
__global__ void k_tile_windows(int n_tiles, const int* __restrict__ tile_lm, const int64_t* __restrict__ lm_obs_begin,
                               const int* __restrict__ o_rec, int n_rec, int* __restrict__ rlo, int* __restrict__ rcnt) {
  const int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (t >= n_tiles) return;
  int mn = 0x7fffffff, mx = -1;
  for (int64_t i = lm_obs_begin[tile_lm[t]] + lane; i < lm_obs_begin[tile_lm[t + 1]]; i += 32) { const int r = o_rec[i]; mn = min(mn, r); mx = max(mx, r); }
  for (int o = 16; o > 0; o >>= 1) { mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o)); mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o)); }
  if (lane == 0) {
    if (mx < 0) { rlo[t] = 0; rcnt[t] = 0; }
    else { rlo[t] = mn; rcnt[t] = min(min(mx - mn + 1, GPBA_WIN_ROWS), n_rec - mn); }
  }
}

// run keys (chunk, r1, r2) -> record pair part; idx = iota
__global__ void k_item_low(int n, const unsigned long long* __restrict__ run_key, unsigned long long nrec2,
                           unsigned long long* __restrict__ low, int* __restrict__ idx) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { low[i] = run_key[i] % nrec2; idx[i] = i; }
}
__global__ void k_slot_heads(int n, const unsigned long long* __restrict__ low_sorted, int* __restrict__ head) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p < n) head[p] = (p == 0 || low_sorted[p] != low_sorted[p - 1]) ? 1 : 0;
}
// slot of every item (= rank of its record pair among the unique record pairs), the unique record pairs themselves
__global__ void k_slot_assign(int n, const unsigned long long* __restrict__ low_sorted, const int* __restrict__ idx_sorted,
                              const int* __restrict__ head, const int* __restrict__ slot_p1, unsigned long long n_rec,
                              int* __restrict__ item_rp, unsigned char* __restrict__ item_flags, unsigned long long* __restrict__ rp_key) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const int slot = slot_p1[p] - 1;
  const unsigned long long k = low_sorted[p];
  item_rp[idx_sorted[p]] = slot;
  item_flags[idx_sorted[p]] = (unsigned char)(((k / n_rec == k % n_rec) ? 1 : 0) | 2);  // several chunks feed one slot: accumulate atomically
  if (head[p]) rp_key[slot] = k;
}
// a run of the sorted pair list longer than GPBA_ITEM_PAIRS is cut into several work items (the diagonal runs hold one
// self pair per observation of a record inside the chunk -- hundreds -- and would otherwise be the kernel's critical path)
#define GPBA_ITEM_PAIRS 64
__global__ void k_run_items(int n, const int* __restrict__ count, int* __restrict__ n_sub) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) n_sub[i] = (count[i] + GPBA_ITEM_PAIRS - 1) / GPBA_ITEM_PAIRS;
}
__global__ void k_item_ranges(int n, const int64_t* __restrict__ begin_excl, const int* __restrict__ count, const int* __restrict__ sub_begin,
                              const int* __restrict__ run_rp, const unsigned char* __restrict__ run_flags, int64_t* __restrict__ item_begin,
                              int64_t* __restrict__ item_end, int* __restrict__ item_rp, unsigned char* __restrict__ item_flags) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t b = begin_excl[i], e = b + count[i];
  int it = sub_begin[i];
  for (int64_t q = b; q < e; q += GPBA_ITEM_PAIRS, ++it) {
    item_begin[it] = q;
    item_end[it] = q + GPBA_ITEM_PAIRS < e ? q + GPBA_ITEM_PAIRS : e;
    item_rp[it] = run_rp[i];
    item_flags[it] = run_flags[i];
  }
}

// ---- Hschur block pattern and K4c contribution lists from the unique record pairs (all integer, all on the device)
// pose-like vertex `which` of record r: 0 previous keyframe, 1 current keyframe, 2 the camera's extrinsic (GP records only)
GPBA_D int rec_pose(const DevView& V, int r, int which) {
  if (which == 2) return V.rec_kf1[r] >= 0 ? V.ext_h[V.rec_cam[r]] : -1;
  const int k = which ? V.rec_kf2[r] : V.rec_kf1[r];
  return k >= 0 ? V.kf_h[k] : -1;
}
GPBA_D unsigned long long block_key(int pa, int pb) {  // (col, row) with row <= col: the order of SparseBlockMatrix columns
  const unsigned lo = (unsigned)(pa < pb ? pa : pb), hi = (unsigned)(pa < pb ? pb : pa);
  return ((unsigned long long)hi << 32) | lo;
}
// nine candidate pose-pair blocks per record pair (~0 = none)
#define GPBA_KEYS_PER_RP 9
__global__ void k_emit_block_keys(DevView V, int n, const unsigned long long* __restrict__ rp_key, unsigned long long* __restrict__ out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const int r1 = (int)(rp_key[t] / (unsigned long long)V.n_rec), r2 = (int)(rp_key[t] % (unsigned long long)V.n_rec);
#pragma unroll
  for (int q = 0; q < GPBA_KEYS_PER_RP; ++q) {
    const int pa = rec_pose(V, r1, q / 3), pb = rec_pose(V, r2, q % 3);
    out[GPBA_KEYS_PER_RP * (size_t)t + q] = (pa >= 0 && pb >= 0) ? block_key(pa, pb) : ~0ull;
  }
}
__global__ void k_split_block_keys(int n, const unsigned long long* __restrict__ key, int* __restrict__ row, int* __restrict__ col) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) { row[t] = (int)(unsigned)key[t]; col[t] = (int)(key[t] >> 32); }
}
GPBA_D int find_block(const unsigned long long* __restrict__ hs_key, int n_hs, unsigned long long key) {
  int lo = 0, hi = n_hs - 1;
  while (lo < hi) { const int mid = (lo + hi) >> 1; if (hs_key[mid] < key) lo = mid + 1; else hi = mid; }
  return lo;
}
// up to twelve (block, left record slice, right record slice, transpose) contributions per record pair, see k_schur_expand
// (nine slice pairs; a pair whose two poses coincide across two different records lands twice in the diagonal block);
// sort key = (block, left record, left slice)
#define GPBA_MAX_CON_PER_RP 12
__global__ void k_emit_contribs(DevView V, int n_rp, const unsigned long long* __restrict__ rp_key, const unsigned long long* __restrict__ hs_key,
                                int n_hs, unsigned long long* __restrict__ keys, HsContrib* __restrict__ vals, int* __restrict__ n_valid) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_rp) return;
  const int r1 = (int)(rp_key[t] / (unsigned long long)V.n_rec), r2 = (int)(rp_key[t] % (unsigned long long)V.n_rec);
  int w = 0;
  auto emit = [&](int blk, int rL, int aL, int rR, int aR, int tr, int g) {
    if (w >= GPBA_MAX_CON_PER_RP) return;   // cannot happen: at most 9 slice pairs + 3 coinciding poses
    keys[(size_t)t * GPBA_MAX_CON_PER_RP + w] = ((unsigned long long)blk * (unsigned long long)V.n_rec + (unsigned long long)rL) * 4ull + (unsigned long long)aL;
    vals[(size_t)t * GPBA_MAX_CON_PER_RP + w] = HsContrib{t, rL, rR, aL | (aR << 2) | (tr << 4) | (g << 5)};
    ++w;
  };
  for (int a = 0; a < 3; ++a)
    for (int b = 0; b < 3; ++b) {
      if (r1 == r2 && a > b) continue;  // the mirror image of (b, a)
      const int pa = rec_pose(V, r1, a), pb = rec_pose(V, r2, b);
      if (pa < 0 || pb < 0) continue;
      const int blk = find_block(hs_key, n_hs, block_key(pa, pb));
      if (pa < pb) emit(blk, r1, a, r2, b, 0, 0);
      else if (pa > pb) emit(blk, r2, b, r1, a, 1, 0);       // upper storage holds the transposed product
      else if (r1 == r2) emit(blk, r1, a, r2, b, 0, 1);      // a == b: symmetric, carries g'_r for bschur
      else { emit(blk, r1, a, r2, b, 0, 0); emit(blk, r2, b, r1, a, 1, 0); }  // both ordered pairs land in (pa, pa)
    }
  if (w) atomicAdd(n_valid, w);
  for (; w < GPBA_MAX_CON_PER_RP; ++w) keys[(size_t)t * GPBA_MAX_CON_PER_RP + w] = ~0ull;
}
// group size into the first entry of every (block, left slice) run
__global__ void k_mark_groups(int n_groups, const int* __restrict__ start, const int* __restrict__ count, HsContrib* __restrict__ con) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n_groups) con[start[t]].code |= count[t] << 8;
}
// first contribution of every block (sorted keys)
__global__ void k_con_begin(int n_hs, int n_con, unsigned long long keys_per_block, const unsigned long long* __restrict__ keys, int* __restrict__ con_begin) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b > n_hs) return;
  const unsigned long long want = (unsigned long long)b * keys_per_block;
  int lo = 0, hi = n_con;
  while (lo < hi) { const int mid = (lo + hi) >> 1; if (keys[mid] < want) lo = mid + 1; else hi = mid; }
  con_begin[b] = lo;
}

// cub temp storage that grows on demand
struct CubTemp {
  void* p = nullptr;
  size_t bytes = 0;
  cudaStream_t owner = nullptr;
  ~CubTemp() { if (p) cudaFreeAsync(p, owner); }
  cudaError_t reserve(size_t need, cudaStream_t s) {
    if (need <= bytes) return cudaSuccess;
    if (p) cudaFreeAsync(p, owner);
    p = nullptr; bytes = 0; owner = s;
    cudaError_t e = cudaMallocAsync(&p, need, s);
    if (e == cudaSuccess) bytes = need;
    return e;
  }
};

static inline int bits_for(unsigned long long max_value) {
  int b = 1;
  while (b < 64 && (max_value >> b)) ++b;
  return b;
}

// keys/vals (n entries) -> sorted in place (through the alternate buffers), unique keys + run lengths.
// Returns the number of unique keys in *h_runs (synchronises the stream).
static inline cudaError_t sort_and_encode(CubTemp& tmp, unsigned long long*& keys, unsigned long long*& vals,
                                          unsigned long long*& keys_alt, unsigned long long*& vals_alt, int64_t n, int key_bits,
                                          unsigned long long* uniq, int* counts, int* d_runs, int* h_runs, cudaStream_t s) {
  cub::DoubleBuffer<unsigned long long> dk(keys, keys_alt), dv(vals, vals_alt);
  size_t need = 0;
  cudaError_t e = cub::DeviceRadixSort::SortPairs(nullptr, need, dk, dv, n, 0, key_bits, s);
  if (e != cudaSuccess) return e;
  if ((e = tmp.reserve(need, s)) != cudaSuccess) return e;
  if ((e = cub::DeviceRadixSort::SortPairs(tmp.p, need, dk, dv, n, 0, key_bits, s)) != cudaSuccess) return e;
  if (dk.Current() != keys) { std::swap(keys, keys_alt); }
  if (dv.Current() != vals) { std::swap(vals, vals_alt); }
  need = 0;
  if ((e = cub::DeviceRunLengthEncode::Encode(nullptr, need, keys, uniq, counts, d_runs, n, s)) != cudaSuccess) return e;
  if ((e = tmp.reserve(need, s)) != cudaSuccess) return e;
  if ((e = cub::DeviceRunLengthEncode::Encode(tmp.p, need, keys, uniq, counts, d_runs, n, s)) != cudaSuccess) return e;
  if ((e = cudaMemcpyAsync(h_runs, d_runs, sizeof(int), cudaMemcpyDeviceToHost, s)) != cudaSuccess) return e;
  return cudaStreamSynchronize(s);
}

}  // namespace gpba
