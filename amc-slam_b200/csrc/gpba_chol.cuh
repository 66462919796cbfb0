// gpba_chol.cuh -- K5b: tile-sparse blocked FP64 Cholesky of the reduced camera system.
//
// Replaces LinearSolverDense::solve (Eigen::LDLT, g2o/solvers/linear_solver_dense.h:65-113) and, for
// global BA, LinearSolverEigen::solve (SimplicialLDLT, linear_solver_eigen.h:94-124).  The matrix is
// stored as NB x NB tiles of the lower triangle; only tiles that are structurally non-zero after a
// tile-level symbolic factorization (host, once per buildStructure -- the analogue of SimplicialLDLT's
// analyzePattern, linear_solver_eigen.h:147-201) are allocated and visited, so a banded covisibility
// pattern costs O(n b^2) while a fully dense system degenerates to the classic right-looking blocked
// algorithm.  Every GEMM-shaped part runs on the FP64 tensor pipe (mma.sync.m8n8k4.f64 -> DMMA).
//
// Schedule (captured once per structure in two CUDA graphs and replayed per LM trial).  Tile columns are grouped in
// LEVELS (a column depends on column k iff tile (j, k) != 0; the nested-dissection order of gpba_order.h makes the levels
// few and wide); the columns of a level are independent and share launches through CTA tables:
//   graph A  k_chol_load                  Hschur blocks -> tiles, bschur -> permuted rhs
//            for every level:
//              k_chol_lupdate(level)  LEFT-LOOKING: one CTA per tile (i, j) of the level's columns:
//                                     A_ij -= sum_k L_ik L_jk^T over the finished columns k that touch both rows, the
//                                     operand tiles streamed by TMA bulk copies (cp.async.bulk + mbarrier, 3 stages)
//                                     while the products accumulate in DMMA fragments; the tile is written ONCE, by one
//                                     CTA, in a fixed order (deterministic, no atomics; a right-looking update re-read
//                                     and re-wrote the target tile for every product: 72 KB of L2 traffic per product
//                                     against 36 KB here).  Diagonal CTAs also do b_j -= sum_k L_jk y_k.
//              k_chol_panel(level)    every CTA: potrf(j,j) in shared memory (redundant: saves a launch on the
//                                     critical path); CTA 0: L_jj^-1 (kept for the backward solve), y_j = L_jj^-1 b_j;
//                                     CTA q>0: L_ij = A_ij L_jj^-T for one tile below the diagonal
//   graph B  for every level, last to first:
//              k_chol_back(level)     x_i = L_ii^-T y_i; one CTA per tile (i,k): y_k -= L_ik^T x_i
//            k_chol_unpermute
// A tile is stored in global memory as its shared-memory image (48 rows padded to 52 doubles: DMMA fragment loads are
// bank-conflict free), so one 1-D bulk copy moves a tile and no thread touches it on the way.
// The code is loop-structured on purpose: these kernels run one short CTA per launch, so straight-line
// unrolled substitutions (~100 KB of SASS) were instruction-fetch bound (ncu: stalled_no_instruction 7/issue).
// A non-positive pivot sets *fail (=> solve() returns false => LM rejects the trial, like
// !_cholesky.isPositive(), linear_solver_dense.h:108-112).
#pragma once
#include "gpba_kernels.cuh"

namespace gpba {

#define GPBA_NB 48
#define GPBA_LD 52   // row stride of a tile (global and shared): DMMA fragment loads are bank-conflict free, rows 16-byte aligned
#define GPBA_TILE (GPBA_NB * GPBA_LD)            // doubles per stored tile
#define GPBA_TILE_BYTES (GPBA_TILE * 8)          // 19 968 B: a multiple of 16, as cp.async.bulk requires
#define GPBA_PANEL_THREADS 192
#define GPBA_LU_STAGES 3

#ifdef GPBA_CHOL_TIMING
__device__ long long g_chol_clk[64];
__device__ __forceinline__ long long* chol_clk_smem() { __shared__ long long b[64]; return b; }
// ticks go to shared memory and are dumped at the end: a global store in front of a barrier stalls ~900 cycles
#define GPBA_TICK(slot) do { if (threadIdx.x == 0 && q_ == 1 && k == 100) chol_clk_smem()[slot] = clock64(); } while (0)
#define GPBA_TICK_DUMP() do { if (threadIdx.x == 0 && q_ == 1 && k == 100) for (int q_ = 0; q_ < 64; ++q_) g_chol_clk[q_] = chol_clk_smem()[q_]; } while (0)
#define GPBA_TICKP(slot) do { GPBA_TICK(16 + pb * 4 + (slot) - 5); \
  if (q_ == 1 && k == 100 && pb == 5) { if (threadIdx.x == 74) chol_clk_smem()[40 + (slot) - 5] = clock64(); if (threadIdx.x == 160) chol_clk_smem()[44 + (slot) - 5] = clock64(); if (threadIdx.x == 64) chol_clk_smem()[48 + (slot) - 5] = clock64(); } } while (0)
#else
#define GPBA_TICK(slot) do { } while (0)
#define GPBA_TICK_DUMP() do { } while (0)
#define GPBA_TICKP(slot) do { } while (0)
#endif

// Programmatic dependent launch (PDL): every kernel of the factorization chain first waits for its predecessor
// (griddepcontrol.wait: predecessor complete, its stores visible), then lets its own successor become resident
// (griddepcontrol.launch_dependents), so the successor's launch latency is hidden behind this kernel's work.
// Both are no-ops when the kernel was launched without the PDL attribute.
GPBA_D void pdl_wait_then_release() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

struct CholView {
  int NT;                       // tiles per side
  int n;                        // true dimension (12 * n_pose)
  const int64_t* tile_off;      // [NT*NT] offset (doubles) of tile (i,j), i>=j, or -1
  const int* col_begin;         // [NT+1] into col_rows
  const int* col_rows;          // rows i>k with L_ik != 0, ascending
  const int* row_begin;         // [NT+1] into row_cols
  const int* row_cols;          // columns k<i with L_ik != 0, ascending
  double* tiles;
  const int* perm;              // [n/12] fill-reducing order of the pose blocks: permuted position of block b
  const unsigned char* pos_used; // [NT*NB/12] the position holds a pose block (parts of the order start on tile boundaries)
  double* dinv;                 // [NT][NB*NB] inverse of the diagonal factor tiles (lower, row-major)
  double* work;                 // [NT*NB] rhs -> y in permuted order
  double* xsol;                 // [NT*NB] solution in permuted order
};

// Pad the unused positions, scatter the upper Hschur blocks (row-major 12x12, block (bi,bj), bi<=bj) into the lower
// tiles and load the permuted right-hand side.  The tiles were zeroed by a memset node before.
__global__ void k_chol_load(CholView C, int n_hs, const int* __restrict__ hs_row, const int* __restrict__ hs_col,
                            const double* __restrict__ hs, const double* __restrict__ rhs) {
  const int64_t n = (int64_t)n_hs * 144;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x, t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (int64_t j = t0; j < n; j += stride) {
    const int blk = (int)(j / 144), e = (int)(j % 144);
    const int r = e / 12, c = e % 12;
    const int bi = hs_row[blk], bj = hs_col[blk];
    const int pi = C.perm[bi], pj = C.perm[bj];
    // symmetric matrix: element (bi*12+r, bj*12+c) == element (bj*12+c, bi*12+r); store whichever lands in the
    // lower triangle of the permuted matrix
    int R, Cc;
    if (pi == pj) { if (c < r) continue; R = pi * 12 + c; Cc = pi * 12 + r; }
    else if (pi > pj) { R = pi * 12 + r; Cc = pj * 12 + c; }
    else { R = pj * 12 + c; Cc = pi * 12 + r; }
    const int ti = R / GPBA_NB, tj = Cc / GPBA_NB;
    C.tiles[C.tile_off[(size_t)ti * C.NT + tj] + (R % GPBA_NB) * GPBA_LD + (Cc % GPBA_NB)] = hs[j];
  }
  const int NTNB = C.NT * GPBA_NB;
  for (int64_t j = t0; j < NTNB; j += stride) {
    if (C.pos_used[j / 12]) continue;
    const int t = (int)j / GPBA_NB, o = (int)j % GPBA_NB;
    C.tiles[C.tile_off[(size_t)t * C.NT + t] + o * GPBA_LD + o] = 1.0;  // identity on the padding positions
    C.work[j] = 0.0;
  }
  for (int64_t j = t0; j < C.n; j += stride) C.work[C.perm[j / 12] * 12 + j % 12] = rhs[j];
}

// reciprocal for the pivot chain: MUFU seed + two Newton steps (55 cycles dependent vs 85 for the IEEE division,
// tools/microbench.cu); relative error ~1 ulp, far below what the factorization needs
GPBA_D double fast_rcp(double d) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
  double e = fma(-d, y, 1.0);
  y = fma(y, e, y);
  e = fma(-d, y, 1.0);
  return fma(y, e, y);
}

// 1/sqrt(d) for the column scaling: MUFU seed + two Newton steps instead of the library routine (~40 instructions, issued
// 48 times per thread and tile by the two warps that also carry the pivot chain)
GPBA_D double fast_rsqrt(double d) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
  const double h = 0.5 * d;
  double e = fma(-h * y, y, 0.5);   // 0.5 - 0.5 d y^2
  y = fma(y, e, y);
  e = fma(-h * y, y, 0.5);
  return fma(y, e, y);
}

// inverse of the b-th 8x8 diagonal block of L (read from S), column `lane` per thread, lanes 0..7 of one warp
GPBA_D void diag_block_inverse(const double (*S)[GPBA_LD], int b, int lane, double (*D8)[8][8]) {
  if (lane >= 8) return;
  const int c0 = 8 * b;
  double x[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) x[q] = (q == lane) ? 1.0 : 0.0;
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    x[q] *= fast_rcp(S[c0 + q][c0 + q]);
#pragma unroll
    for (int q2 = q + 1; q2 < 8; ++q2) x[q2] = fma(-S[c0 + q2][c0 + q], x[q], x[q2]);
  }
#pragma unroll
  for (int q = 0; q < 8; ++q) D8[b][q][lane] = x[q];
}

// Cholesky of the 48x48 tile in S (lower part; the upper part is never read), GPBA_PANEL_THREADS threads.
// Six panels of 8 columns, left-looking: before a panel is factorized its (48 - 8 pb) x 8 column strip receives the
// updates of all previous panels in one go -- one 8x8 tile per warp, two DMMA accumulator chains of depth pb each
// (a right-looking trailing update touched m(m+1)/2 tiles per panel in up to three rounds per warp and cost 470-940
// cycles per panel against ~150-400 here; same flops, fewer and longer chains, one shared-memory round trip per tile).
// Then warps 0-1, which own the 48 rows, factorize: every one of their threads factorizes the 8x8 diagonal block of the
// panel redundantly in registers (LDL^T form: the pivot chain is one reciprocal + one FMA per pivot and needs no
// communication -- a shared-memory + barrier round trip per pivot cost ~380 cycles) and solves its own row of the panel
// with it.  Meanwhile warp 3 inverts the previous panel's diagonal block for the blocked triangular solves; the other
// warps stay off the FP64 pipe.  On exit S = L, D8[b] = (b-th 8x8 diagonal block of L)^-1.
// CTA-wide barrier of the 192 factorizing threads: the whole CTA in k_chol_panel, the six consumer warps (named barrier 2)
// in the persistent kernel, whose producer warp must not take part
template <bool NAMED> GPBA_D void tile_sync() {
  if (NAMED) asm volatile("barrier.sync 2, 192;" ::: "memory");
  else __syncthreads();
}

template <bool NAMED = false>
GPBA_D void potrf48(double (*S)[GPBA_LD], double (*D8)[8][8], int* fail, int k = -1, int q_ = -1) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, gid = lane >> 2, tig = lane & 3;
  const int r = tid;
#pragma unroll 1
  for (int pb = 0; pb < GPBA_NB / 8; ++pb) {
    const int c0 = 8 * pb;
    if (pb > 0) {
      // S[8 mt + gid][c0 + 2 tig ..] -= sum over previous columns kc of L[8 mt + gid][kc] L[c0 + n][kc], mt = pb + warp
      const int mt = pb + warp;
      if (mt < GPBA_NB / 8) {
        double* cp = &S[8 * mt + gid][c0 + 2 * tig];
        double2 c = *reinterpret_cast<double2*>(cp);
        double e0 = 0.0, e1 = 0.0;   // second accumulator pair: two independent DMMA chains
#pragma unroll 1
        for (int kp = 0; kp < pb; ++kp) {
          dmma884(c.x, c.y, -S[8 * mt + gid][8 * kp + tig], S[c0 + gid][8 * kp + tig]);
          dmma884(e0, e1, -S[8 * mt + gid][8 * kp + 4 + tig], S[c0 + gid][8 * kp + 4 + tig]);
        }
        c.x += e0; c.y += e1;
        *reinterpret_cast<double2*>(cp) = c;   // columns >= c0: nobody reads them in this phase (operands are columns < c0)
      }
      tile_sync<NAMED>();
    }
    GPBA_TICKP(7);
    if (warp < 2) {
      double u[8][8], p[8], is[8];
      const bool mine = r >= c0 && r < GPBA_NB;
#pragma unroll
      for (int q = 0; q < 8; ++q)
#pragma unroll
        for (int i = 0; i <= q; ++i) u[q][i] = S[c0 + q][c0 + i];
#pragma unroll
      for (int i = 0; i < 8; ++i) p[i] = mine ? S[r][c0 + i] : 0.0;  // rows of the block itself: entries i > r - c0 are unused
      asm volatile("barrier.sync 1, 64;" ::: "memory");  // the block is in registers: its rows may be overwritten
      GPBA_TICKP(5);
      bool bad = false;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const double d = u[i][i];
        bad = bad || !(d > 0.0);
        const double rc = fast_rcp(d);
        is[i] = fast_rsqrt(d);  // off the pivot chain: L[q][i] = u[q][i] / sqrt(d_i)
#pragma unroll
        for (int q = i + 1; q < 8; ++q) {
          const double t = u[q][i] * rc;
#pragma unroll
          for (int i2 = i + 1; i2 <= q; ++i2) u[q][i2] = fma(-t, u[i2][i], u[q][i2]);
        }
        const double t = p[i] * rc;
#pragma unroll
        for (int i2 = i + 1; i2 < 8; ++i2) p[i2] = fma(-t, u[i2][i], p[i2]);
      }
      GPBA_TICKP(6);
      if (bad && tid == 0) atomicExch(fail, 1);
      if (mine) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
          if (r >= c0 + i) S[r][c0 + i] = p[i] * is[i];
      }
    } else if (warp == 3 && pb > 0) {
      diag_block_inverse(S, pb - 1, lane, D8);
    }
    tile_sync<NAMED>();
    GPBA_TICKP(8);
  }
  if (warp == 3) diag_block_inverse(S, GPBA_NB / 8 - 1, lane, D8);
  tile_sync<NAMED>();
}

// X = T L^-T in place (T: 48 x 48 in shared memory), blocked by 8: warp w owns the 8 rows of row-tile w, so the
// six chains are independent.  X_j = (T_j - sum_{p<j} X_p L_jp^T) L_jj^-T, both products on DMMA.
GPBA_D void trsm48(double (*T)[GPBA_LD], const double (*S)[GPBA_LD], const double (*D8)[8][8]) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, gid = lane >> 2, tig = lane & 3;
  if (warp >= GPBA_NB / 8) return;
  const int row = 8 * warp + gid;
#pragma unroll 1
  for (int j = 0; j < GPBA_NB / 8; ++j) {
    double* cp = &T[row][8 * j + 2 * tig];
    double2 c = *reinterpret_cast<double2*>(cp);
    double e0 = 0.0, e1 = 0.0;  // second accumulator pair: two independent DMMA chains
#pragma unroll 1
    for (int p = 0; p < j; ++p) {
      dmma884(c.x, c.y, -T[row][8 * p + tig], S[8 * j + gid][8 * p + tig]);
      dmma884(e0, e1, -T[row][8 * p + 4 + tig], S[8 * j + gid][8 * p + 4 + tig]);
    }
    c.x += e0; c.y += e1;
    *reinterpret_cast<double2*>(cp) = c;
    __syncwarp();
    double2 x = make_double2(0.0, 0.0);
#pragma unroll
    for (int kk = 0; kk < 2; ++kk) dmma884(x.x, x.y, T[row][8 * j + 4 * kk + tig], D8[j][gid][4 * kk + tig]);
    __syncwarp();
    *reinterpret_cast<double2*>(cp) = x;
    __syncwarp();
  }
}

// Panel step of the tile columns of one level.  One CTA per table entry (k, q): q = 0 is the diagonal CTA of column k,
// q > 0 its q-th non-zero tile below the diagonal.
__global__ void __launch_bounds__(GPBA_PANEL_THREADS) k_chol_panel(CholView C, const int2* __restrict__ tab, int* __restrict__ fail,
                                                                   int* __restrict__ level_done) {
  const int k = tab[blockIdx.x].x, q_ = tab[blockIdx.x].y;
  __shared__ __align__(16) double S[GPBA_NB][GPBA_LD];
  __shared__ __align__(16) double T[GPBA_NB][GPBA_LD];
  __shared__ double D8[GPBA_NB / 8][8][8];
  __shared__ double Sy[GPBA_NB];
  const int tid = threadIdx.x;
  __shared__ __align__(8) unsigned long long bar;
  GPBA_TICK(0);
  const double* Tkk = C.tiles + C.tile_off[(size_t)k * C.NT + k];   // static index data: may be read before the wait
  double* A = q_ == 0 ? nullptr : C.tiles + C.tile_off[(size_t)C.col_rows[C.col_begin[k] + q_ - 1] * C.NT + k];
  if (tid == 0) { mbar_init(&bar, 1); mbar_init_fence(); }
  pdl_wait_then_release();
  if (tid == 0) {   // the tiles are stored as their shared-memory image: one bulk copy each, no thread touches the data
    mbar_expect_tx(&bar, (q_ == 0 ? 1u : 2u) * GPBA_TILE_BYTES);
    bulk_g2s(&S[0][0], Tkk, GPBA_TILE_BYTES, &bar);
    if (q_ != 0) bulk_g2s(&T[0][0], A, GPBA_TILE_BYTES, &bar);
  }
  if (q_ == 0) {
    for (int j = tid; j < GPBA_NB * GPBA_NB; j += blockDim.x) T[j / GPBA_NB][j % GPBA_NB] = (j / GPBA_NB == j % GPBA_NB) ? 1.0 : 0.0;
    if (tid < GPBA_NB) Sy[tid] = C.work[k * GPBA_NB + tid];
  }
  __syncthreads();   // the barrier's initialisation is visible to every waiter
  mbar_wait(&bar, 0);
  GPBA_TICK(1);
  potrf48(S, D8, fail, k, q_);
  GPBA_TICK(2);
  trsm48(T, S, D8);
  __syncthreads();
  GPBA_TICK(3);
  if (q_ == 0) {
    // T = L_kk^-T.  dinv[k] = L_kk^-1 = T^T (row-major), y_k = L_kk^-1 b_k
    double* D = C.dinv + (size_t)k * GPBA_NB * GPBA_NB;
    for (int j = tid; j < GPBA_NB * GPBA_NB; j += blockDim.x) D[j] = T[j % GPBA_NB][j / GPBA_NB];
    if (tid < GPBA_NB) {
      double y0 = 0.0, y1 = 0.0;
#pragma unroll 4
      for (int c = 0; c < GPBA_NB; c += 2) { y0 = fma(T[c][tid], Sy[c], y0); y1 = fma(T[c + 1][tid], Sy[c + 1], y1); }
      C.work[k * GPBA_NB + tid] = y0 + y1;
    }
  } else {
    for (int j = tid; j < GPBA_NB * GPBA_NB / 2; j += blockDim.x) {
      const int r = j / (GPBA_NB / 2), c = 2 * (j % (GPBA_NB / 2));
      *reinterpret_cast<double2*>(A + r * GPBA_LD + c) = *reinterpret_cast<double2*>(&T[r][c]);
    }
  }
  GPBA_TICK(4);
  GPBA_TICK_DUMP();
  // The next level's left-looking update already runs (it was launched when this grid became resident) and consumes the
  // older columns; its producers wait for this counter before they touch a column of THIS level.
  __threadfence();
  __syncthreads();
  if (tid == 0) atomicAdd(level_done, 1);
}

// Left-looking update of the tile columns of one level: tile (i, j) of a column of the level receives
//     A_ij -= sum_k L_ik L_jk^T                 diagonal tiles also:  b_j -= sum_k L_jk y_k
// over the finished columns k < j with L_ik != 0 and L_jk != 0 (klist: the symbolic phase intersected the two row
// patterns).  The host cuts the lists into CHUNKS of <= GPBA_LU_CHUNK products, table entry {j, q, kb, ke} (q = 0: the
// diagonal tile, q > 0: the q-th non-zero row below it).  One SM needs 0.88 us of DMMA issue per 48^3 product at the
// measured 37 TFLOP/s, so what matters is that all SMs stay busy for the whole launch: PERSISTENT CTAs (one per SM) draw
// chunks from an atomic counter, and the operand stream never drains: a PRODUCER WARP (one thread) walks the chunk
// stream and keeps GPBA_LU_STAGES operand pairs in flight with cp.async.bulk (one bulk copy per tile, completing on the
// stage's `full` mbarrier) ACROSS chunk boundaries -- its index loads and the atomic never sit on the consumers' path -- and
// leaves a small descriptor per stage telling the consumers where a chunk begins and ends and where it goes.  Four
// CONSUMER WARPS take a stage when its barrier flips -- each owns a 24 x 24 corner = 3 x 3 DMMA tiles (9 independent
// accumulator chains) -- and hand it back through the stage's `empty` mbarrier (no CTA-wide barrier in the loop).  A chunk's
// partial sum leaves with red.global.add.f64 (fire and forget: no read of the target tile on the way; several chunks,
// possibly on different SMs, may feed one tile).
#define GPBA_LU_CHUNK 2
#define GPBA_LU_THREADS 160   // warps 0-3 consume (DMMA), warp 4 produces (TMA)
struct LuDesc { double* target; int j; int flags; int task; };   // flags: 1 first product of its chunk, 2 last product, 4 stop, 8 diagonal tile
#define GPBA_LU_LATE 0x80000000u   // klist entry: the source column belongs to the level whose panel step is still in flight
__global__ void __launch_bounds__(GPBA_LU_THREADS) k_chol_lupdate(CholView C, const int4* __restrict__ tab, int n_chunks,
                                                                  const unsigned* __restrict__ klist, int* __restrict__ counter,
                                                                  const int* __restrict__ prev_done, int prev_count, int* __restrict__ fail) {
  extern __shared__ __align__(128) unsigned char lu_smem[];
  __shared__ __align__(8) unsigned long long full[GPBA_LU_STAGES], empty[GPBA_LU_STAGES];
  __shared__ LuDesc desc[GPBA_LU_STAGES];
  typedef double (*TileP)[GPBA_LD];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // stage s: [La | Lb | y_k]
  constexpr int STAGE_BYTES = 2 * GPBA_TILE_BYTES + GPBA_NB * 8;
  auto stage_a = [&](int s) { return reinterpret_cast<TileP>(lu_smem + (size_t)s * STAGE_BYTES); };
  auto stage_b = [&](int s) { return reinterpret_cast<TileP>(lu_smem + (size_t)s * STAGE_BYTES + GPBA_TILE_BYTES); };
  auto stage_y = [&](int s) { return reinterpret_cast<double*>(lu_smem + (size_t)s * STAGE_BYTES + 2 * GPBA_TILE_BYTES); };
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < GPBA_LU_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 4); }
    mbar_init_fence();
  }
  // LOOK-AHEAD.  This grid is the programmatic dependent of the previous level's panel grid and deliberately does NOT wait
  // for it (no griddepcontrol.wait): that grid executed its own wait -- everything older than it is complete and visible --
  // before it let this one start, so only the previous level's columns are still being written.  Products whose source is
  // such a column (GPBA_LU_LATE, sorted to the end of every list and of the chunk table) wait for the panel grid's completion
  // counter; all others (the bulk: contributions of columns finished long ago) overlap with the panel step.  The panel grid
  // is resident in full before this grid exists (that is when a programmatic dependent may start), so the wait cannot
  // deadlock.  Without programmatic launches the grids run in stream order and the counter is already there.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  __syncthreads();           // barrier initialisation visible to both roles
  if (warp == 4) {
    // ---------------------------------------------------------------- producer (one thread): the operand stream
    if (lane != 0) return;
    int pos = 0, begin = 0, end = 0, pi = 0, pj = 0;
    double* target = nullptr;
    bool diag = false, prev_ready = prev_count <= 0;
    for (int n = 0;; ++n) {
      const int s = n % GPBA_LU_STAGES;
      if (n >= GPBA_LU_STAGES) mbar_wait(&empty[s], (unsigned)(n / GPBA_LU_STAGES - 1) & 1u);   // the four warps released the stage
      if (pos == end) {
        const int c = atomicAdd(counter, 1);
        if (c >= n_chunks) {
          desc[s].target = nullptr; desc[s].j = -1; desc[s].flags = 4;
          mbar_arrive(&full[s]);
          return;
        }
        const int4 e = tab[c];
        pj = e.x; diag = e.y == 0;
        pi = diag ? e.x : C.col_rows[C.col_begin[e.x] + e.y - 1];
        target = C.tiles + C.tile_off[(size_t)pi * C.NT + pj];
        begin = pos = e.z; end = e.w;
      }
      const unsigned ke = klist[pos];
      const int k = (int)(ke & ~GPBA_LU_LATE);
      if ((ke & GPBA_LU_LATE) && !prev_ready) {
        // the source column is being written by the panel grid in flight: wait for its last CTA (bounded: a grid that
        // never completes is reported as a failed factorization instead of a hang)
        unsigned spins = 0;
        int seen;
        do {
          asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(seen) : "l"(prev_done) : "memory");
        } while (seen < prev_count && ++spins < (1u << 26));
        if (seen < prev_count) atomicExch(fail, 1);
        asm volatile("fence.proxy.async;" ::: "memory");   // the bulk copies below (async proxy) read what that grid wrote
        prev_ready = true;
      }
      desc[s].target = target; desc[s].j = pj;
      desc[s].flags = (pos == begin ? 1 : 0) | (pos + 1 == end ? 2 : 0) | (diag ? 8 : 0);
      mbar_expect_tx(&full[s], diag ? GPBA_TILE_BYTES + GPBA_NB * 8u : 2u * GPBA_TILE_BYTES);
      bulk_g2s(stage_a(s), C.tiles + C.tile_off[(size_t)pi * C.NT + k], GPBA_TILE_BYTES, &full[s]);
      if (!diag) bulk_g2s(stage_b(s), C.tiles + C.tile_off[(size_t)pj * C.NT + k], GPBA_TILE_BYTES, &full[s]);
      else bulk_g2s(stage_y(s), C.work + (size_t)k * GPBA_NB, GPBA_NB * 8, &full[s]);
      ++pos;
    }
  }
  // ------------------------------------------------------------------ consumers (warps 0-3)
  const int gid = lane >> 2, tig = lane & 3;
  const int m0 = 3 * (warp >> 1), n0 = 3 * (warp & 1);
  double2 acc[3][3];
  double s0 = 0.0, s1 = 0.0;   // rhs rows (diagonal tiles): warp w owns rows 12 w .. 12 w + 11, lanes 0..11
  for (int n = 0;; ++n) {
    const int s = n % GPBA_LU_STAGES;
    mbar_wait(&full[s], (unsigned)(n / GPBA_LU_STAGES) & 1u);
    const LuDesc d = desc[s];
    if (d.flags & 4) break;    // the stream is exhausted
    const bool diag = (d.flags & 8) != 0;
    if (d.flags & 1) {
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b) acc[a][b] = make_double2(0.0, 0.0);
      s0 = s1 = 0.0;
    }
    const TileP La = stage_a(s);
    const TileP Lb = diag ? La : stage_b(s);
#pragma unroll 2
    for (int k0 = 0; k0 < GPBA_NB; k0 += 4) {
      double af[3], bf[3];
#pragma unroll
      for (int a = 0; a < 3; ++a) { af[a] = La[8 * (m0 + a) + gid][k0 + tig]; bf[a] = Lb[8 * (n0 + a) + gid][k0 + tig]; }
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b) dmma884(acc[a][b].x, acc[a][b].y, af[a], bf[b]);
    }
    const int rrow = 12 * warp + lane;   // forward substitution with the tile already staged
    if (diag && lane < 12) {
      const double* yk = stage_y(s);
#pragma unroll 4
      for (int c = 0; c < GPBA_NB; c += 2) { s0 = fma(La[rrow][c], yk[c], s0); s1 = fma(La[rrow][c + 1], yk[c + 1], s1); }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[s]);   // this warp is done with stage s (and its descriptor, held in registers)
    if (d.flags & 2) {   // the chunk is complete: its partial sum joins the target tile
      double* Tc = d.target;
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b) {
          double* out = Tc + (8 * (m0 + a) + gid) * GPBA_LD + 8 * (n0 + b) + 2 * tig;
          atomicAdd(out, -acc[a][b].x); atomicAdd(out + 1, -acc[a][b].y);
        }
      if (diag && lane < 12) atomicAdd(&C.work[(size_t)d.j * GPBA_NB + rrow], -(s0 + s1));
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// k_chol_factor: the WHOLE factorization (and forward substitution) as one persistent dataflow kernel.
//
// The level-by-level schedule above pays two kernel boundaries per level on the critical path (~25 us per level at C4
// against ~8 us of dependent work).  Here the CTAs stay resident (as many as fit: the host sizes the grid with the
// occupancy calculator, so every CTA is resident from the start) and draw TASKS from one list in level order through an
// atomic counter:
//   update chunk {j, q, kb, ke}   A_ij -= sum_k L_ik L_jk^T over klist[kb..ke) (and b_j -= L_jk y_k on diagonal tiles)
//   panel task   {j, q, -1, 0}    potrf(A_jj); q = 0: L_jj^-1 and y_j; q > 0: L_ij = A_ij L_jj^-T
// Dependencies are per tile COLUMN and travel through two completion counters in global memory:
//   upd_cnt[j]  += 1 per consumer warp and chunk once its partial sum has been added to column j   (panel tasks of j wait)
//   pan_cnt[k]  += 1 per finished panel task of column k                                            (products reading k wait)
// A task only depends on tasks in front of it in the list; tasks are claimed in list order and every claimed task is held by
// a resident CTA that executes its tasks in order, so the waits cannot deadlock (induction over the task index).  The
// waits are bounded anyway: a counter that never arrives is reported as a failed factorization instead of a hang.
// Inside a CTA the producer warp (one thread) claims tasks, performs the waits (ld.acquire.gpu, then fence.proxy.async
// because the operands are fetched by the async proxy) and streams the operands with cp.async.bulk into a ring of stages;
// six consumer warps execute the items in order: an update product on 24 x 24 corners (four of the warps), or a
// panel item with the 192-thread potrf48 / trsm48 above, IN the stage buffers.
#define GPBA_CF_THREADS 256   // warps 0-5 consume, warp 6 produces, warp 7 publishes finished chunks
#define GPBA_CF_SIG 4          // ring of finished-chunk messages between the accumulating warps and the publishing thread
#define GPBA_CF_D8_BYTES (GPBA_NB / 8 * 64 * 8)   // the six inverted 8 x 8 diagonal blocks of a factorized diagonal tile
#define GPBA_CF_STAGE_BYTES (2 * GPBA_TILE_BYTES + GPBA_NB * 8 + GPBA_CF_D8_BYTES)
struct ChFactorArgs {
  const int4* tab; int n_tasks;   // chunk {j | DIAG, target tile, pb, pe}; panel task {j | DIAG, target tile, -1 / -2 / -3, diagonal tile of j}
                                  //   -1: potrf + solve (every task of the column factorizes the diagonal tile itself: shortest chain)
                                  //   -2: diagonal task that also PUBLISHES L_jj and its inverted diagonal blocks; -3: solve only, with the
                                  //       published factor (levels with more panel tasks than CTAs: 5.3 us of potrf per task saved)
  const int4* prod;               // per product {tile of L_ik, tile of L_jk, k, -}: everything the producer needs in one load
  const int* upd_need;   // [NT] chunks that target column j
  const int* pan_need;   // [NT] panel tasks of column k
  int* task_counter;     // zeroed by a memset node of the graph, like the two arrays below
  int* upd_cnt;          // [NT]
  int* pan_cnt;          // [NT]
  int* diag_cnt;         // [NT] 1 once the published factor of the diagonal tile is in place
  double* d8;            // [NT][6][8][8] published inverted diagonal blocks
  int* fail;
  long long* trace;      // optional [4 n_tasks] ns: claimed, dependencies met (last wait), first item consumed, done (tools only)
};
#define GPBA_CF_DIAG 0x40000000
GPBA_D long long cf_now() { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
GPBA_D void cf_wait_counter(const int* cnt, int need, int* fail) {
  unsigned spins = 0;
  int seen;
  do {
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(seen) : "l"(cnt) : "memory");
  } while (seen < need && ++spins < (1u << 26));
  if (seen < need) atomicExch(fail, 1);
  asm volatile("fence.proxy.async;" ::: "memory");   // the bulk copies (async proxy) read what the counted tasks wrote
}
#define GPBA_CF_STAGES 3
__global__ void __launch_bounds__(GPBA_CF_THREADS, 1) k_chol_factor(CholView C, ChFactorArgs F) {
  extern __shared__ __align__(128) unsigned char cf_smem[];
  __shared__ __align__(8) unsigned long long full[GPBA_CF_STAGES], empty[GPBA_CF_STAGES];
  __shared__ LuDesc desc[GPBA_CF_STAGES];
  __shared__ double D8[GPBA_NB / 8][8][8];
  __shared__ __align__(8) unsigned long long sig_full[GPBA_CF_SIG], sig_empty[GPBA_CF_SIG];
  __shared__ int sig_j[GPBA_CF_SIG], sig_task[GPBA_CF_SIG];
  typedef double (*TileP)[GPBA_LD];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  auto stage_a = [&](int s) { return reinterpret_cast<TileP>(cf_smem + (size_t)s * GPBA_CF_STAGE_BYTES); };
  auto stage_b = [&](int s) { return reinterpret_cast<TileP>(cf_smem + (size_t)s * GPBA_CF_STAGE_BYTES + GPBA_TILE_BYTES); };
  auto stage_y = [&](int s) { return reinterpret_cast<double*>(cf_smem + (size_t)s * GPBA_CF_STAGE_BYTES + 2 * GPBA_TILE_BYTES); };
  auto stage_d8 = [&](int s) { return reinterpret_cast<double(*)[8][8]>(cf_smem + (size_t)s * GPBA_CF_STAGE_BYTES + 2 * GPBA_TILE_BYTES + GPBA_NB * 8); };
  // columns this CTA has already seen complete (completion is final, so a positive answer can be cached): one bit per column
  unsigned* ready_bits = reinterpret_cast<unsigned*>(cf_smem + (size_t)GPBA_CF_STAGES * GPBA_CF_STAGE_BYTES);
  for (int j = tid; j < (C.NT + 31) / 32; j += blockDim.x) ready_bits[j] = 0u;
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < GPBA_CF_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 6); }
#pragma unroll
    for (int s = 0; s < GPBA_CF_SIG; ++s) { mbar_init(&sig_full[s], 4); mbar_init(&sig_empty[s], 1); }
    mbar_init_fence();
  }
  __syncthreads();
  if (warp == 7) {
    // ---------------------------------------------------------------- publisher (one thread)
    // A finished chunk ends with 36 reductions per thread into the target tile.  Making them visible (a gpu-scope fence that
    // waits for all of them) and counting the chunk took 2 us per chunk on the accumulating warps; they now only announce
    // the chunk on a shared-memory barrier (release at CTA scope) and go on, and this thread observes the barrier (acquire),
    // fences at gpu scope -- cumulative: it covers the reductions it has observed through the barrier, the pattern of a
    // grid-wide barrier (__syncthreads, then ONE thread fences and signals) -- and counts.
    if (lane != 0) return;
    for (int m = 0;; ++m) {
      const int q = m % GPBA_CF_SIG;
      mbar_wait(&sig_full[q], (unsigned)(m / GPBA_CF_SIG) & 1u);
      const int j = sig_j[q], task = sig_task[q];
      mbar_arrive(&sig_empty[q]);
      if (j < 0) return;
      __threadfence();
      atomicAdd(F.upd_cnt + j, 1);
      if (F.trace) F.trace[8 * (size_t)task + 3] = cf_now();
    }
  }
  if (warp == 6) {
    // ---------------------------------------------------------------- producer (one thread)
    // Everything it needs per item comes from one record (no dependent index chains), and the NEXT task is claimed while
    // the current one is being streamed: the atomic's round trip and the record loads are off the consumers' path.
    if (lane != 0) return;
    int pos = 0, begin = 0, end = 0, pj = 0, task = 0;
    double* target = nullptr;
    bool diag = false;
    int next_c = atomicAdd(F.task_counter, 1);
    int4 next_e = make_int4(0, 0, 0, 0);
    bool next_loaded = false;
    int4 rec = make_int4(0, 0, 0, 0);   // record of product `pos` (prefetched one item ahead inside a chunk)
    for (int n = 0;; ++n) {
      const int s = n % GPBA_CF_STAGES;
      {
        const long long tw = F.trace ? cf_now() : 0;
        if (n >= GPBA_CF_STAGES) mbar_wait(&empty[s], (unsigned)(n / GPBA_CF_STAGES - 1) & 1u);
        if (F.trace && pos != end) F.trace[8 * (size_t)task + 6] += cf_now() - tw;
      }
      if (pos == end) {
        if (next_c >= F.n_tasks) {
          desc[s].target = nullptr; desc[s].j = -1; desc[s].flags = 4;
          mbar_arrive(&full[s]);
          return;
        }
        const int4 e = next_loaded ? next_e : F.tab[next_c];
        task = next_c;
        next_c = atomicAdd(F.task_counter, 1);   // claim ahead: the result is not needed before the next item
        next_loaded = false;
        if (F.trace) F.trace[8 * (size_t)task] = cf_now();
        pj = e.x & ~GPBA_CF_DIAG; diag = (e.x & GPBA_CF_DIAG) != 0;
        target = C.tiles + (size_t)e.y * GPBA_TILE;
        if (e.z < 0) {
          // panel item: the column has received all its updates (solve-only tasks: and its diagonal tile is factorized)
          const bool solve_only = e.z == -3;
          if (solve_only) cf_wait_counter(F.diag_cnt + pj, 1, F.fail);
          else cf_wait_counter(F.upd_cnt + pj, F.upd_need[pj], F.fail);
          if (F.trace) F.trace[8 * (size_t)task + 1] = cf_now();
          desc[s].target = target; desc[s].j = pj; desc[s].task = task;
          desc[s].flags = 16 | (diag ? 8 : 0) | (e.z == -2 ? 32 : 0) | (solve_only ? 64 : 0);
          mbar_expect_tx(&full[s], diag ? GPBA_TILE_BYTES + GPBA_NB * 8u : 2u * GPBA_TILE_BYTES + (solve_only ? (unsigned)GPBA_CF_D8_BYTES : 0u));
          bulk_g2s(stage_a(s), C.tiles + (size_t)e.w * GPBA_TILE, GPBA_TILE_BYTES, &full[s]);
          if (!diag) bulk_g2s(stage_b(s), target, GPBA_TILE_BYTES, &full[s]);
          else bulk_g2s(stage_y(s), C.work + (size_t)pj * GPBA_NB, GPBA_NB * 8, &full[s]);
          if (solve_only) bulk_g2s(stage_d8(s), F.d8 + (size_t)pj * (GPBA_CF_D8_BYTES / 8), GPBA_CF_D8_BYTES, &full[s]);
          continue;
        }
        begin = pos = e.z; end = e.w;
        rec = F.prod[pos];
      }
      const int k = rec.z;
      if (!(ready_bits[k >> 5] >> (k & 31) & 1u)) {
        cf_wait_counter(F.pan_cnt + k, F.pan_need[k], F.fail);   // the source column is final (its y_k too)
        ready_bits[k >> 5] |= 1u << (k & 31);
      }
      if (F.trace) F.trace[8 * (size_t)task + 1] = cf_now();
      desc[s].target = target; desc[s].j = pj; desc[s].task = task;
      desc[s].flags = (pos == begin ? 1 : 0) | (pos + 1 == end ? 2 : 0) | (diag ? 8 : 0);
      mbar_expect_tx(&full[s], diag ? GPBA_TILE_BYTES + GPBA_NB * 8u : 2u * GPBA_TILE_BYTES);
      bulk_g2s(stage_a(s), C.tiles + (size_t)rec.x * GPBA_TILE, GPBA_TILE_BYTES, &full[s]);
      if (!diag) bulk_g2s(stage_b(s), C.tiles + (size_t)rec.y * GPBA_TILE, GPBA_TILE_BYTES, &full[s]);
      else bulk_g2s(stage_y(s), C.work + (size_t)k * GPBA_NB, GPBA_NB * 8, &full[s]);
      ++pos;
      if (pos < end) rec = F.prod[pos];   // in flight while the consumers work on this stage
      else if (next_c < F.n_tasks) { next_e = F.tab[next_c]; next_loaded = true; }   // the claim issued at the start of this chunk has returned
    }
  }
  // ------------------------------------------------------------------ consumers (warps 0-5)
  const int gid = lane >> 2, tig = lane & 3;
  const int m0 = 3 * ((warp & 3) >> 1), n0 = 3 * (warp & 1);
  int cn = 0;   // finished chunks of this CTA (message ring position)
  auto announce = [&](int j, int task) {   // warps 0-3, after their reductions have been issued
    const int q = cn % GPBA_CF_SIG;
    if (cn >= GPBA_CF_SIG) mbar_wait(&sig_empty[q], (unsigned)(cn / GPBA_CF_SIG - 1) & 1u);
    if (tid == 0) { sig_j[q] = j; sig_task[q] = task; }
    __syncwarp();
    if (lane == 0) mbar_arrive(&sig_full[q]);
    ++cn;
  };
  for (int n = 0;;) {
    int s = n % GPBA_CF_STAGES;
    mbar_wait(&full[s], (unsigned)(n / GPBA_CF_STAGES) & 1u);
    LuDesc d = desc[s];
    if (d.flags & 4) { if (warp < 4) announce(-1, 0); break; }
    const bool diag = (d.flags & 8) != 0;
    if (F.trace && tid == 0) F.trace[8 * (size_t)d.task + 2] = cf_now();
    if (d.flags & 16) {
      // ---- panel item: S = A_jj, T = A_ij (or the identity for the diagonal task), both in the stage
      const TileP S = stage_a(s), T = stage_b(s);
      double* Sy = stage_y(s);
      if (d.flags & 64) {
        trsm48(T, S, stage_d8(s));   // the column's diagonal task has published L_jj and its inverted diagonal blocks
      } else {
        if (diag)
          for (int j = tid; j < GPBA_NB * GPBA_NB; j += 192) T[j / GPBA_NB][j % GPBA_NB] = (j / GPBA_NB == j % GPBA_NB) ? 1.0 : 0.0;
        tile_sync<true>();
        potrf48<true>(S, D8, F.fail);
        if (F.trace && tid == 0) F.trace[8 * (size_t)d.task + 4] = cf_now();
        if (d.flags & 32) {
          // publish the factor first: the solve-only tasks of the column are waiting for it
          double* A = d.target;   // the diagonal tile itself
          for (int j = tid; j < GPBA_TILE / 2; j += 192) reinterpret_cast<double2*>(A)[j] = reinterpret_cast<const double2*>(&S[0][0])[j];
          double* G = F.d8 + (size_t)d.j * (GPBA_CF_D8_BYTES / 8);
          for (int j = tid; j < GPBA_CF_D8_BYTES / 8; j += 192) G[j] = (&D8[0][0][0])[j];
          tile_sync<true>();   // then ONE thread fences (cumulative over what the barrier made it observe) and signals: the grid-barrier pattern
          if (tid == 0) { __threadfence(); atomicAdd(F.diag_cnt + d.j, 1); }
        }
        trsm48(T, S, D8);
      }
      tile_sync<true>();
      if (F.trace && tid == 0) F.trace[8 * (size_t)d.task + 5] = cf_now();
      if (diag) {
        double* D = C.dinv + (size_t)d.j * GPBA_NB * GPBA_NB;
        for (int j = tid; j < GPBA_NB * GPBA_NB; j += 192) D[j] = T[j % GPBA_NB][j / GPBA_NB];
        if (tid < GPBA_NB) {
          double y0 = 0.0, y1 = 0.0;
#pragma unroll 4
          for (int c = 0; c < GPBA_NB; c += 2) { y0 = fma(T[c][tid], Sy[c], y0); y1 = fma(T[c + 1][tid], Sy[c + 1], y1); }
          C.work[d.j * GPBA_NB + tid] = y0 + y1;
        }
      } else {
        double* A = d.target;
        for (int j = tid; j < GPBA_NB * GPBA_NB / 2; j += 192) {
          const int r = j / (GPBA_NB / 2), c = 2 * (j % (GPBA_NB / 2));
          *reinterpret_cast<double2*>(A + r * GPBA_LD + c) = *reinterpret_cast<double2*>(&T[r][c]);
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic writes to the stage before the next bulk copy into it
      tile_sync<true>();
      if (tid == 0) { __threadfence(); atomicAdd(F.pan_cnt + d.j, 1); if (F.trace) F.trace[8 * (size_t)d.task + 3] = cf_now(); }
      if (lane == 0) mbar_arrive(&empty[s]);
      ++n;
      continue;
    }
    // ---- update chunk: its products arrive back to back (the accumulators live only here, not across panel items).
    // Warps 0-3 -- one per SM sub-partition, so the four FP64 tensor pipes carry equal shares -- own a 24 x 24 corner each
    // (3 x 3 DMMA tiles, 9 independent accumulator chains); warps 4-5 only take part in panel items and hand the stage back.
    if (warp >= 4) {
      for (;;) {
        if (lane == 0) mbar_arrive(&empty[s]);
        ++n;
        if (d.flags & 2) break;
        s = n % GPBA_CF_STAGES;
        mbar_wait(&full[s], (unsigned)(n / GPBA_CF_STAGES) & 1u);
        d = desc[s];
      }
      continue;
    }
    double2 acc[3][3];
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int b = 0; b < 3; ++b) acc[a][b] = make_double2(0.0, 0.0);
    double s0 = 0.0, s1 = 0.0;   // rhs rows (diagonal tiles): warp w owns rows 12 w .. 12 w + 11, lanes 0..11
    const int rrow = 12 * warp + lane;
    long long twait = 0;
    for (;;) {
      const TileP La = stage_a(s);
      const TileP Lb = diag ? La : stage_b(s);
#pragma unroll 2
      for (int k0 = 0; k0 < GPBA_NB; k0 += 4) {
        double af[3], bf[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) { af[a] = La[8 * (m0 + a) + gid][k0 + tig]; bf[a] = Lb[8 * (n0 + a) + gid][k0 + tig]; }
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
          for (int b = 0; b < 3; ++b) dmma884(acc[a][b].x, acc[a][b].y, af[a], bf[b]);
      }
      if (diag && lane < 12) {
        const double* yk = stage_y(s);
#pragma unroll 4
        for (int c = 0; c < GPBA_NB; c += 2) { s0 = fma(La[rrow][c], yk[c], s0); s1 = fma(La[rrow][c + 1], yk[c + 1], s1); }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[s]);
      ++n;
      if (d.flags & 2) break;
      s = n % GPBA_CF_STAGES;
      const long long tw = (F.trace && tid == 0) ? cf_now() : 0;
      mbar_wait(&full[s], (unsigned)(n / GPBA_CF_STAGES) & 1u);
      if (F.trace && tid == 0) twait += cf_now() - tw;
      d = desc[s];
    }
    if (F.trace && tid == 0) { F.trace[8 * (size_t)d.task + 4] = twait; F.trace[8 * (size_t)d.task + 5] = cf_now(); }
    double* Tc = d.target;
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int b = 0; b < 3; ++b) {
        double* out = Tc + (8 * (m0 + a) + gid) * GPBA_LD + 8 * (n0 + b) + 2 * tig;
        atomicAdd(out, -acc[a][b].x); atomicAdd(out + 1, -acc[a][b].y);
      }
    if (diag && lane < 12) atomicAdd(&C.work[(size_t)d.j * GPBA_NB + rrow], -(s0 + s1));
    announce(d.j, d.task);   // the publisher makes the partial sum visible and counts the chunk
  }
}

// Backward substitution, tile row i (launched for i = NT-1 .. 0): every CTA computes x_i = L_ii^-T y_i
// (y_i has received the contributions of all later rows); CTA 0 stores it, CTA q > 0 applies one tile of
// row i to an earlier segment: y_k -= L_ik^T x_i.  Distinct k per CTA and rows are stream ordered, so the
// result is deterministic without atomics.
template <bool ATOMIC>
__global__ void __launch_bounds__(192) k_chol_back(CholView C, const int2* __restrict__ tab) {
  __shared__ double yi[GPBA_NB], xi[GPBA_NB], part[4][GPBA_NB];
  const int i = tab[blockIdx.x].x, q_ = tab[blockIdx.x].y;   // tile row i; q_ = 0: store x_i, q_ > 0: its q_-th tile
  const int tid = threadIdx.x, c = tid % GPBA_NB, h = tid / GPBA_NB;  // 4 row-quarters x 48 columns
  // all global operands are requested up front (they do not depend on x_i): one memory round trip per launch
  const double* D = C.dinv + (size_t)i * GPBA_NB * GPBA_NB;  // (L^-T)[c][r] = Linv[r][c], zero for r < c
  const int k = q_ > 0 ? C.row_cols[C.row_begin[i] + q_ - 1] : 0;
  const double* L = q_ > 0 ? C.tiles + C.tile_off[(size_t)i * C.NT + k] : D;
  const int ldl = q_ > 0 ? GPBA_LD : GPBA_NB;   // factor tiles are stored padded, the inverses of the diagonal tiles dense
  pdl_wait_then_release();
  double dv[12], lv[12];
#pragma unroll
  for (int r = 0; r < 12; ++r) { dv[r] = D[(12 * h + r) * GPBA_NB + c]; lv[r] = L[(12 * h + r) * ldl + c]; }
  const double yk_old = (!ATOMIC && q_ > 0 && tid < GPBA_NB) ? C.work[k * GPBA_NB + tid] : 0.0;
  if (tid < GPBA_NB) yi[tid] = C.work[i * GPBA_NB + tid];
  __syncthreads();
  double s = 0.0;
#pragma unroll
  for (int r = 0; r < 12; ++r) s = fma(dv[r], yi[12 * h + r], s);
  part[h][c] = s;
  __syncthreads();
  if (tid < GPBA_NB) {
    const double x = (part[0][tid] + part[1][tid]) + (part[2][tid] + part[3][tid]);
    xi[tid] = x;
    if (q_ == 0) C.xsol[i * GPBA_NB + tid] = x;  // not in place: the other CTAs of this launch still read y_i
  }
  if (q_ == 0) return;
  __syncthreads();
  s = 0.0;
#pragma unroll
  for (int r = 0; r < 12; ++r) s = fma(lv[r], xi[12 * h + r], s);
  part[h][c] = s;
  __syncthreads();
  if (tid < GPBA_NB) {
    const double sum = (part[0][tid] + part[1][tid]) + (part[2][tid] + part[3][tid]);
    if (ATOMIC) atomicAdd(&C.work[k * GPBA_NB + tid], -sum);   // rows of one level may feed the same earlier segment
    else C.work[k * GPBA_NB + tid] = yk_old - sum;
  }
}

__global__ void k_chol_unpermute(CholView C, double* __restrict__ xout) {
  pdl_wait_then_release();
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < C.n) xout[j] = C.xsol[C.perm[j / 12] * 12 + j % 12];
}

}  // namespace gpba
