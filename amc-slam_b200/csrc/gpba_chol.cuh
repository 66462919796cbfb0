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
// Schedule (captured once per structure in two CUDA graphs and replayed per LM trial):
//   graph A  k_chol_load                Hschur blocks -> tiles, bschur -> permuted rhs
//            for every tile column k:
//              k_chol_panel(k)  every CTA: potrf(k,k) in shared memory (redundant: saves a launch on the
//                               critical path); CTA 0: L_kk^-1 (kept for the backward solve), y_k = L_kk^-1 b_k;
//                               CTA q>0: L_ik = A_ik L_kk^-T for one tile below the diagonal
//              k_chol_update(k) one CTA per tile pair (a >= b) of column k: A_ab -= L_ak L_bk^T;
//                               last CTA: b_i -= L_ik y_k  (the forward substitution rides along)
//   graph B  for every tile row i, last to first:
//              k_chol_back(i)   x_i = L_ii^-T y_i; one CTA per tile (i,k): y_k -= L_ik^T x_i
//            k_chol_unpermute
// The code is loop-structured on purpose: these kernels run one short CTA per launch, so straight-line
// unrolled substitutions (~100 KB of SASS) were instruction-fetch bound (ncu: stalled_no_instruction 7/issue).
// A non-positive pivot sets *fail (=> solve() returns false => LM rejects the trial, like
// !_cholesky.isPositive(), linear_solver_dense.h:108-112).
#pragma once
#include "gpba_kernels.cuh"

namespace gpba {

#define GPBA_NB 48
#define GPBA_LD 52   // shared-memory row stride: DMMA fragment loads are bank-conflict free, rows 16-byte aligned
#define GPBA_PANEL_THREADS 192

struct CholView {
  int NT;                       // tiles per side
  int n;                        // true dimension (12 * n_pose)
  const int64_t* tile_off;      // [NT*NT] offset (doubles) of tile (i,j), i>=j, or -1
  const int* col_begin;         // [NT+1] into col_rows
  const int* col_rows;          // rows i>k with L_ik != 0, ascending
  const int* row_begin;         // [NT+1] into row_cols
  const int* row_cols;          // columns k<i with L_ik != 0, ascending
  double* tiles;
  const int* perm;              // [n/12] fill-reducing order of the pose blocks: permuted index of block b
  double* dinv;                 // [NT][NB*NB] inverse of the diagonal factor tiles (lower, row-major)
  double* work;                 // [NT*NB] rhs -> y in permuted order
  double* xsol;                 // [NT*NB] solution in permuted order
};

// Pad the diagonal, scatter the upper Hschur blocks (row-major 12x12, block (bi,bj), bi<=bj) into the lower
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
    C.tiles[C.tile_off[(size_t)ti * C.NT + tj] + (R % GPBA_NB) * GPBA_NB + (Cc % GPBA_NB)] = hs[j];
  }
  const int NTNB = C.NT * GPBA_NB;
  for (int64_t j = t0; j < NTNB; j += stride) {
    if (j < C.n) continue;
    const int t = (int)j / GPBA_NB, o = (int)j % GPBA_NB;
    C.tiles[C.tile_off[(size_t)t * C.NT + t] + o * GPBA_NB + o] = 1.0;  // identity on the padding
    C.work[j] = 0.0;
  }
  for (int64_t j = t0; j < C.n; j += stride) C.work[C.perm[j / 12] * 12 + j % 12] = rhs[j];
}

// the two warps that own the 48 rows during a panel step
GPBA_D void bar64() { asm volatile("barrier.sync 1, 64;" ::: "memory"); }

// global 48x48 row-major tile -> shared [48][GPBA_LD] (16-byte loads), any CTA size
GPBA_D void tile_to_smem(const double* __restrict__ T, double (*S)[GPBA_LD]) {
  const double2* T2 = reinterpret_cast<const double2*>(T);
  for (int j = threadIdx.x; j < GPBA_NB * GPBA_NB / 2; j += blockDim.x) {
    const double2 v = T2[j];
    const int r = j / (GPBA_NB / 2), c = 2 * (j % (GPBA_NB / 2));
    *reinterpret_cast<double2*>(&S[r][c]) = v;
  }
}

// Cholesky of the 48x48 tile in S (lower part; the upper part is never read), GPBA_PANEL_THREADS threads.
// Six panels of 8 columns: the panel is factorized one row per thread (registers, one 64-thread barrier per
// pivot), the trailing tile is updated with DMMA by all warps.  On exit S = L, D8[b] = (b-th 8x8 diagonal block
// of L)^-1, Sinv[j] = 1 / L[j][j].
GPBA_D void potrf48(double (*S)[GPBA_LD], double (*D8)[8][8], double* Sinv, int* fail) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, gid = lane >> 2, tig = lane & 3;
  const int r = tid;
#pragma unroll 1
  for (int pb = 0; pb < GPBA_NB / 8; ++pb) {
    const int c0 = 8 * pb;
    if (tid < 64) {
      const bool own = r >= c0 && r < GPBA_NB;
      double p[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) p[i] = own ? S[r][c0 + i] : 0.0;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int col = c0 + i;
        if (i > 0) {
          if (own && r >= col) S[r][col] = p[i];  // publish the updated, still unscaled column
          bar64();
        }
        const double d = S[col][col];
        const double inv = rsqrt(d);
        if (r == col) {
          if (!(d > 0.0)) atomicExch(fail, 1);
          Sinv[col] = inv;
        }
        if (own && r >= col) p[i] = (r == col) ? d * inv : p[i] * inv;
#pragma unroll
        for (int i2 = i + 1; i2 < 8; ++i2)
          if (own && r >= c0 + i2) p[i2] = fma(-p[i], S[c0 + i2][col] * inv, p[i2]);
      }
      bar64();  // every thread has read the unscaled diagonal before the scaled panel replaces it
      if (own) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
          if (r >= c0 + i) S[r][c0 + i] = p[i];
      }
    }
    __syncthreads();
    // trailing update S[mt][nt] -= P[mt] P[nt]^T over the tiles pb < nt <= mt < 6
    const int m = GPBA_NB / 8 - 1 - pb, cnt = m * (m + 1) / 2;
    for (int t = warp; t < cnt; t += GPBA_PANEL_THREADS / 32) {
      int a = 0;
      while ((a + 1) * (a + 2) / 2 <= t) ++a;
      const int mt = pb + 1 + a, nt = pb + 1 + (t - a * (a + 1) / 2);
      double* cp = &S[8 * mt + gid][8 * nt + 2 * tig];
      double2 c = *reinterpret_cast<double2*>(cp);
#pragma unroll
      for (int kk = 0; kk < 2; ++kk)
        dmma884(c.x, c.y, -S[8 * mt + gid][c0 + 4 * kk + tig], S[8 * nt + gid][c0 + 4 * kk + tig]);
      *reinterpret_cast<double2*>(cp) = c;
    }
    // inverse of the 8x8 diagonal block (column t per thread) for the blocked triangular solves
    if (tid >= GPBA_PANEL_THREADS - 8) {
      const int t = tid - (GPBA_PANEL_THREADS - 8);
      double x[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) x[q] = (q == t) ? 1.0 : 0.0;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        x[q] *= Sinv[c0 + q];
#pragma unroll
        for (int q2 = q + 1; q2 < 8; ++q2) x[q2] = fma(-S[c0 + q2][c0 + q], x[q], x[q2]);
      }
#pragma unroll
      for (int q = 0; q < 8; ++q) D8[pb][q][t] = x[q];
    }
    __syncthreads();
  }
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

// Panel step of tile column k.  grid = 1 + (#non-zero tiles below the diagonal).
__global__ void __launch_bounds__(GPBA_PANEL_THREADS) k_chol_panel(CholView C, int k, int* __restrict__ fail) {
  __shared__ __align__(16) double S[GPBA_NB][GPBA_LD];
  __shared__ __align__(16) double T[GPBA_NB][GPBA_LD];
  __shared__ double D8[GPBA_NB / 8][8][8];
  __shared__ double Sinv[GPBA_NB], Sy[GPBA_NB];
  const int tid = threadIdx.x;
  tile_to_smem(C.tiles + C.tile_off[(size_t)k * C.NT + k], S);
  double* A = nullptr;
  if (blockIdx.x == 0) {
    for (int j = tid; j < GPBA_NB * GPBA_NB; j += blockDim.x) T[j / GPBA_NB][j % GPBA_NB] = (j / GPBA_NB == j % GPBA_NB) ? 1.0 : 0.0;
    if (tid < GPBA_NB) Sy[tid] = C.work[k * GPBA_NB + tid];
  } else {
    A = C.tiles + C.tile_off[(size_t)C.col_rows[C.col_begin[k] + blockIdx.x - 1] * C.NT + k];
    tile_to_smem(A, T);
  }
  __syncthreads();
  potrf48(S, D8, Sinv, fail);
  trsm48(T, S, D8);
  __syncthreads();
  if (blockIdx.x == 0) {
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
    double2* A2 = reinterpret_cast<double2*>(A);
    for (int j = tid; j < GPBA_NB * GPBA_NB / 2; j += blockDim.x)
      A2[j] = *reinterpret_cast<double2*>(&T[j / (GPBA_NB / 2)][2 * (j % (GPBA_NB / 2))]);
  }
}

// Trailing update of step k: A_ab -= L_ak L_bk^T for all pairs a >= b of column k's non-zero rows.
// One CTA (4 warps) per pair; both L tiles staged in shared memory, the C tile prefetched into registers while
// they arrive; each warp owns a 24 x 24 corner = 3 x 3 DMMA tiles (9 independent accumulator chains).
// The extra last CTA applies column k to the right-hand side: b_i -= L_ik y_k.
__global__ void __launch_bounds__(128) k_chol_update(CholView C, int k) {
  __shared__ __align__(16) double La[GPBA_NB][GPBA_LD], Lb[GPBA_NB][GPBA_LD];
  const int cb = C.col_begin[k];
  const int nr = C.col_begin[k + 1] - cb;
  if ((int)blockIdx.x == nr * (nr + 1) / 2) {
    double* yk = &La[0][0];
    if (threadIdx.x < GPBA_NB) yk[threadIdx.x] = C.work[k * GPBA_NB + threadIdx.x];
    __syncthreads();
    for (int q = threadIdx.x; q < nr * GPBA_NB; q += blockDim.x) {
      const int i = C.col_rows[cb + q / GPBA_NB], r = q % GPBA_NB;
      const double2* Lik = reinterpret_cast<const double2*>(C.tiles + C.tile_off[(size_t)i * C.NT + k] + r * GPBA_NB);
      double s0 = 0.0, s1 = 0.0;
#pragma unroll 8
      for (int c = 0; c < GPBA_NB / 2; ++c) {
        const double2 l = Lik[c];
        s0 = fma(l.x, yk[2 * c], s0); s1 = fma(l.y, yk[2 * c + 1], s1);
      }
      C.work[i * GPBA_NB + r] -= s0 + s1;
    }
    return;
  }
  // decode the triangular pair index: blockIdx.x = a*(a+1)/2 + b, a >= b
  int a = (int)((sqrt(8.0 * (double)blockIdx.x + 1.0) - 1.0) * 0.5);
  while ((a + 1) * (a + 2) / 2 <= (int)blockIdx.x) ++a;
  while (a * (a + 1) / 2 > (int)blockIdx.x) --a;
  const int b = blockIdx.x - a * (a + 1) / 2;
  const int ra = C.col_rows[cb + a], rb = C.col_rows[cb + b];
  const double* Ta = C.tiles + C.tile_off[(size_t)ra * C.NT + k];
  const double* Tb = C.tiles + C.tile_off[(size_t)rb * C.NT + k];
  double* Tc = C.tiles + C.tile_off[(size_t)ra * C.NT + rb];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gid = lane >> 2, tig = lane & 3;
  const int m0 = 3 * (warp >> 1), n0 = 3 * (warp & 1);
  tile_to_smem(Ta, La);
  tile_to_smem(Tb, Lb);
  double2 cin[3][3];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      cin[i][j] = *reinterpret_cast<const double2*>(Tc + (8 * (m0 + i) + gid) * GPBA_NB + 8 * (n0 + j) + 2 * tig);
  __syncthreads();
  double2 acc[3][3];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) acc[i][j] = make_double2(0.0, 0.0);
#pragma unroll 2
  for (int k0 = 0; k0 < GPBA_NB; k0 += 4) {
    double af[3], bf[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) { af[i] = La[8 * (m0 + i) + gid][k0 + tig]; bf[i] = Lb[8 * (n0 + i) + gid][k0 + tig]; }
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) dmma884(acc[i][j].x, acc[i][j].y, af[i], bf[j]);
  }
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      double2 o = cin[i][j];
      o.x -= acc[i][j].x; o.y -= acc[i][j].y;
      *reinterpret_cast<double2*>(Tc + (8 * (m0 + i) + gid) * GPBA_NB + 8 * (n0 + j) + 2 * tig) = o;
    }
}

// Backward substitution, tile row i (launched for i = NT-1 .. 0): every CTA computes x_i = L_ii^-T y_i
// (y_i has received the contributions of all later rows); CTA 0 stores it, CTA q > 0 applies one tile of
// row i to an earlier segment: y_k -= L_ik^T x_i.  Distinct k per CTA and rows are stream ordered, so the
// result is deterministic without atomics.
__global__ void __launch_bounds__(192) k_chol_back(CholView C, int i) {
  __shared__ double yi[GPBA_NB], xi[GPBA_NB], part[4][GPBA_NB];
  const int tid = threadIdx.x, c = tid % GPBA_NB, h = tid / GPBA_NB;  // 4 row-quarters x 48 columns
  if (tid < GPBA_NB) yi[tid] = C.work[i * GPBA_NB + tid];
  __syncthreads();
  {
    const double* D = C.dinv + (size_t)i * GPBA_NB * GPBA_NB;  // (L^-T)[c][r] = Linv[r][c], zero for r < c
    double s = 0.0;
#pragma unroll
    for (int r = 12 * h; r < 12 * h + 12; ++r) s = fma(D[r * GPBA_NB + c], yi[r], s);
    part[h][c] = s;
  }
  __syncthreads();
  if (tid < GPBA_NB) {
    const double x = (part[0][tid] + part[1][tid]) + (part[2][tid] + part[3][tid]);
    xi[tid] = x;
    if (blockIdx.x == 0) C.xsol[i * GPBA_NB + tid] = x;  // not in place: the other CTAs of this launch still read y_i
  }
  if (blockIdx.x == 0) return;
  __syncthreads();
  const int k = C.row_cols[C.row_begin[i] + blockIdx.x - 1];
  const double* L = C.tiles + C.tile_off[(size_t)i * C.NT + k];
  double s = 0.0;
#pragma unroll
  for (int r = 12 * h; r < 12 * h + 12; ++r) s = fma(L[r * GPBA_NB + c], xi[r], s);
  part[h][c] = s;
  __syncthreads();
  if (tid < GPBA_NB) C.work[k * GPBA_NB + tid] -= (part[0][tid] + part[1][tid]) + (part[2][tid] + part[3][tid]);
}

__global__ void k_chol_unpermute(CholView C, double* __restrict__ xout) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < C.n) xout[j] = C.xsol[C.perm[j / 12] * 12 + j % 12];
}

}  // namespace gpba
