// gpba_chol.cuh -- K5b: tile-sparse blocked FP64 Cholesky of the reduced camera system.
//
// Replaces LinearSolverDense::solve (Eigen::LDLT, g2o/solvers/linear_solver_dense.h:65-113) and, for
// global BA, LinearSolverEigen::solve (SimplicialLDLT, linear_solver_eigen.h:94-124).  The matrix is
// stored as NB x NB tiles of the lower triangle; only tiles that are structurally non-zero after a
// tile-level symbolic factorization (host, once per buildStructure -- the analogue of SimplicialLDLT's
// analyzePattern, linear_solver_eigen.h:147-201) are allocated and visited, so a banded covisibility
// pattern costs O(n b^2) while a fully dense system degenerates to the classic right-looking blocked
// algorithm.  Trailing updates run on the FP64 tensor pipe (mma.sync.m8n8k4.f64 -> DMMA).
// A non-positive pivot sets *fail (=> solve() returns false => LM rejects the trial, like
// !_cholesky.isPositive(), linear_solver_dense.h:108-112).
#pragma once
#include "gpba_kernels.cuh"

namespace gpba {

#define GPBA_NB 48
#define GPBA_NBP 49  // padded shared-memory row stride

struct CholView {
  int NT;                       // tiles per side
  int n;                        // true dimension (12 * n_pose)
  const int64_t* tile_off;      // [NT*NT] offset (doubles) of tile (i,j), i>=j, or -1
  const int* col_begin;         // [NT+1] into col_rows
  const int* col_rows;          // rows i>k with L_ik != 0, ascending
  double* tiles;
  const int* perm;              // [n/12] fill-reducing order of the pose blocks: permuted index of block b
};

// zero the allocated tiles and put 1 on the padded part of the diagonal
__global__ void k_chol_clear(CholView C, int64_t n_doubles) {
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n_doubles; j += (int64_t)gridDim.x * blockDim.x)
    C.tiles[j] = 0.0;
}
__global__ void k_chol_pad(CholView C) {
  const int r = C.n + blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= C.NT * GPBA_NB) return;
  const int t = r / GPBA_NB, o = r % GPBA_NB;
  C.tiles[C.tile_off[(size_t)t * C.NT + t] + o * GPBA_NB + o] = 1.0;
}
// scatter the upper Hschur blocks (row-major 12x12, block (bi,bj), bi<=bj) into the lower tiles
__global__ void k_chol_scatter(CholView C, int n_hs, const int* __restrict__ hs_row, const int* __restrict__ hs_col,
                               const double* __restrict__ hs) {
  const int64_t n = (int64_t)n_hs * 144;
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x) {
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
}

// Cholesky of the diagonal tile (k,k): one CTA, tile in shared memory.
__global__ void __launch_bounds__(256) k_chol_potrf(CholView C, int k, int* __restrict__ fail) {
  __shared__ double A[GPBA_NB][GPBA_NBP];
  double* T = C.tiles + C.tile_off[(size_t)k * C.NT + k];
  const int tid = threadIdx.x;
  for (int j = tid; j < GPBA_NB * GPBA_NB; j += blockDim.x) A[j / GPBA_NB][j % GPBA_NB] = T[j];
  __syncthreads();
  for (int j = 0; j < GPBA_NB; ++j) {
    const double d = A[j][j];
    if (!(d > 0.0)) { if (tid == 0) atomicExch(fail, 1); }
    const double l = sqrt(d);
    __syncthreads();
    if (tid == 0) A[j][j] = l;
    for (int i = j + 1 + tid; i < GPBA_NB; i += blockDim.x) A[i][j] = A[i][j] / l;
    __syncthreads();
    // trailing update of the lower triangle
    const int rem = GPBA_NB - j - 1;
    for (int q = tid; q < rem * rem; q += blockDim.x) {
      const int r = j + 1 + q / rem, c = j + 1 + q % rem;
      if (c <= r) A[r][c] = fma(-A[r][j], A[c][j], A[r][c]);
    }
    __syncthreads();
  }
  for (int j = tid; j < GPBA_NB * GPBA_NB; j += blockDim.x) {
    const int r = j / GPBA_NB, c = j % GPBA_NB;
    T[j] = c <= r ? A[r][c] : 0.0;
  }
}

// L_ik = A_ik L_kk^-T for every non-zero tile below the diagonal of column k: one CTA per tile, one thread per row.
__global__ void __launch_bounds__(GPBA_NB) k_chol_trsm(CholView C, int k) {
  __shared__ double L[GPBA_NB][GPBA_NBP];
  const int i = C.col_rows[C.col_begin[k] + blockIdx.x];
  const double* Lkk = C.tiles + C.tile_off[(size_t)k * C.NT + k];
  double* A = C.tiles + C.tile_off[(size_t)i * C.NT + k];
  for (int j = threadIdx.x; j < GPBA_NB * GPBA_NB; j += blockDim.x) L[j / GPBA_NB][j % GPBA_NB] = Lkk[j];
  __syncthreads();
  const int r = threadIdx.x;
  double x[GPBA_NB];
#pragma unroll
  for (int c = 0; c < GPBA_NB; ++c) x[c] = A[r * GPBA_NB + c];
#pragma unroll
  for (int c = 0; c < GPBA_NB; ++c) {
    double s = x[c];
#pragma unroll
    for (int p = 0; p < c; ++p) s = fma(-x[p], L[c][p], s);
    x[c] = s / L[c][c];
  }
#pragma unroll
  for (int c = 0; c < GPBA_NB; ++c) A[r * GPBA_NB + c] = x[c];
}

// Trailing update of step k: A_ab -= L_ak L_bk^T for all pairs a >= b of column k's non-zero rows.
// One CTA (4 warps) per pair; both L tiles staged in shared memory; 6x6 DMMA output tiles of 8x8.
__global__ void __launch_bounds__(128) k_chol_update(CholView C, int k) {
  __shared__ double La[GPBA_NB][GPBA_NBP], Lb[GPBA_NB][GPBA_NBP];
  const int cb = C.col_begin[k];
  // decode the triangular pair index: blockIdx.x = a*(a+1)/2 + b, a >= b
  int a = (int)((sqrt(8.0 * (double)blockIdx.x + 1.0) - 1.0) * 0.5);
  while ((a + 1) * (a + 2) / 2 <= (int)blockIdx.x) ++a;
  while (a * (a + 1) / 2 > (int)blockIdx.x) --a;
  const int b = blockIdx.x - a * (a + 1) / 2;
  const int ra = C.col_rows[cb + a], rb = C.col_rows[cb + b];
  const double* Ta = C.tiles + C.tile_off[(size_t)ra * C.NT + k];
  const double* Tb = C.tiles + C.tile_off[(size_t)rb * C.NT + k];
  double* Tc = C.tiles + C.tile_off[(size_t)ra * C.NT + rb];
  for (int j = threadIdx.x; j < GPBA_NB * GPBA_NB; j += blockDim.x) {
    La[j / GPBA_NB][j % GPBA_NB] = Ta[j];
    Lb[j / GPBA_NB][j % GPBA_NB] = Tb[j];
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gid = lane >> 2, tig = lane & 3;
  for (int t = warp; t < 36; t += 4) {  // 6 x 6 output tiles
    const int mt = t / 6, nt = t % 6;
    double c0 = 0.0, c1 = 0.0;
#pragma unroll
    for (int k0 = 0; k0 < GPBA_NB; k0 += 4) dmma884(c0, c1, La[mt * 8 + gid][k0 + tig], Lb[nt * 8 + gid][k0 + tig]);
    double* out = Tc + (mt * 8 + gid) * GPBA_NB + nt * 8 + 2 * tig;
    out[0] -= c0;
    out[1] -= c1;
  }
}

// Forward / backward substitution with the tile factor, one CTA (sequential over tile columns).
// rhs (length n) -> x (length n).
__global__ void __launch_bounds__(256) k_chol_solve(CholView C, const double* __restrict__ rhs, double* __restrict__ x,
                                                    double* __restrict__ work /* NT*NB */) {
  __shared__ double yk[GPBA_NB];
  __shared__ double Ld[GPBA_NB][GPBA_NBP];
  const int tid = threadIdx.x, NTNB = C.NT * GPBA_NB;
  for (int j = tid; j < NTNB; j += blockDim.x) work[j] = 0.0;
  __syncthreads();
  for (int j = tid; j < C.n; j += blockDim.x) work[C.perm[j / 12] * 12 + j % 12] = rhs[j];
  __syncthreads();
  for (int k = 0; k < C.NT; ++k) {  // L y = b
    const double* Lkk = C.tiles + C.tile_off[(size_t)k * C.NT + k];
    for (int j = tid; j < GPBA_NB * GPBA_NB; j += blockDim.x) Ld[j / GPBA_NB][j % GPBA_NB] = Lkk[j];
    if (tid < GPBA_NB) yk[tid] = work[k * GPBA_NB + tid];
    __syncthreads();
    if (tid < 32) {  // one warp: 48-dim forward substitution
      for (int c = 0; c < GPBA_NB; ++c) {
        const double yc = yk[c] / Ld[c][c];
        __syncwarp();
        if (tid == 0) yk[c] = yc;
        for (int r = c + 1 + tid; r < GPBA_NB; r += 32) yk[r] = fma(-Ld[r][c], yc, yk[r]);
        __syncwarp();
      }
    }
    __syncthreads();
    if (tid < GPBA_NB) work[k * GPBA_NB + tid] = yk[tid];
    const int nb = C.col_begin[k + 1] - C.col_begin[k];
    for (int q = tid; q < nb * GPBA_NB; q += blockDim.x) {
      const int i = C.col_rows[C.col_begin[k] + q / GPBA_NB], r = q % GPBA_NB;
      const double* Lik = C.tiles + C.tile_off[(size_t)i * C.NT + k] + r * GPBA_NB;
      double s = 0.0;
#pragma unroll 8
      for (int c = 0; c < GPBA_NB; ++c) s = fma(Lik[c], yk[c], s);
      work[i * GPBA_NB + r] -= s;
    }
    __syncthreads();
  }
  for (int k = C.NT - 1; k >= 0; --k) {  // L^T x = y
    const int nb = C.col_begin[k + 1] - C.col_begin[k];
    if (tid < GPBA_NB) yk[tid] = work[k * GPBA_NB + tid];
    __syncthreads();
    // yk[c] -= sum_i sum_r L_ik[r][c] x_i[r] : thread c handles column c (strided over the tile rows)
    if (tid < GPBA_NB) {
      double s = 0.0;
      for (int q = 0; q < nb; ++q) {
        const int i = C.col_rows[C.col_begin[k] + q];
        const double* Lik = C.tiles + C.tile_off[(size_t)i * C.NT + k];
        const double* xi = work + i * GPBA_NB;
#pragma unroll 8
        for (int r = 0; r < GPBA_NB; ++r) s = fma(Lik[r * GPBA_NB + tid], xi[r], s);
      }
      yk[tid] -= s;
    }
    const double* Lkk = C.tiles + C.tile_off[(size_t)k * C.NT + k];
    for (int j = tid; j < GPBA_NB * GPBA_NB; j += blockDim.x) Ld[j / GPBA_NB][j % GPBA_NB] = Lkk[j];
    __syncthreads();
    if (tid < 32) {
      for (int c = GPBA_NB - 1; c >= 0; --c) {
        const double xc = yk[c] / Ld[c][c];
        __syncwarp();
        if (tid == 0) yk[c] = xc;
        for (int r = tid; r < c; r += 32) yk[r] = fma(-Ld[c][r], xc, yk[r]);
        __syncwarp();
      }
    }
    __syncthreads();
    if (tid < GPBA_NB) work[k * GPBA_NB + tid] = yk[tid];
    __syncthreads();
  }
  for (int j = tid; j < C.n; j += blockDim.x) x[j] = work[C.perm[j / 12] * 12 + j % 12];
}

}  // namespace gpba
