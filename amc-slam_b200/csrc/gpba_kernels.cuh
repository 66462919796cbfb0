// gpba_kernels.cuh -- hot-path kernels of libgpba (sm_100a), K0..K8 of SURVEY.md §2.3.
//
//  K0  k_records          GaussianProcess::QueryPose + the pose-chain part of EdgeMonoGP::linearizeOplus
//                         (src/GaussianProcess.cc:23-42, src/G2oTypes.cc:343-357) once per (KF,cam) record
//  K1  k_residual         SparseOptimizer::computeActiveErrors + activeRobustChi2 (sparse_optimizer.cpp:61-114)
//  K2  k_lin_points       BlockSolver::buildSystem, landmark side: Hll, b_l and W_o = J1^T (rho' Omega) J_p (6x3 per
//                         observation; Hpl_(pose,l) = sum_o M_r^T W_o is never materialised) (block_solver.hpp:502-560)
//      k_lin_records      ... pose side in the 6-dim record tangent space: S_r = sum w J1^T J1, g_r
//      k_rec_to_hpp       ... Hpp += M_r^T S_r M_r, b_p += M_r^T g_r
//  K3  k_priors           EdgeGaussianPrior / EdgeVelocity (G2oTypes.cc:100-118, G2oTypes.h:155-163,496-519)
//  K4  k_schur_prep       D = Hll + lambda I = L L^T, U_o = W_o L^-T, z = L^-1 b_l
//      k_schur_pairs      C_(r,r') = sum over observation pairs of a landmark of U_o U_o'^T (6x6 per record pair, DMMA)
//      k_schur_expand     Hschur_ij = Hpp_ij + lambda - sum M_r^T C_(r,r') M_r', bschur  (block_solver.hpp:367-439)
//  K6  k_rec_y / k_backsub / k_update_poses   landmark back-substitution + oplus (block_solver.hpp:459-483,
//                         sparse_optimizer.cpp:422-435, G2oTypes.cc:41-46)
// The Schur complement is evaluated in the record tangent space: every reprojection edge of a
// (KF_prev, KF_cur, camera) record shares the 6x24 chain matrix M_r, so Hpl D^-1 Hpl^T = sum M_r^T (W_o D^-1 W_o'^T) M_r'
// needs a 6x6x3 product per observation pair instead of a 12x12x3 one per pose pair (about 7x fewer flops and 3.5x
// less traffic at 10 observations per landmark); only the summation order differs from the reference.
//  K8  k_flags            LocalGPBA inlier check (src/Optimizer.cc:1263-1348)
#pragma once
#include "gpba_device.cuh"

namespace gpba {

// ------------------------------------------------------------------------------------------------ helpers
GPBA_D double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// block-wide sum to thread 0 (blockDim.x multiple of 32, <= 1024)
GPBA_D double block_sum(double v, double* smem32) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) smem32[w] = v;
  __syncthreads();
  double r = 0.0;
  if (w == 0) {
    r = lane < (int)((blockDim.x + 31) >> 5) ? smem32[lane] : 0.0;
    r = warp_sum(r);
  }
  __syncthreads();
  return r;
}

// R_cw | t_cw of a record row (12 doubles, 16-byte aligned rows): six 16-byte loads instead of twelve scattered 8-byte
// ones -- neighbouring observations belong to different records, so every load instruction of a warp touches up to
// 32 lines and the load count, not the byte count, is what these kernels pay for.
GPBA_D void load_rec12(const double* __restrict__ p, double (&R)[12]) {
  const double2* q = reinterpret_cast<const double2*>(p);
#pragma unroll
  for (int k = 0; k < 6; ++k) { const double2 v = __ldg(q + k); R[2 * k] = v.x; R[2 * k + 1] = v.y; }
}

// ---- TMA bulk copies (cp.async.bulk, SASS UBLKCP) completing on an mbarrier
GPBA_D unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
GPBA_D void mbar_init(unsigned long long* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
GPBA_D void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
GPBA_D void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
GPBA_D void mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
GPBA_D void mbar_wait(unsigned long long* bar, unsigned parity) {
  unsigned ok;
  do {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  } while (!ok);
}
// global -> shared, `bytes` a multiple of 16, both addresses 16-byte aligned; completes `bytes` on the barrier
GPBA_D void bulk_g2s(void* dst_smem, const void* src_global, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst_smem)), "l"(src_global), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// Record window of an observation tile: the tile's observations belong to landmarks first seen by neighbouring
// keyframes, so the records they touch form a short contiguous run of the (keyframe, camera)-ordered record table
// (C4: ~126 rows per 1024 observations).  One thread stages the run with a single cp.async.bulk (TMA) while the others
// fetch their observation scalars; afterwards a row is 6 LDS.128 away instead of 6 scattered LDG.128 (ncu, round 1 and
// profiles/r02_ncu_streaming.txt: these kernels were bound by L1 tag throughput, l1tex 89-90 %, DRAM 20-45 %).
// Observations of a revisited place (records far from the run) fall back to the global row.
struct RecWindow {
  const double* smem;   // staged rows (12 doubles each)
  const double* global; // the lite record table
  int lo, cnt;
  GPBA_D void load12(int r, double (&R)[12]) const {
    const unsigned d = (unsigned)(r - lo);
    const double2* q = reinterpret_cast<const double2*>(d < (unsigned)cnt ? smem + (size_t)d * 12 : global + (size_t)r * GPBA_REC_LITE_STRIDE);
    if (d < (unsigned)cnt) {
#pragma unroll
      for (int k = 0; k < 6; ++k) { const double2 v = q[k]; R[2 * k] = v.x; R[2 * k + 1] = v.y; }
    } else {
#pragma unroll
      for (int k = 0; k < 6; ++k) { const double2 v = __ldg(q + k); R[2 * k] = v.x; R[2 * k + 1] = v.y; }
    }
  }
};

// Per-observation geometry shared by K1/K2/K8: residual, chi2, robust weight, J1 (rows x 6), Jp (rows x 3).
template <bool STEREO>
struct ObsEval {
  double e[STEREO ? 3 : 2];
  double chi2, rho, rho1;
  int rows;
};

template <bool STEREO, bool JAC>
GPBA_D void eval_obs(const DevView& V, const double* __restrict__ R /* R_cw[9], t_cw[3] */, const CamConst& cam,
                     double X0, double X1, double X2, double u, double v, double ur, double w, unsigned flags,
                     ObsEval<STEREO>& E, double (*J1)[6], double (*Jp)[3]) {
  const double xc = fma(R[0], X0, fma(R[1], X1, fma(R[2], X2, R[9])));
  const double yc = fma(R[3], X0, fma(R[4], X1, fma(R[5], X2, R[10])));
  const double zc = fma(R[6], X0, fma(R[7], X1, fma(R[8], X2, R[11])));
  const double pu = cam.fx * xc / zc + cam.cx;  // Pinhole::project (src/CameraModels/Pinhole.cpp:35-41)
  const double pv = cam.fy * yc / zc + cam.cy;
  E.e[0] = u - pu;
  E.e[1] = v - pv;
  double chi2 = E.e[0] * (w * E.e[0]) + E.e[1] * (w * E.e[1]);  // BaseEdge::chi2, Omega = I * invSigma2
  bool stereo = false;
  if (STEREO) {
    stereo = ur >= 0.0;
    if (stereo) {
      const double invz = 1 / zc;
      E.e[2] = ur - (pu - V.bf * invz);  // EdgeStereoGP::computeError (src/G2oTypes.cc:380-385)
      chi2 += E.e[2] * (w * E.e[2]);
    } else {
      E.e[2] = 0.0;
    }
  }
  E.rows = stereo ? 3 : 2;
  E.chi2 = chi2;
  const double delta = stereo ? V.hub_stereo_delta : V.hub_mono_delta;
  const double dsqr = stereo ? V.hub_stereo_dsqr : V.hub_mono_dsqr;
  if (delta > 0.0 && !(flags & 0x4u)) {
    E.rho = huber(chi2, delta, dsqr, &E.rho1);
  } else {
    E.rho = chi2;
    E.rho1 = 1.0;
  }
  if (JAC) {
    // proj_jac (Pinhole.cpp:71-81), stereo row (G2oTypes.cc:407-412)
    double P[STEREO ? 3 : 2][3];
    P[0][0] = cam.fx / zc; P[0][1] = 0.0; P[0][2] = -cam.fx * xc / (zc * zc);
    P[1][0] = 0.0; P[1][1] = cam.fy / zc; P[1][2] = -cam.fy * yc / (zc * zc);
    if (STEREO) {
      P[2][0] = stereo ? P[0][0] : 0.0; P[2][1] = 0.0;
      P[2][2] = stereo ? P[0][2] + V.bf * (1.0 / (zc * zc)) : 0.0;
    }
    // X_b = T_bc X_c
    const double xb = fma(cam.Rbc[0], xc, fma(cam.Rbc[1], yc, fma(cam.Rbc[2], zc, cam.tbc[0])));
    const double yb = fma(cam.Rbc[3], xc, fma(cam.Rbc[4], yc, fma(cam.Rbc[5], zc, cam.tbc[1])));
    const double zb = fma(cam.Rbc[6], xc, fma(cam.Rbc[7], yc, fma(cam.Rbc[8], zc, cam.tbc[2])));
#pragma unroll
    for (int r = 0; r < (STEREO ? 3 : 2); ++r) {
      // G = P * R_cb ; J1 = -P [-R_cb, R_cb X_b^] = [G, -G X_b^]   (G2oTypes.cc:337-341)
      const double g0 = P[r][0] * cam.Rcb[0] + P[r][1] * cam.Rcb[3] + P[r][2] * cam.Rcb[6];
      const double g1 = P[r][0] * cam.Rcb[1] + P[r][1] * cam.Rcb[4] + P[r][2] * cam.Rcb[7];
      const double g2 = P[r][0] * cam.Rcb[2] + P[r][1] * cam.Rcb[5] + P[r][2] * cam.Rcb[8];
      J1[r][0] = g0; J1[r][1] = g1; J1[r][2] = g2;
      // (G X_b^)[j]: X_b^ = [[0,-zb,yb],[zb,0,-xb],[-yb,xb,0]]
      J1[r][3] = -(g1 * zb - g2 * yb);
      J1[r][4] = -(-g0 * zb + g2 * xb);
      J1[r][5] = -(g0 * yb - g1 * xb);
      // J_point = -P R_cb R_bw = -P R_cw   (G2oTypes.cc:366)
      Jp[r][0] = -(P[r][0] * R[0] + P[r][1] * R[3] + P[r][2] * R[6]);
      Jp[r][1] = -(P[r][0] * R[1] + P[r][1] * R[4] + P[r][2] * R[7]);
      Jp[r][2] = -(P[r][0] * R[2] + P[r][1] * R[5] + P[r][2] * R[8]);
    }
  }
}

// ------------------------------------------------------------------------------------------------ K0
// One record row: interpolated camera pose (R_cw | t_cw) and, if FULL, the 6 x 24 chain matrix M.
// pose1 / vel1 == nullptr marks a synchronous record.
template <bool FULL, int MS = GPBA_REC_MS>
GPBA_D void record_row(const double* __restrict__ pose1, const double* __restrict__ vel1, double t1,
                       const double* __restrict__ pose2, const double* __restrict__ vel2, double t2, double t,
                       const CamConst& cam, double* __restrict__ out) {
  const SE3 T2 = load_se3(pose2);
  SE3 Twb = T2;
  if (pose1) {
    const SE3 T1 = load_se3(pose1);
    const V6 v1 = load_v6(vel1), v2 = load_v6(vel2);
    const GpWeights gw = gp_weights(t1, t2, t);
    const V6 xi12 = se3_log(se3_mul(se3_inv(T1), T2));
    const M6 K = RightJacobianPose3Inv(xi12);
    const V6 Kv2 = mul(K, v2);
    V6 arg;
#pragma unroll
    for (int i = 0; i < 6; ++i) arg[i] = gw.l12 * v1[i] + gw.p11 * xi12[i] + gw.p12 * Kv2[i];
    const SE3 dT = se3_exp(arg);
    Twb = se3_mul(T1, dT);
    if (FULL) {
      const V6 dxi = se3_log(dT);
      const M6 Ad = se3_Adj(se3_exp(neg6(dxi)));
      const M6 Jd = RightJacobianPose3(dxi);
      const M6 a = se3Adj(v2);
      const M6 A12inv = se3_Adj(se3_inv(se3_exp(xi12)));  // Adj(T)^-1 == Adj(T^-1)
      const M6 Jt = scale(-1.0, mul(K, A12inv));          // JinT1 top   (G2oTypes.cc:352)
      const M6 haJt = mul(scale(-0.5, a), Jt);            // JinT1 bottom (:353)
      const M6 haK = mul(scale(-0.5, a), K);              // JinT2 bottom (:356)
      const M6 MT1 = add(mul(Jd, add(scale(gw.p11, Jt), scale(gw.p12, haJt))), Ad);
      const M6 MV1 = scale(gw.l12, Jd);
      const M6 MT2 = mul(Jd, add(scale(gw.p11, K), scale(gw.p12, haK)));
      const M6 MV2 = scale(gw.p12, mul(Jd, K));
      double* M = out + GPBA_REC_M;
#pragma unroll
      for (int m = 0; m < 6; ++m)
#pragma unroll
        for (int c = 0; c < 6; ++c) {
          M[m * MS + c] = MT1(m, c);
          M[m * MS + 6 + c] = MV1(m, c);
          M[m * MS + 12 + c] = MT2(m, c);
          M[m * MS + 18 + c] = MV2(m, c);
        }
    }
  } else if (FULL) {
    // synchronous record (EdgeMono / EdgeStereo): pose block = J1, velocity block = 0 (G2oTypes.cc:465-467),
    // i.e. M = [0 | 0 | I | 0], so that every consumer can treat the two record kinds alike
    double* M = out + GPBA_REC_M;
#pragma unroll
    for (int m = 0; m < 6; ++m)
#pragma unroll
      for (int c = 0; c < 24; ++c) M[m * MS + c] = (c == 12 + m) ? 1.0 : 0.0;
  }
  SE3 Tbc;
  Tbc.q.x = cam.qbc[0]; Tbc.q.y = cam.qbc[1]; Tbc.q.z = cam.qbc[2]; Tbc.q.w = cam.qbc[3];
  Tbc.t = v3(cam.tbc[0], cam.tbc[1], cam.tbc[2]);
  const SE3 Tcw = se3_inv(se3_mul(Twb, Tbc));
  const M3 R = quat_to_R(Tcw.q);
#pragma unroll
  for (int i = 0; i < 9; ++i) out[i] = R.a[i];
  out[9] = Tcw.t[0]; out[10] = Tcw.t[1]; out[11] = Tcw.t[2];
}

template <bool FULL>
__global__ void k_records(DevView V, const double* __restrict__ pose, const double* __restrict__ vel,
                          double* __restrict__ rec_out, double* __restrict__ lite_out = nullptr) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= V.n_rec) return;
  const int k1 = V.rec_kf1[r], k2 = V.rec_kf2[r];
  double* out = rec_out + (size_t)r * (FULL ? GPBA_REC_STRIDE : GPBA_REC_LITE_STRIDE);
  const CamConst& cam = V.cam[V.rec_cam[r]];
  record_row<FULL>(k1 >= 0 ? pose + 7 * k1 : nullptr, k1 >= 0 ? vel + 6 * k1 : nullptr, k1 >= 0 ? V.kf_time[k1] : 0.0,
                   pose + 7 * k2, vel + 6 * k2, V.kf_time[k2], V.rec_t[r], cam, out);
  if (FULL) {   // third slice: the extrinsic's 12-slot [Adj(T_bc) | 0] (EdgeMonoGPExtrinsic, src/G2oTypes.cc:311-313)
    const bool ext = k1 >= 0 && V.ext_h[V.rec_cam[r]] >= 0;
    double* M = out + GPBA_REC_M;
#pragma unroll
    for (int m = 0; m < 6; ++m)
#pragma unroll
      for (int c = 0; c < 12; ++c) M[m * GPBA_REC_MS + 24 + c] = (ext && c < 6) ? cam.AdjTbc[m * 6 + c] : 0.0;
    if (lite_out) {   // the compact (R_cw | t_cw) table the streaming kernels stage in shared memory
#pragma unroll
      for (int c = 0; c < 12; ++c) lite_out[(size_t)r * GPBA_REC_LITE_STRIDE + c] = out[c];
    }
  }
}

// ------------------------------------------------------------------------------------------------ K1
// One thread per observation (sorted by landmark: u/v/w/rec/lm loads are coalesced, the landmark and record rows come
// from L1/L2).  Writes nothing per observation unless chi2_out / err_out are given.  A TMA-staged variant of this kernel
// (record window per observation tile, as K2a has) was measured and not kept: with only 36 B of HBM traffic per
// observation the per-tile barrier + bulk-copy latency cost as much as the scattered row loads it removed
// (profiles/r02_ncu_streaming.txt).
template <bool STEREO>
__global__ void __launch_bounds__(256) k_residual(DevView V, const double* __restrict__ rec_lite,
                                                  const double* __restrict__ pt, double* __restrict__ partial,
                                                  double* __restrict__ chi2_out, double* __restrict__ err_out = nullptr) {
  __shared__ double red[32];
  double acc = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < V.n_aobs; i += (int64_t)gridDim.x * blockDim.x) {
    const int r = V.o_rec[i];
    const int lm = V.o_lm[i];
    double R[12];
    load_rec12(rec_lite + (size_t)r * GPBA_REC_LITE_STRIDE, R);
    ObsEval<STEREO> E;
    eval_obs<STEREO, false>(V, R, V.cam[V.rec_cam[r]], pt[3 * (size_t)lm], pt[3 * (size_t)lm + 1], pt[3 * (size_t)lm + 2],
                            V.o_u[i], V.o_v[i], STEREO ? V.o_ur[i] : -1.0, V.o_w[i], V.o_flags[i], E, nullptr, nullptr);
    acc += E.rho;
    if (chi2_out) chi2_out[V.o_orig[i]] = E.chi2;
    if (err_out) {   // BaseEdge::_error of the edge (3 slots; the third is 0 for a monocular edge)
      double* eo = err_out + 3 * V.o_orig[i];
      eo[0] = E.e[0]; eo[1] = E.e[1]; eo[2] = (STEREO && E.rows == 3) ? E.e[2] : 0.0;
    }
  }
  const double s = block_sum(acc, red);
  if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

// Sum of `n` partials + `m` prior terms in a fixed order -> out[0] (deterministic, single block).
__global__ void __launch_bounds__(256) k_reduce(const double* __restrict__ a, int n, const double* __restrict__ b, int m,
                                                double* __restrict__ out) {
  __shared__ double red[32];
  double acc = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) acc += a[i];
  for (int i = threadIdx.x; i < m; i += blockDim.x) acc += b[i];
  const double s = block_sum(acc, red);
  if (threadIdx.x == 0) out[0] = s;
}

// ------------------------------------------------------------------------------------------------ K2a
// One thread per observation (observations are sorted by landmark, so a warp covers ~3 landmarks at 10 observations
// each; a warp-per-landmark mapping left 2/3 of the lanes idle).  Residual, robust weight, J1, Jp per lane;
// W_o = J1^T (rho' w) Jp is staged through shared memory and leaves as fully coalesced 256-byte stores; Hll / b_l are
// reduced over the lanes of equal landmark with a segmented shuffle reduction, one atomicAdd per value and segment
// (a landmark that straddles two warps gets two commutative adds into the zeroed accumulators).
#define GPBA_K2_THREADS 128
#define GPBA_K2_WSTRIDE 19   // padded row stride of the staging tile (18 values per observation)
template <bool STEREO>
__global__ void __launch_bounds__(GPBA_K2_THREADS) k_lin_points(DevView V, const double* __restrict__ rec,
                                                                const double* __restrict__ pt, double* __restrict__ hll,
                                                                double* __restrict__ bl, double* __restrict__ W) {
  __shared__ double sW[GPBA_K2_THREADS / 32][32 * GPBA_K2_WSTRIDE];
  __shared__ __align__(16) double win[GPBA_WIN_ROWS * 12];
  __shared__ __align__(8) unsigned long long bar;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int ROWS = STEREO ? 3 : 2;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init_fence(); }
  __syncthreads();
  unsigned phase = 0;
  // persistent CTAs over the landmark-aligned tiles: the tile's record rows are staged by one TMA bulk copy (RecWindow)
  for (int t = blockIdx.x; t < V.n_tiles; t += gridDim.x) {
   const int wlo = V.tile_rlo[t], wcnt = V.tile_rcnt[t];
   if (threadIdx.x == 0) {
     mbar_expect_tx(&bar, (unsigned)wcnt * 96u);
     bulk_g2s(win, rec + (size_t)wlo * GPBA_REC_LITE_STRIDE, (unsigned)wcnt * 96u, &bar);
   }
   const RecWindow RW{win, rec, wlo, wcnt};
   const int64_t t_ob = V.lm_obs_begin[V.tile_lm[t]], t_oe = V.lm_obs_begin[V.tile_lm[t + 1]];
   mbar_wait(&bar, phase);
   phase ^= 1u;
   for (int64_t base = t_ob + warp * 32; base < t_oe; base += GPBA_K2_THREADS) {
    const int64_t i = base + lane;
    const bool live = i < t_oe;
    int lm = -1;
    double h[6] = {0, 0, 0, 0, 0, 0}, b[3] = {0, 0, 0};
    if (live) {
      lm = V.o_lm[i];
      const int r = V.o_rec[i];
      double R[12];
      RW.load12(r, R);
      ObsEval<STEREO> E;
      double J1[ROWS][6], Jp[ROWS][3];
      const double w = V.o_w[i];
      eval_obs<STEREO, true>(V, R, V.cam[V.rec_cam[r]], pt[3 * (size_t)lm], pt[3 * (size_t)lm + 1], pt[3 * (size_t)lm + 2],
                             V.o_u[i], V.o_v[i], STEREO ? V.o_ur[i] : -1.0, w, V.o_flags[i], E, J1, Jp);
      const double wr = E.rho1 * w;  // robustInformation = rho' * Omega (base_edge.h:96-102)
#pragma unroll
      for (int m = 0; m < 6; ++m)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          double s = 0.0;
#pragma unroll
          for (int rr = 0; rr < ROWS; ++rr) s = fma(wr * J1[rr][m], Jp[rr][c], s);
          sW[warp][lane * GPBA_K2_WSTRIDE + m * 3 + c] = s;
        }
#pragma unroll
      for (int rr = 0; rr < ROWS; ++rr) {
        const double w0 = wr * Jp[rr][0], w1 = wr * Jp[rr][1], w2 = wr * Jp[rr][2];
        h[0] = fma(w0, Jp[rr][0], h[0]); h[1] = fma(w0, Jp[rr][1], h[1]); h[2] = fma(w0, Jp[rr][2], h[2]);
        h[3] = fma(w1, Jp[rr][1], h[3]); h[4] = fma(w1, Jp[rr][2], h[4]); h[5] = fma(w2, Jp[rr][2], h[5]);
        b[0] = fma(-w0, E.e[rr], b[0]); b[1] = fma(-w1, E.e[rr], b[1]); b[2] = fma(-w2, E.e[rr], b[2]);
      }
    }
    __syncwarp();
    // coalesced copy-out of the warp's W rows
    const int nlive = (int)((t_oe - base) < 32 ? (t_oe - base) : 32);
    double* out = W + (size_t)base * 18;
    for (int j = lane; j < nlive * 18; j += 32) out[j] = sW[warp][(j / 18) * GPBA_K2_WSTRIDE + j % 18];
    // segmented reduction over equal landmark (segments are contiguous)
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int lm_o = __shfl_down_sync(0xffffffffu, lm, o);
      const bool take = (lane + o < 32) && lm_o == lm;
#pragma unroll
      for (int k = 0; k < 6; ++k) { const double t = __shfl_down_sync(0xffffffffu, h[k], o); if (take) h[k] += t; }
#pragma unroll
      for (int k = 0; k < 3; ++k) { const double t = __shfl_down_sync(0xffffffffu, b[k], o); if (take) b[k] += t; }
    }
    const int lm_prev = __shfl_up_sync(0xffffffffu, lm, 1);
    if (live && (lane == 0 || lm_prev != lm)) {
      double* H = hll + 9 * (size_t)lm;
      atomicAdd(H + 0, h[0]); atomicAdd(H + 1, h[1]); atomicAdd(H + 2, h[2]);
      atomicAdd(H + 3, h[1]); atomicAdd(H + 4, h[3]); atomicAdd(H + 5, h[4]);
      atomicAdd(H + 6, h[2]); atomicAdd(H + 7, h[4]); atomicAdd(H + 8, h[5]);
      atomicAdd(bl + 3 * (size_t)lm, b[0]); atomicAdd(bl + 3 * (size_t)lm + 1, b[1]); atomicAdd(bl + 3 * (size_t)lm + 2, b[2]);
    }
    __syncwarp();
   }
   __syncthreads();   // every warp is done with the window: the next tile may overwrite it
  }
}

// ------------------------------------------------------------------------------------------------ K2b
// One warp per record segment (observations grouped by record).  Accumulates the 6x6 S_r = sum w J1^T J1
// (upper, 21 values) and g_r = -sum w J1^T e in registers, warp-shuffle reduces, one atomicAdd per value.
template <bool STEREO>
__global__ void __launch_bounds__(128, 3) k_lin_records(DevView V, const double* __restrict__ rec,
                                                     const double* __restrict__ pt, double* __restrict__ recS,
                                                     const double* __restrict__ r_u, const double* __restrict__ r_v,
                                                     const double* __restrict__ r_ur, const double* __restrict__ r_w,
                                                     const int* __restrict__ r_lm, const uint8_t* __restrict__ r_flags) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int ROWS = STEREO ? 3 : 2;
  for (int s = blockIdx.x * 4 + warp; s < V.n_rseg; s += gridDim.x * 4) {
    const int r = V.rseg_rec[s];
    double R[12];
    load_rec12(rec + (size_t)r * GPBA_REC_STRIDE, R);
    const CamConst& cam = V.cam[V.rec_cam[r]];
    double acc[27];
#pragma unroll
    for (int k = 0; k < 27; ++k) acc[k] = 0.0;
    // record-major copies of the per-observation inputs (made once per structure): coalesced streams, only the
    // landmark position is gathered
    for (int64_t j = V.rseg_begin[s] + lane; j < V.rseg_begin[s + 1]; j += 32) {
      const int lm = r_lm[j];
      const double w = r_w[j];
      ObsEval<STEREO> E;
      double J1[ROWS][6], Jp[ROWS][3];
      eval_obs<STEREO, true>(V, R, cam, pt[3 * (size_t)lm], pt[3 * (size_t)lm + 1], pt[3 * (size_t)lm + 2], r_u[j],
                             r_v[j], STEREO ? r_ur[j] : -1.0, w, r_flags[j], E, J1, Jp);
      const double wr = E.rho1 * w;
#pragma unroll
      for (int rr = 0; rr < ROWS; ++rr) {
        int k = 0;
#pragma unroll
        for (int m = 0; m < 6; ++m) {
          const double wj = wr * J1[rr][m];
#pragma unroll
          for (int n = m; n < 6; ++n) { acc[k] = fma(wj, J1[rr][n], acc[k]); ++k; }
          acc[21 + m] = fma(-wj, E.e[rr], acc[21 + m]);
        }
      }
    }
#pragma unroll
    for (int k = 0; k < 27; ++k) acc[k] = warp_sum(acc[k]);
    if (lane == 0) {
#pragma unroll
      for (int k = 0; k < 27; ++k) atomicAdd(&recS[(size_t)r * 27 + k], acc[k]);
    }
  }
}

// ------------------------------------------------------------------------------------------------ K2c
// One CTA per record: Hpp blocks += M_r^T S_r M_r, b_p += M_r^T g_r, over the record's (up to) three pose-like vertices:
// previous keyframe, current keyframe, extrinsic (12-column slices 0, 1, 2 of M_r).
__global__ void __launch_bounds__(128) k_rec_to_hpp(DevView V, const double* __restrict__ rec,
                                                    const double* __restrict__ recS, double* __restrict__ hpp,
                                                    double* __restrict__ bp) {
  __shared__ double S[36], g[6], M[6 * GPBA_REC_MS], T[6 * GPBA_REC_MS];
  const int r = blockIdx.x;
  const int k1 = V.rec_kf1[r], k2 = V.rec_kf2[r];
  const int h1 = k1 >= 0 ? V.kf_h[k1] : -1, h2 = V.kf_h[k2];
  const int h3 = k1 >= 0 ? V.ext_h[V.rec_cam[r]] : -1;
  if (h1 < 0 && h2 < 0 && h3 < 0) return;
  const int tid = threadIdx.x;
  if (tid < 27) {
    const double v = recS[(size_t)r * 27 + tid];
    if (tid < 21) {
      int m = 0, k = tid;
      while (k >= 6 - m) { k -= 6 - m; ++m; }
      const int n = m + k;
      S[m * 6 + n] = v; S[n * 6 + m] = v;
    } else {
      g[tid - 21] = v;
    }
  }
  const int b22 = V.rec_hpp22[r];
  if (k1 < 0) {  // synchronous record: pose block = S, velocity rows/cols = 0
    __syncthreads();
    if (h2 < 0) return;
    if (tid < 36) atomicAdd(&hpp[(size_t)b22 * 144 + (tid / 6) * 12 + (tid % 6)], S[tid]);
    if (tid < 6) atomicAdd(&bp[(size_t)h2 * 12 + tid], g[tid]);
    return;
  }
  const int NC = h3 >= 0 ? 36 : 24;   // columns of M in use
  for (int j = tid; j < 6 * GPBA_REC_MS; j += blockDim.x) M[j] = rec[(size_t)r * GPBA_REC_STRIDE + GPBA_REC_M + j];
  __syncthreads();
  for (int j = tid; j < 6 * GPBA_REC_MS; j += blockDim.x) {
    const int m = j / GPBA_REC_MS, c = j % GPBA_REC_MS;
    double s = 0.0;
#pragma unroll
    for (int n = 0; n < 6; ++n) s = fma(S[m * 6 + n], M[n * GPBA_REC_MS + c], s);
    T[j] = s;
  }
  __syncthreads();
  const int hh[3] = {h1, h2, h3};
  // block of slice pair (A, B), A <= B: 11, 12, 22, 13, 23, 33
  const int blk_of[3][3] = {{V.rec_hpp11[r], V.rec_hpp12[r], V.rec_hpp13[r]}, {-1, b22, V.rec_hpp23[r]}, {-1, -1, V.rec_hpp33[r]}};
  for (int j = tid; j < NC * NC; j += blockDim.x) {
    const int ap = j / NC, bpp = j % NC;
    const int A = ap / 12, B = bpp / 12;
    if (A > B || hh[A] < 0 || hh[B] < 0) continue;
    const int ra = ap % 12, cb = bpp % 12;
    if ((A == 2 && ra >= 6) || (B == 2 && cb >= 6)) continue;   // padding of the extrinsic's slot
    double s = 0.0;
#pragma unroll
    for (int m = 0; m < 6; ++m) s = fma(M[m * GPBA_REC_MS + ap], T[m * GPBA_REC_MS + bpp], s);
    const int b = blk_of[A][B];
    if (b < 0) continue;
    const int blk = b & 0x3fffffff;
    if (b & 0x40000000) atomicAdd(&hpp[(size_t)blk * 144 + cb * 12 + ra], s);  // stored transposed
    else atomicAdd(&hpp[(size_t)blk * 144 + ra * 12 + cb], s);
  }
  if (tid < NC) {
    const int A = tid / 12;
    if (hh[A] >= 0 && !(A == 2 && tid % 12 >= 6)) {
      double s = 0.0;
#pragma unroll
      for (int m = 0; m < 6; ++m) s = fma(M[m * GPBA_REC_MS + tid], g[m], s);
      atomicAdd(&bp[(size_t)hh[A] * 12 + (tid % 12)], s);
    }
  }
}

// ------------------------------------------------------------------------------------------------ K3
// One CTA (64 threads) per EdgeGaussianPrior; EdgeVelocity handled by the tail CTAs (one thread each).
// mode 0: robustified chi2 -> prior_rho[]; mode 1: accumulate Hpp / b_p.
__global__ void __launch_bounds__(64) k_priors(DevView V, const double* __restrict__ pose, const double* __restrict__ vel,
                                               int mode, double* __restrict__ prior_rho, double* __restrict__ hpp,
                                               double* __restrict__ bp) {
  __shared__ double Ji[144], Jj[144], OJi[144], OJj[144], e[12], Oe[12];
  __shared__ double s_rho1;
  const int tid = threadIdx.x;
  const int idx = blockIdx.x;
  if (idx >= V.n_prior) {  // EdgeVelocity: e = Vel[2], J = [0_6 | 0 0 1 0 0 0], Omega = QcInv(2,2)
    const int j = (idx - V.n_prior) * 64 + tid;
    if (j >= V.n_velp) return;
    const int k = V.velp_kf[j];
    const int h = V.kf_h[k];
    const double ev = vel[6 * k + 2];
    const double O = V.qc_inv[2];
    if (mode == 0) { prior_rho[V.n_prior + j] = h >= 0 ? ev * (O * ev) : 0.0; return; }
    if (h < 0) return;
    atomicAdd(&hpp[(size_t)V.pose_hpp_diag[h] * 144 + 8 * 12 + 8], O);
    atomicAdd(&bp[(size_t)h * 12 + 8], -(O * ev));
    return;
  }
  const int k1 = V.prior_kf1[idx], k2 = V.prior_kf2[idx];
  const int h1 = V.kf_h[k1], h2 = V.kf_h[k2];
  if (h1 < 0 && h2 < 0) {  // allVerticesFixed: inactive
    if (mode == 0 && tid == 0) prior_rho[idx] = 0.0;
    return;
  }
  const double dt = V.kf_time[k2] - V.kf_time[k1];
  if (tid == 0) {
    const SE3 T1 = load_se3(pose + 7 * k1), T2 = load_se3(pose + 7 * k2);
    const V6 v1 = load_v6(vel + 6 * k1), v2 = load_v6(vel + 6 * k2);
    const SE3 T = se3_mul(se3_inv(T1), T2);
    const V6 xi = se3_log(T);
    const M6 K = RightJacobianPose3Inv(xi);
    const V6 Kv2 = mul(K, v2);
    for (int i = 0; i < 6; ++i) { e[i] = xi[i] - dt * v1[i]; e[6 + i] = Kv2[i] - v1[i]; }
    if (mode == 1) {
      const M6 a = se3Adj(v2);
      const M6 A = scale(-1.0, mul(K, se3_Adj(se3_inv(T))));  // -Jr^-1 Adj(T)^-1  (G2oTypes.cc:110)
      const M6 haA = mul(scale(-0.5, a), A);
      const M6 haK = mul(scale(-0.5, a), K);
      for (int j = 0; j < 144; ++j) { Ji[j] = 0.0; Jj[j] = 0.0; }
      for (int r = 0; r < 6; ++r)
        for (int c = 0; c < 6; ++c) {
          Ji[r * 12 + c] = A(r, c);
          Ji[(6 + r) * 12 + c] = haA(r, c);
          Jj[r * 12 + c] = K(r, c);
          Jj[(6 + r) * 12 + c] = haK(r, c);
          Jj[(6 + r) * 12 + 6 + c] = K(r, c);
        }
      for (int r = 0; r < 6; ++r) { Ji[r * 12 + 6 + r] = -dt; Ji[(6 + r) * 12 + 6 + r] = -1.0; }
    }
  }
  __syncthreads();
  // Omega = QiInv(dt) = [[12/dt^3, -6/dt^2], [-6/dt^2, 4/dt]] (x) QcInv   (GaussianProcess.h:31-41)
  const double dt2 = dt * dt, dt3 = dt2 * dt;
  const double o11 = 12.0 / dt3, o12 = -6.0 / dt2, o22 = 4.0 / dt;
  if (tid < 12) {
    const int i = tid % 6;
    Oe[tid] = tid < 6 ? V.qc_inv[i] * o11 * e[i] + V.qc_inv[i] * o12 * e[6 + i]
                      : V.qc_inv[i] * o12 * e[i] + V.qc_inv[i] * o22 * e[6 + i];
  }
  __syncthreads();
  if (tid == 0) {
    double chi2 = 0.0;
    for (int i = 0; i < 12; ++i) chi2 += e[i] * Oe[i];
    double rho1 = 1.0, rho = chi2;
    if (V.hub_prior_delta > 0.0) rho = huber(chi2, V.hub_prior_delta, V.hub_prior_dsqr, &rho1);
    s_rho1 = rho1;
    if (mode == 0) prior_rho[idx] = rho;
  }
  __syncthreads();
  if (mode == 0) return;
  const double rho1 = s_rho1;
  for (int j = tid; j < 144; j += 64) {
    const int r = j / 12, c = j % 12, i = r % 6;
    const double q = V.qc_inv[i] * rho1;
    if (r < 6) { OJi[j] = q * (o11 * Ji[i * 12 + c] + o12 * Ji[(6 + i) * 12 + c]); OJj[j] = q * (o11 * Jj[i * 12 + c] + o12 * Jj[(6 + i) * 12 + c]); }
    else { OJi[j] = q * (o12 * Ji[i * 12 + c] + o22 * Ji[(6 + i) * 12 + c]); OJj[j] = q * (o12 * Jj[i * 12 + c] + o22 * Jj[(6 + i) * 12 + c]); }
  }
  __syncthreads();
  const int b11 = V.prior_hpp11[idx], b12 = V.prior_hpp12[idx], b22 = V.prior_hpp22[idx];
  for (int j = tid; j < 144; j += 64) {
    const int r = j / 12, c = j % 12;
    double sii = 0.0, sjj = 0.0, sij = 0.0;
#pragma unroll
    for (int k = 0; k < 12; ++k) {
      sii = fma(Ji[k * 12 + r], OJi[k * 12 + c], sii);
      sjj = fma(Jj[k * 12 + r], OJj[k * 12 + c], sjj);
      sij = fma(Ji[k * 12 + r], OJj[k * 12 + c], sij);
    }
    if (h1 >= 0) atomicAdd(&hpp[(size_t)b11 * 144 + j], sii);
    if (h2 >= 0) atomicAdd(&hpp[(size_t)b22 * 144 + j], sjj);
    if (b12 >= 0) {
      const int blk = b12 & 0x3fffffff;
      if (b12 & 0x40000000) atomicAdd(&hpp[(size_t)blk * 144 + c * 12 + r], sij);
      else atomicAdd(&hpp[(size_t)blk * 144 + j], sij);
    }
  }
  if (tid < 24) {
    const int r = tid % 12;
    const double* J = tid < 12 ? Ji : Jj;
    const int hh = tid < 12 ? h1 : h2;
    double s = 0.0;
    for (int k = 0; k < 12; ++k) s = fma(J[k * 12 + r], -rho1 * Oe[k], s);
    if (hh >= 0) atomicAdd(&bp[(size_t)hh * 12 + r], s);
  }
}

// ------------------------------------------------------------------------------------------------ K4
// K4a: one warp per landmark: D = Hll + lambda I = L L^T, U_o = W_o L^-T (so that W_o D^-1 W_o'^T = U_o U_o'^T),
// z = L^-1 b_l.  ptL[9*lm] = {l00,l10,l11,l20,l21,l22, z0,z1,z2}.  Pure streaming: each W row is read and each U row
// written once.
// U layout (GPBA_U_STRIDE = 18 doubles per observation, dense): three k-slices of six, U[k][m] = (U_o)_(m,k).  A DMMA operand
// fetch of K4b wants 6 rows x one k per pair: here that is one contiguous 48-byte run, the three fetches of an operand cover
// 144 contiguous bytes = 5 sectors.  History: the 6 x 3 row-major layout of round 1 spread every fetch over the whole row
// (15 sector requests per operand, K4b on the L1 tag limit: l1tex 98 %); three padded 64-byte slices (with z_l in slot 6) cut
// that to 6 sectors but moved 192 B per row through the L1 miss path, which is what bounds K4b (profiles/r02_k4b_experiments.txt).
#define GPBA_U_STRIDE 18
#define GPBA_U_K 6
__global__ void __launch_bounds__(128) k_schur_prep(DevView V, double lambda, const double* __restrict__ hll,
                                                    const double* __restrict__ bl, const double* __restrict__ W,
                                                    double* __restrict__ U, double* __restrict__ ptL, int* __restrict__ fail) {
  // eight lanes per landmark, four landmarks per warp (a landmark has ~60 rows; the factorization of its 3 x 3 block is a
  // chain of three square roots and six divisions that a whole warp per landmark only waited for)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, grp = lane >> 3, sl = lane & 7;
  for (int base = (blockIdx.x * 4 + warp) * 4; base < V.n_lm; base += gridDim.x * 16) {
    const int lm = base + grp;
    if (lm >= V.n_lm) continue;
    const double* H = hll + 9 * (size_t)lm;
    const double d00 = H[0] + lambda, d10 = H[3], d11 = H[4] + lambda, d20 = H[6], d21 = H[7], d22 = H[8] + lambda;
    const double l00 = sqrt(d00);
    const double l10 = d10 / l00, l20 = d20 / l00;
    const double l11 = sqrt(d11 - l10 * l10);
    const double l21 = (d21 - l20 * l10) / l11;
    const double l22 = sqrt(d22 - l20 * l20 - l21 * l21);
    const double z0 = bl[3 * (size_t)lm] / l00;
    const double z1 = (bl[3 * (size_t)lm + 1] - l10 * z0) / l11;
    const double z2 = (bl[3 * (size_t)lm + 2] - l20 * z0 - l21 * z1) / l22;
    if (sl == 0) {
      if (!(l00 > 0.0) || !(l11 > 0.0) || !(l22 > 0.0)) atomicExch(fail, 1);
      double* o = ptL + 9 * (size_t)lm;
      o[0] = l00; o[1] = l10; o[2] = l11; o[3] = l20; o[4] = l21; o[5] = l22; o[6] = z0; o[7] = z1; o[8] = z2;
    }
    const int64_t ob = V.lm_obs_begin[lm];
    const int nobs = (int)(V.lm_obs_begin[lm + 1] - ob);
    const double* B = W + (size_t)ob * 18;
    double* Uo = U + (size_t)ob * GPBA_U_STRIDE;
    const double i00 = 1.0 / l00, i11 = 1.0 / l11, i22 = 1.0 / l22;
    for (int rr = sl; rr < nobs * 6; rr += 8) {
      const int o = rr / 6, m = rr - 6 * o;
      const double b0 = B[rr * 3], b1 = B[rr * 3 + 1], b2 = B[rr * 3 + 2];
      const double u0 = b0 * i00;
      const double u1 = (b1 - u0 * l10) * i11;
      const double u2 = (b2 - u0 * l20 - u1 * l21) * i22;
      double* d = Uo + (size_t)o * GPBA_U_STRIDE + m;
      d[0] = u0; d[GPBA_U_K] = u1; d[2 * GPBA_U_K] = u2;
    }
  }
}

// K4b: C_(r,r') = sum U_o U_o'^T over the observation pairs (o in r, o' in r') of common landmarks: a 6 x 6 x 3L GEMM
// per record pair on the FP64 tensor pipe (mma.sync.m8n8k4.f64 -> DMMA), one warp per work item = (record pair, chunk
// of its pair list).  Diagonal record pairs hold the self pairs (o, o): there row 6 of the second operand carries z_l (from
// ptL), which yields g'_r = sum_o U_o z_l (the record's share of Hpl D^-1 b_l) in column 6.  C is stored 6 x 8 row-major.
GPBA_D void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

#define GPBA_RP_STRIDE 48
__global__ void __launch_bounds__(128) k_schur_pairs(int n_items, const int* __restrict__ item_rp,
                                                     const int64_t* __restrict__ item_begin, const int64_t* __restrict__ item_end,
                                                     const unsigned char* __restrict__ item_flags /* 1: diagonal, 2: atomic */,
                                                     const unsigned long long* __restrict__ pairs, const int* __restrict__ o_lm,
                                                     const double* __restrict__ U, const double* __restrict__ ptL,
                                                     double* __restrict__ C, int* __restrict__ next_item, int batch) {
  const int lane = threadIdx.x & 31;
  const int gid = lane >> 2, tig = lane & 3;
  __shared__ int s_base, s_taken;
  // Items are handed out in list order through a counter: the warps in flight then always work on one tight window
  // of the (landmark chunk, record pair) list, whose U rows stay L2 resident.  (With a static grid-stride assignment
  // the warps drift apart -- item lengths differ 100x -- and every L1 miss went to HBM: ncu lts hit rate 20 %.)
  // A CTA takes `batch` CONSECUTIVE items at a time and its warps share them: the list is sorted by (chunk, first record,
  // second record), so the items of a batch mostly pair the same first-side rows with different partners, and those rows
  // are then served by this SM's L1 instead of crossing the L2 -> SM fabric once per pair (round 1: 9.0 GB per launch).
  for (;;) {
    __syncthreads();   // the previous batch is consumed
    if (threadIdx.x == 0) { s_base = atomicAdd(next_item, batch); s_taken = 0; }
    __syncthreads();
    const int base = s_base;
    if (base >= n_items) break;
   for (;;) {
    int it = 0;
    if (lane == 0) it = atomicAdd(&s_taken, 1);
    it = __shfl_sync(0xffffffffu, it, 0);
    if (it >= batch || base + it >= n_items) break;
    it += base;
    const int64_t pb = item_begin[it];
    const int np = (int)(item_end[it] - pb);
    const unsigned fl = item_flags[it];
    // Four pairs per trip: K slot `tig` of the three DMMAs (one per column cc of the 6 x 3 blocks) belongs to pair
    // 4 j + tig, so a lane reads one pair index and three consecutive doubles of each block per trip (no div / mod,
    // one index load per three DMMAs); three accumulator pairs = independent DMMA chains.  The next trip's pair index
    // is fetched before this trip's gathers so the dependent loads overlap.
    double c0 = 0.0, c1 = 0.0, e0 = 0.0, e1 = 0.0, f0 = 0.0, f1 = 0.0;
    unsigned long long pr = tig < np ? pairs[pb + tig] : 0ull;
    for (int j = 0; j < np; j += 4) {
      const bool live = j + tig < np;
      const unsigned long long cur = pr;
      if (j + 4 + tig < np) pr = pairs[pb + j + 4 + tig];
      double a0 = 0.0, a1 = 0.0, a2 = 0.0, b0 = 0.0, b1 = 0.0, b2 = 0.0;
      if (live) {
        // six lanes fetch one contiguous 48-byte k-slice of a row; lane group 6 of the second operand supplies z_l on
        // diagonal record pairs (column 6 of C: g'_r), lane group 7 stays zero
        const unsigned oa = (unsigned)(cur >> 32), ob = (unsigned)cur;
        if (gid < 6) {
          const double* pa = U + (size_t)oa * GPBA_U_STRIDE + gid;
          const double* pbb = U + (size_t)ob * GPBA_U_STRIDE + gid;
          a0 = pa[0]; a1 = pa[GPBA_U_K]; a2 = pa[2 * GPBA_U_K];
          b0 = pbb[0]; b1 = pbb[GPBA_U_K]; b2 = pbb[2 * GPBA_U_K];
        } else if (gid == 6 && (fl & 1u)) {
          const double* z = ptL + 9 * (size_t)o_lm[ob] + 6;
          b0 = z[0]; b1 = z[1]; b2 = z[2];
        }
      }
      dmma884(c0, c1, a0, b0);
      dmma884(e0, e1, a1, b1);
      dmma884(f0, f1, a2, b2);
    }
    c0 += e0 + f0; c1 += e1 + f1;
    if (gid < 6) {
      double* out = C + (size_t)item_rp[it] * GPBA_RP_STRIDE + gid * 8 + 2 * tig;
      if (fl & 2u) { atomicAdd(out, c0); atomicAdd(out + 1, c1); }
      else *reinterpret_cast<double2*>(out) = make_double2(c0, c1);
    }
   }
  }
}

// K4c: one WARP per Hschur block (i, j), i <= j:
//   Hs_ij = Hpp_ij (+ lambda on the diagonal) - sum over the block's contributions of M_L^T C^ M_R
// where L / R are the records (and their keyframe slice) touching pose i / pose j and C^ = C_(r1,r2) or its transpose
// when the pair is listed the other way round.  Contributions are grouped by L on the device (build_structure):
// T = sum_R C^ M_R (6 x 12) accumulates in DMMA accumulator fragments (m8n8k4: C^ padded to 8 x 8, M_R to 8 x 16), and
// M_L^T T (12 x 12, four 8 x 8 tiles) is formed once per group with T moved from accumulator to B-operand layout by
// warp shuffles.  Per contribution a lane issues 6 loads and 4 DMMAs (the scalar version issued 12 loads per thread of a
// 72-thread half CTA and was L1-bound: ncu l1tex throughput 81 %, 0.83 ms at C4).  The block is written once, in a
// fixed order: no atomics.  On diagonal blocks the diagonal record pairs also give bschur_i = b_p,i - sum M_L^T g'_r.
struct HsContrib { int rp; int rL; int rR; int code; };  // code: bits 0-1 slice of L, bits 2-3 slice of R (0 previous keyframe, 1 current keyframe, 2 extrinsic), bit 4 transpose C, bit 5 g' entry, bits 8.. group size (first entry)
#define GPBA_K4C_WARPS 4
__global__ void __launch_bounds__(32 * GPBA_K4C_WARPS) k_schur_expand(DevView V, double lambda, const double* __restrict__ rec,
                                                                     const double* __restrict__ hpp, const double* __restrict__ bp,
                                                                     const int* __restrict__ con_begin, const HsContrib* __restrict__ con,
                                                                     const double* __restrict__ C, double* __restrict__ hs,
                                                                     double* __restrict__ bs) {
  const int blk = blockIdx.x * GPBA_K4C_WARPS + (threadIdx.x >> 5);
  if (blk >= V.n_hs) return;
  const int lane = threadIdx.x & 31, gid = lane >> 2, tig = lane & 3;
  const int src = V.hs_from_hpp[blk];
  const int diag = V.hs_diag_pose[blk];
  // accumulator tiles of the 12 x 12 block: acc[mt][nt] = rows 8 mt + gid, columns 8 nt + 2 tig + {0, 1}
  double2 acc[2][2];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) {
      const int r = 8 * mt + gid, c = 8 * nt + 2 * tig;
      double2 v = make_double2(0.0, 0.0);
      if (src >= 0 && r < 12 && c < 12) v = *reinterpret_cast<const double2*>(hpp + (size_t)src * 144 + r * 12 + c);
      if (diag >= 0 && r < 12) {
        // lambda on the diagonal; the padding dimensions of an extrinsic's 12-slot get a unit diagonal (zero rows otherwise)
        const double pad = (diag >= V.n_pose_kf) ? 1.0 : 0.0;
        if (r == c) v.x += lambda + (r >= 6 ? pad : 0.0);
        if (r == c + 1) v.y += lambda + (r >= 6 ? pad : 0.0);
      }
      acc[mt][nt] = v;
    }
  double bpart = 0.0;   // lanes < 12
  const int end = con_begin[blk + 1];
#pragma unroll 1
  for (int e = con_begin[blk]; e < end;) {
    const HsContrib g = con[e];
    const int gs = g.code >> 8;
    double2 T[2] = {make_double2(0.0, 0.0), make_double2(0.0, 0.0)};   // T[gid][8 nt + 2 tig + {0,1}], rows 6, 7 stay zero
#pragma unroll 1
    for (int q = e; q < e + gs; ++q) {
      const HsContrib cn = q == e ? g : con[q];
      const double* Cp = C + (size_t)cn.rp * GPBA_RP_STRIDE;
      const double* Mb = rec + (size_t)cn.rR * GPBA_REC_STRIDE + GPBA_REC_M + 12 * ((cn.code >> 2) & 3);
      const int sm = (cn.code & 16) ? 1 : 8, sn = (cn.code & 16) ? 8 : 1;  // C^[m][n] = C[n][m] when transposed
      double a[2], b[2][2];
#pragma unroll
      for (int kk = 0; kk < 2; ++kk) {
        const int k = 4 * kk + tig;
        a[kk] = (gid < 6 && k < 6) ? Cp[gid * sm + k * sn] : 0.0;
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) b[kk][nt] = (k < 6 && 8 * nt + gid < 12) ? Mb[k * GPBA_REC_MS + 8 * nt + gid] : 0.0;
      }
#pragma unroll
      for (int kk = 0; kk < 2; ++kk)
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) dmma884(T[nt].x, T[nt].y, a[kk], b[kk][nt]);
      if ((cn.code & 32) && lane < 12) {
        const double* Ma = rec + (size_t)cn.rL * GPBA_REC_STRIDE + GPBA_REC_M + 12 * (cn.code & 3) + lane;
#pragma unroll
        for (int mm = 0; mm < 6; ++mm) bpart = fma(Ma[mm * GPBA_REC_MS], Cp[mm * 8 + 6], bpart);
      }
    }
    // acc -= M_L^T T : A[m'][k] = -M_L[k][m'], B[k][n] = T[k][n] (accumulator layout -> B-operand layout by shuffles)
    const double* Ma = rec + (size_t)g.rL * GPBA_REC_STRIDE + GPBA_REC_M + 12 * (g.code & 3);
#pragma unroll
    for (int kk = 0; kk < 2; ++kk) {
      const int k = 4 * kk + tig;
      const int srcl = (k << 2) | (gid >> 1);   // lane holding T[k][8 nt + gid] in its T[nt].x / .y (gid even / odd)
      double bt[2];
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        const double x = __shfl_sync(0xffffffffu, T[nt].x, srcl), y = __shfl_sync(0xffffffffu, T[nt].y, srcl);
        bt[nt] = (gid & 1) ? y : x;
      }
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        const double am = (k < 6 && 8 * mt + gid < 12) ? -Ma[k * GPBA_REC_MS + 8 * mt + gid] : 0.0;
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) dmma884(acc[mt][nt].x, acc[mt][nt].y, am, bt[nt]);
      }
    }
    e += gs;
  }
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) {
      const int r = 8 * mt + gid, c = 8 * nt + 2 * tig;
      if (r < 12 && c < 12) *reinterpret_cast<double2*>(hs + (size_t)blk * 144 + r * 12 + c) = acc[mt][nt];
    }
  if (diag >= 0 && lane < 12) bs[(size_t)diag * 12 + lane] = bp[(size_t)diag * 12 + lane] - bpart;
}

// ------------------------------------------------------------------------------------------------ K6
// y_r = M_r [x_kf1; x_kf2]: the pose update seen from record r (6-vector), one thread per (record, row).
__global__ void k_rec_y(DevView V, const double* __restrict__ rec, const double* __restrict__ xp, double* __restrict__ Y) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= V.n_rec * 6) return;
  const int r = t / 6, m = t % 6;
  const int k1 = V.rec_kf1[r], k2 = V.rec_kf2[r];
  const int h1 = k1 >= 0 ? V.kf_h[k1] : -1, h2 = V.kf_h[k2];
  const int h3 = k1 >= 0 ? V.ext_h[V.rec_cam[r]] : -1;
  const double* M = rec + (size_t)r * GPBA_REC_STRIDE + GPBA_REC_M + m * GPBA_REC_MS;
  double s = 0.0;
  if (h1 >= 0) {
#pragma unroll
    for (int c = 0; c < 12; ++c) s = fma(M[c], xp[(size_t)h1 * 12 + c], s);
  }
  if (h2 >= 0) {
#pragma unroll
    for (int c = 0; c < 12; ++c) s = fma(M[12 + c], xp[(size_t)h2 * 12 + c], s);
  }
  if (h3 >= 0) {
#pragma unroll
    for (int c = 0; c < 6; ++c) s = fma(M[24 + c], xp[(size_t)h3 * 12 + c], s);
  }
  Y[t] = s;
}

// Landmarks: x_l = D^-1 (b_l - Hpl^T x_p) = L^-T (z - sum_o U_o^T y_r(o)); pt_new = pt + x_l.
// partial[] receives sum x_l (lambda x_l + b_l) for computeScale (optimization_algorithm_levenberg.cpp:187-194).
// Eight lanes per landmark, four landmarks per warp: a landmark has ~60 rows (10 observations x 6), so a whole warp per
// landmark spent most of its time in three 5-step reductions and a one-lane triangular solve with three divisions (ncu:
// latency-bound, 34 % of HBM); here the reductions are 3 steps and four solves run side by side.
__global__ void __launch_bounds__(128) k_backsub(DevView V, double lambda, const double* __restrict__ U,
                                                 const double* __restrict__ ptL, const double* __restrict__ bl,
                                                 const double* __restrict__ Y, const double* __restrict__ pt_cur,
                                                 double* __restrict__ pt_new, double* __restrict__ xl,
                                                 double* __restrict__ partial) {
  __shared__ double red[32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, grp = lane >> 3, sl = lane & 7;
  double sc = 0.0;
  for (int base = (blockIdx.x * 4 + warp) * 4; base < V.n_lm; base += gridDim.x * 16) {
    const int lm = base + grp;
    const bool live = lm < V.n_lm;
    const int64_t ob = live ? V.lm_obs_begin[lm] : 0;
    const int nrow = live ? (int)(V.lm_obs_begin[lm + 1] - ob) * 6 : 0;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0;
    for (int rr = sl; rr < nrow; rr += 8) {
      const int o = rr / 6, m = rr - 6 * o;
      const double y = Y[(size_t)V.o_rec[ob + o] * 6 + m];
      const double* u = U + (size_t)(ob + o) * GPBA_U_STRIDE + m;
      a0 = fma(u[0], y, a0); a1 = fma(u[GPBA_U_K], y, a1); a2 = fma(u[2 * GPBA_U_K], y, a2);
    }
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {   // sum over the eight lanes of the group
      a0 += __shfl_xor_sync(0xffffffffu, a0, o); a1 += __shfl_xor_sync(0xffffffffu, a1, o); a2 += __shfl_xor_sync(0xffffffffu, a2, o);
    }
    if (live && sl == 0) {
      const double* L = ptL + 9 * (size_t)lm;
      const double y0 = L[6] - a0, y1 = L[7] - a1, y2 = L[8] - a2;
      const double x2 = y2 / L[5];
      const double x1 = (y1 - L[4] * x2) / L[2];
      const double x0 = (y0 - L[1] * x1 - L[3] * x2) / L[0];
      pt_new[3 * (size_t)lm] = pt_cur[3 * (size_t)lm] + x0;          // VertexSBAPointXYZ::oplusImpl (types_sba.h:41-57)
      pt_new[3 * (size_t)lm + 1] = pt_cur[3 * (size_t)lm + 1] + x1;
      pt_new[3 * (size_t)lm + 2] = pt_cur[3 * (size_t)lm + 2] + x2;
      xl[3 * (size_t)lm] = x0; xl[3 * (size_t)lm + 1] = x1; xl[3 * (size_t)lm + 2] = x2;
      sc += x0 * (lambda * x0 + bl[3 * (size_t)lm]) + x1 * (lambda * x1 + bl[3 * (size_t)lm + 1]) +
            x2 * (lambda * x2 + bl[3 * (size_t)lm + 2]);
    }
  }
  const double s = block_sum(sc, red);
  if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

// Keyframes: Twb <- Twb * exp(x[0:6]), Vel += x[6:12]   (PoseVelocity::Update, src/G2oTypes.cc:41-46).
// pose_scale[h] = sum over the 12 dims of x (lambda x + b).
__global__ void k_update_poses(DevView V, double lambda, const double* __restrict__ xp, const double* __restrict__ bp,
                               const double* __restrict__ pose_cur, const double* __restrict__ vel_cur,
                               double* __restrict__ pose_new, double* __restrict__ vel_new, double* __restrict__ pose_scale) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= V.n_kf) return;
  const int h = V.kf_h[k];
  if (h < 0) return;  // fixed / inactive: both buffers already hold the same value
  const double* x = xp + (size_t)h * 12;
  const SE3 T = se3_mul(load_se3(pose_cur + 7 * k), se3_exp(load_v6(x)));
  store_se3(T, pose_new + 7 * k);
  double sc = 0.0;
  for (int i = 0; i < 6; ++i) vel_new[6 * k + i] = vel_cur[6 * k + i] + x[6 + i];
  for (int i = 0; i < 12; ++i) sc += x[i] * (lambda * x[i] + bp[(size_t)h * 12 + i]);
  pose_scale[h] = sc;
}

// Extrinsics: Tbc <- Tbc * exp(x[0:6])  (VertexExtrinsic::oplusImpl, include/G2oTypes.h:98-100); fixed ones are copied.
// One thread per camera; also derives the per-camera constants of the new state.
GPBA_D void cam_from_tbc(const SE3& Tbc, CamConst& cc) {
  const SE3 Tcb = se3_inv(Tbc);
  const M3 Rcb = quat_to_R(Tcb.q), Rbc = quat_to_R(Tbc.q);
#pragma unroll
  for (int i = 0; i < 9; ++i) { cc.Rcb[i] = Rcb.a[i]; cc.Rbc[i] = Rbc.a[i]; }
#pragma unroll
  for (int i = 0; i < 3; ++i) { cc.tcb[i] = Tcb.t[i]; cc.tbc[i] = Tbc.t[i]; }
  cc.qbc[0] = Tbc.q.x; cc.qbc[1] = Tbc.q.y; cc.qbc[2] = Tbc.q.z; cc.qbc[3] = Tbc.q.w;
  const M6 A = se3_Adj(Tbc);
#pragma unroll
  for (int r = 0; r < 6; ++r)
#pragma unroll
    for (int c = 0; c < 6; ++c) cc.AdjTbc[r * 6 + c] = A(r, c);
}
__global__ void k_update_ext(DevView V, double lambda, const double* __restrict__ xp, const double* __restrict__ bp,
                             const double* __restrict__ ext_cur, double* __restrict__ ext_new, const CamConst* __restrict__ cam_cur,
                             CamConst* __restrict__ cam_new, double* __restrict__ pose_scale) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= V.n_cam) return;
  const int h = V.ext_h[c];
  SE3 T = load_se3(ext_cur + 7 * c);
  if (h >= 0) {
    const double* x = xp + (size_t)h * 12;
    T = se3_mul(T, se3_exp(load_v6(x)));
    double sc = 0.0;
    for (int i = 0; i < 6; ++i) sc += x[i] * (lambda * x[i] + bp[(size_t)h * 12 + i]);
    pose_scale[h] = sc;   // the padding dimensions carry x = 0
  }
  store_se3(T, ext_new + 7 * c);
  CamConst cc = cam_cur[c];
  if (h >= 0) cam_from_tbc(T, cc);
  cam_new[c] = cc;
}

// EdgeExtrinsicPrior (include/G2oTypes.h:470-494): e = Log(R_ini^-1 R_bc), J = [0 | Jr(e)^-1] (RightJacobianSO3 of
// src/G2oTypes.cc:575-590, inverted like Eigen's fixed 3 x 3 inverse), information = MultiFrame::mRbc_ini_cov[c], no kernel.
// One thread per camera.  mode 0: chi2 -> rho[c]; mode 1: Hpp / b of the extrinsic's rotation block.
__global__ void k_ext_prior(DevView V, const double* __restrict__ ext, int mode, double* __restrict__ rho,
                            double* __restrict__ hpp, double* __restrict__ bp) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= V.n_cam) return;
  const int h = V.ext_h[c];
  if (h < 0 || !V.ext_prior_on[c]) { if (mode == 0) rho[c] = 0.0; return; }
  const SE3 T = load_se3(ext + 7 * c);
  Quat qi; qi.x = V.ext_prior_qinv[4 * c]; qi.y = V.ext_prior_qinv[4 * c + 1]; qi.z = V.ext_prior_qinv[4 * c + 2]; qi.w = V.ext_prior_qinv[4 * c + 3];
  double theta_;
  const V3 e = so3_log(quat_mul(qi, T.q), &theta_);
  const double* O = V.ext_prior_info + 9 * c;
  double Oe[3];
  for (int r = 0; r < 3; ++r) Oe[r] = O[3 * r] * e[0] + O[3 * r + 1] * e[1] + O[3 * r + 2] * e[2];
  if (mode == 0) { rho[c] = e[0] * Oe[0] + e[1] * Oe[1] + e[2] * Oe[2]; return; }
  // Jr(e) = I - W (1 - cos d) / d^2 + W^2 (d - sin d) / d^3, identity below d = 1e-5
  const double d2 = e[0] * e[0] + e[1] * e[1] + e[2] * e[2], d = sqrt(d2);
  double Jr[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  if (!(d < 1e-5)) {
    const double W[9] = {0, -e[2], e[1], e[2], 0, -e[0], -e[1], e[0], 0};
    const double a = (1.0 - cos(d)) / d2, b2 = (d - sin(d)) / (d2 * d);
    for (int r = 0; r < 3; ++r)
      for (int k = 0; k < 3; ++k) {
        double ww = 0.0;
        for (int q = 0; q < 3; ++q) ww += W[3 * r + q] * W[3 * q + k];
        Jr[3 * r + k] += -a * W[3 * r + k] + b2 * ww;
      }
  }
  double Cf[9];   // cofactor inverse
  Cf[0] = Jr[4] * Jr[8] - Jr[5] * Jr[7]; Cf[1] = Jr[2] * Jr[7] - Jr[1] * Jr[8]; Cf[2] = Jr[1] * Jr[5] - Jr[2] * Jr[4];
  Cf[3] = Jr[5] * Jr[6] - Jr[3] * Jr[8]; Cf[4] = Jr[0] * Jr[8] - Jr[2] * Jr[6]; Cf[5] = Jr[2] * Jr[3] - Jr[0] * Jr[5];
  Cf[6] = Jr[3] * Jr[7] - Jr[4] * Jr[6]; Cf[7] = Jr[1] * Jr[6] - Jr[0] * Jr[7]; Cf[8] = Jr[0] * Jr[4] - Jr[1] * Jr[3];
  const double idet = 1.0 / (Jr[0] * Cf[0] + Jr[1] * Cf[3] + Jr[2] * Cf[6]);
  double J[9], JtO[9];
  for (int i = 0; i < 9; ++i) J[i] = Cf[i] * idet;
  for (int a2 = 0; a2 < 3; ++a2)
    for (int k = 0; k < 3; ++k) JtO[3 * a2 + k] = J[a2] * O[k] + J[3 + a2] * O[3 + k] + J[6 + a2] * O[6 + k];
  double* B = hpp + (size_t)V.pose_hpp_diag[h] * 144;
  for (int a2 = 0; a2 < 3; ++a2) {
    for (int k = 0; k < 3; ++k) atomicAdd(&B[(3 + a2) * 12 + 3 + k], JtO[3 * a2] * J[k] + JtO[3 * a2 + 1] * J[3 + k] + JtO[3 * a2 + 2] * J[6 + k]);
    atomicAdd(&bp[(size_t)h * 12 + 3 + a2], -(JtO[3 * a2] * e[0] + JtO[3 * a2 + 1] * e[1] + JtO[3 * a2 + 2] * e[2]));
  }
}

// ------------------------------------------------------------------------------------------------ K8
// flag = chi2 > threshold (close / far) || !isDepthPositive at BOTH keyframe poses (G2oTypes.h:362-370);
// stereo edges: chi2 only (Optimizer.cc:1283-1296).  Runs over ALL observations in original order.
__global__ void k_flags(DevView V, int64_t n_obs, const double* __restrict__ chi2, const double* __restrict__ ur,
                        const int* __restrict__ obs_rec, const uint8_t* __restrict__ obs_flags,
                        const int* __restrict__ obs_pt, const double* __restrict__ pt_all, const double* __restrict__ pose,
                        double th_mono, double th_close, double th_stereo, uint8_t* __restrict__ flags) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_obs; i += (int64_t)gridDim.x * blockDim.x) {
    const double c2 = chi2[i];
    bool out;
    if (ur && ur[i] >= 0.0) {
      out = c2 > th_stereo;
    } else {
      const int r = obs_rec[i];
      const CamConst& cam = V.cam[V.rec_cam[r]];
      SE3 Tbc;
      Tbc.q.x = cam.qbc[0]; Tbc.q.y = cam.qbc[1]; Tbc.q.z = cam.qbc[2]; Tbc.q.w = cam.qbc[3];
      Tbc.t = v3(cam.tbc[0], cam.tbc[1], cam.tbc[2]);
      const V3 X = v3(pt_all[3 * (size_t)obs_pt[i]], pt_all[3 * (size_t)obs_pt[i] + 1], pt_all[3 * (size_t)obs_pt[i] + 2]);
      bool pos = se3_act(se3_inv(se3_mul(load_se3(pose + 7 * V.rec_kf2[r]), Tbc)), X)[2] > 0;
      if (V.rec_kf1[r] >= 0) pos = (se3_act(se3_inv(se3_mul(load_se3(pose + 7 * V.rec_kf1[r]), Tbc)), X)[2] > 0) && pos;
      const bool close = obs_flags[i] & 0x1u;
      out = (c2 > th_mono && !close) || (c2 > th_close && close) || !pos;
    }
    flags[i] = out ? 1 : 0;
  }
}

}  // namespace gpba
