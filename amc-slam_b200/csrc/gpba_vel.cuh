// gpba_vel.cuh -- velocity RANSAC: Tracking::MCRansac (src/Tracking.cc:1939-2002) = maxIt calls of
// Optimizer::OptimizeVel (src/Optimizer.cc:2364-2447), all hypotheses in one launch, one CTA per hypothesis.
//
// A hypothesis is a 6-dimensional LM problem over `set_size` (3) EdgeVelReproj edges (include/G2oTypes.h:521-547,
// src/G2oTypes.cc:497-510) followed by an inlier test of every matched feature.  Thread k < set_size owns sampled edge k
// (error, robust weight and the 2 x 6 Jacobian: exp, right Jacobian of SE(3)), thread 0 plays BlockSolver + the LM
// controller (6 x 6 Cholesky in registers), then the whole CTA sweeps the matches with one (R | t) per camera.  As in the
// pose-only kernel a trial is evaluated together with its Jacobians, so an accepted trial is the next linearisation.
#pragma once
#include "gpba_kernels.cuh"

namespace gpba {

#define GPBA_VEL_MAX_SET 8
#define GPBA_VEL_MAX_CAM 8
#define GPBA_VEL_THREADS 128

struct VelView {
  int n_cam, n_match, n_hyp, set_size, iterations;
  const CamConst* cam;
  const double* cam_dt;
  double Tinv[7];            // (pF2->GetPoseW())^-1
  double vel_init[6];
  const double *obs_u, *obs_v, *obs_w, *obs_xw;
  const int* obs_cam;
  const int* samples;
  double hub_delta, hub_dsqr, threshold;
  double* vel_out;
  int* inliers_out;
  uint8_t* mask_out;
  gpba_lm_trace* traces;
};

// EdgeVelReproj::computeError (+ linearizeOplus if JAC) of match i at velocity v
template <bool JAC>
GPBA_D void vel_edge(const VelView& B, const SE3& Tinv, int i, const double* __restrict__ v, double* __restrict__ e, double (*J)[6]) {
  const int c = B.obs_cam[i];
  const CamConst& cam = B.cam[c];
  const double dt = B.cam_dt[c];
  V6 ndxi;
#pragma unroll
  for (int k = 0; k < 6; ++k) ndxi[k] = -(v[k] * dt);
  SE3 Tcb;
  Tcb.q = quat_inv(Quat{cam.qbc[0], cam.qbc[1], cam.qbc[2], cam.qbc[3]});
  Tcb.t = v3(cam.tcb[0], cam.tcb[1], cam.tcb[2]);
  const SE3 Tcb1 = se3_mul(Tcb, se3_exp(ndxi));   // Tbc^-1 exp(-v dt)
  const V3 Xb = se3_act(Tinv, v3(B.obs_xw[3 * i], B.obs_xw[3 * i + 1], B.obs_xw[3 * i + 2]));
  const V3 Xc = se3_act(Tcb1, Xb);
  e[0] = B.obs_u[i] - (cam.fx * Xc[0] / Xc[2] + cam.cx);
  e[1] = B.obs_v[i] - (cam.fy * Xc[1] / Xc[2] + cam.cy);
  if (JAC) {
    double P[2][3];
    P[0][0] = cam.fx / Xc[2]; P[0][1] = 0.0; P[0][2] = -cam.fx * Xc[0] / (Xc[2] * Xc[2]);
    P[1][0] = 0.0; P[1][1] = cam.fy / Xc[2]; P[1][2] = -cam.fy * Xc[1] / (Xc[2] * Xc[2]);
    const M3 R = quat_to_R(Tcb1.q);
    const M3 Xh = hat(Xb);
    double D[3][6];   // R [I, -Xb^]  (rows 0..2 of Tcb1.matrix() * CircleDot(Xb), Pose3utils.cc:75-80)
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int cc = 0; cc < 3; ++cc) {
        D[r][cc] = R(r, cc);
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < 3; ++k) s += R(r, k) * Xh(k, cc);
        D[r][3 + cc] = -s;
      }
    const M6 Jr = RightJacobianPose3(ndxi);
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int cc = 0; cc < 6; ++cc) {
        double s = 0.0;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          double d = 0.0;
#pragma unroll
          for (int k = 0; k < 6; ++k) d += D[a][k] * Jr(k, cc);
          s += P[r][a] * d;
        }
        J[r][cc] = s * dt;
      }
  }
}

__global__ void __launch_bounds__(GPBA_VEL_THREADS) k_vel_ransac(VelView B) {
  __shared__ double sV[2][6];
  __shared__ double sE[GPBA_VEL_MAX_SET][2], sJ[GPBA_VEL_MAX_SET][2][6], sW[GPBA_VEL_MAX_SET], sRho[GPBA_VEL_MAX_SET];
  __shared__ double sKeepE[GPBA_VEL_MAX_SET][2], sKeepJ[GPBA_VEL_MAX_SET][2][6], sKeepW[GPBA_VEL_MAX_SET];
  __shared__ double sx[6], sCam[GPBA_VEL_MAX_CAM][12], sRed[32];
  __shared__ int sFlag[4];   // [1] continue trial loop, [2] outer result, [3] accepted
  const int h = blockIdx.x, tid = threadIdx.x;
  const int ns = B.set_size;
  const int* set = B.samples + (size_t)h * ns;
  const SE3 Tinv = load_se3(B.Tinv);
  if (tid < 6) sV[0][tid] = B.vel_init[tid];
  __syncthreads();
  int cur = 0;
  gpba_lm_trace* tr = B.traces ? B.traces + h : nullptr;

  // sampled edges at state buffer s: e, J, rho' w -> shared; returns activeRobustChi2 (valid on thread 0)
  auto evaluate = [&](int s) -> double {
    if (tid < ns) {
      const int i = set[tid];
      double e[2], J[2][6];
      vel_edge<true>(B, Tinv, i, sV[s], e, J);
      const double w = B.obs_w[i];
      double rho1;
      const double rho = huber(e[0] * w * e[0] + e[1] * w * e[1], B.hub_delta, B.hub_dsqr, &rho1);
      sE[tid][0] = e[0]; sE[tid][1] = e[1]; sW[tid] = rho1 * w; sRho[tid] = rho;
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int c = 0; c < 6; ++c) sJ[tid][r][c] = J[r][c];
    }
    __syncthreads();
    double sum = 0.0;
    if (tid == 0) for (int k = 0; k < ns; ++k) sum += sRho[k];
    return sum;
  };
  // the evaluation just made becomes the linearisation of the current estimate
  auto keep = [&]() {
    for (int j = tid; j < ns * 12; j += GPBA_VEL_THREADS) (&sKeepJ[0][0][0])[j] = (&sJ[0][0][0])[j];
    if (tid < ns) { sKeepE[tid][0] = sE[tid][0]; sKeepE[tid][1] = sE[tid][1]; sKeepW[tid] = sW[tid]; }
    __syncthreads();
  };

  double currentChi = evaluate(cur);
  keep();
  double lambda = 0.0, ni = 2.0;
  int nBad = 0, result = GPBA_RESULT_OK, cj = 0;
  if (tr && tid == 0) { tr->total_trials = 0; tr->cg_iterations = 0; tr->last_trial_chi2 = 0.0; }
#pragma unroll 1
  for (int it = 0; it < B.iterations; ++it) {
    double H[36], b[6];   // thread 0 only
    const double iniChi = currentChi;
    double tempChi = currentChi, rho = 0.0;
    int qmax = 0;
    if (tid == 0) {   // BaseUnaryEdge::constructQuadraticForm (base_unary_edge.hpp:43-72)
#pragma unroll
      for (int j = 0; j < 36; ++j) H[j] = 0.0;
#pragma unroll
      for (int j = 0; j < 6; ++j) b[j] = 0.0;
      for (int k = 0; k < ns; ++k)
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
          for (int a = 0; a < 6; ++a) {
            const double jo = sKeepJ[k][r][a] * sKeepW[k];
            b[a] -= jo * sKeepE[k][r];
#pragma unroll
            for (int c = 0; c < 6; ++c) H[a * 6 + c] += jo * sKeepJ[k][r][c];
          }
      if (it == 0) {
        double mx = 0.0;
#pragma unroll
        for (int j = 0; j < 6; ++j) mx = fmax(mx, fabs(H[j * 6 + j]));
        lambda = 1e-5 * mx; ni = 2.0; nBad = 0;
      }
    }
#pragma unroll 1
    for (;;) {
      bool ok = true;
      if (tid == 0) {   // (H + lambda I) x = b: Cholesky (LinearSolverDense: !isPositive() => solve fails)
        double L[36];
#pragma unroll
        for (int j = 0; j < 36; ++j) L[j] = H[j];
#pragma unroll
        for (int j = 0; j < 6; ++j) L[j * 6 + j] += lambda;
#pragma unroll
        for (int j = 0; j < 6; ++j) {
          double d = L[j * 6 + j];
#pragma unroll
          for (int k = 0; k < j; ++k) d -= L[j * 6 + k] * L[j * 6 + k];
          if (!(d > 0.0)) ok = false;
          const double l = sqrt(d);
          L[j * 6 + j] = l;
#pragma unroll
          for (int i = j + 1; i < 6; ++i) {
            double s = L[i * 6 + j];
#pragma unroll
            for (int k = 0; k < j; ++k) s -= L[i * 6 + k] * L[j * 6 + k];
            L[i * 6 + j] = s / l;
          }
        }
        double y[6], x[6];
#pragma unroll
        for (int i = 0; i < 6; ++i) { double s = b[i]; for (int k = 0; k < i; ++k) s -= L[i * 6 + k] * y[k]; y[i] = s / L[i * 6 + i]; }
#pragma unroll
        for (int i = 5; i >= 0; --i) { double s = y[i]; for (int k = i + 1; k < 6; ++k) s -= L[k * 6 + i] * x[k]; x[i] = s / L[i * 6 + i]; }
#pragma unroll
        for (int i = 0; i < 6; ++i) { sx[i] = x[i]; sV[1 - cur][i] = sV[cur][i] + x[i]; }   // VertexVel::oplusImpl
      }
      __syncthreads();
      const double trial = evaluate(1 - cur);
      if (tid == 0) {
        tempChi = ok ? trial : 1.7976931348623157e308;
        double sc = 0.0;
#pragma unroll
        for (int j = 0; j < 6; ++j) sc += sx[j] * (lambda * sx[j] + b[j]);
        rho = (currentChi - tempChi) / (sc + 1e-3);
        int accept = 0;
        if (rho > 0 && isfinite(tempChi)) {
          double alpha = 1. - (2 * rho - 1) * (2 * rho - 1) * (2 * rho - 1);
          alpha = fmin(alpha, 2. / 3.);
          lambda *= fmax(1. / 3., alpha);
          ni = 2.0; currentChi = tempChi; accept = 1;
        } else {
          lambda *= ni; ni *= 2.0;
        }
        ++qmax;
        sFlag[1] = (rho < 0 && qmax < 10) ? 1 : 0;
        sFlag[3] = accept;
      }
      __syncthreads();
      if (sFlag[3]) { cur = 1 - cur; keep(); }
      if (!sFlag[1]) break;
    }
    if (tid == 0) {
      if (tr && it < GPBA_MAX_ITERS) {
        tr->levenberg_iterations[it] = qmax; tr->chi2_before[it] = iniChi; tr->chi2_after[it] = currentChi;
        tr->lambda[it] = lambda; tr->total_trials += qmax; tr->last_trial_chi2 = tempChi;
      }
      ++cj;
      result = GPBA_RESULT_OK;
      if (qmax == 10 || rho == 0) result = GPBA_TERMINATE;
      else { if ((iniChi - currentChi) * 1e3 < iniChi) nBad++; else nBad = 0; if (nBad >= 3) result = GPBA_TERMINATE; }
      sFlag[2] = result;
    }
    __syncthreads();
    if (sFlag[2] != GPBA_RESULT_OK) break;
  }
  if (tr && tid == 0) { tr->n_iters = cj; tr->result = result; }

  // ---- inlier test of every match at the estimate (Optimizer.cc:2425-2440): X_c = (Tbc^-1 exp(-v dt_c) T^-1) X_w
  if (tid < B.n_cam) {
    const CamConst& cam = B.cam[tid];
    V6 ndxi;
    for (int k = 0; k < 6; ++k) ndxi[k] = -(sV[cur][k] * B.cam_dt[tid]);
    SE3 Tcb;
    Tcb.q = quat_inv(Quat{cam.qbc[0], cam.qbc[1], cam.qbc[2], cam.qbc[3]});
    Tcb.t = v3(cam.tcb[0], cam.tcb[1], cam.tcb[2]);
    const SE3 Tcw = se3_mul(se3_mul(Tcb, se3_exp(ndxi)), Tinv);
    const M3 R = quat_to_R(Tcw.q);
    for (int k = 0; k < 9; ++k) sCam[tid][k] = R.a[k];
    sCam[tid][9] = Tcw.t[0]; sCam[tid][10] = Tcw.t[1]; sCam[tid][11] = Tcw.t[2];
  }
  __syncthreads();
  int cnt = 0;
  const double thr2 = B.threshold;
  for (int i = tid; i < B.n_match; i += GPBA_VEL_THREADS) {
    const int c = B.obs_cam[i];
    const double* R = sCam[c];
    const CamConst& cam = B.cam[c];
    const double X0 = B.obs_xw[3 * i], X1 = B.obs_xw[3 * i + 1], X2 = B.obs_xw[3 * i + 2];
    const double xc = fma(R[0], X0, fma(R[1], X1, fma(R[2], X2, R[9])));
    const double yc = fma(R[3], X0, fma(R[4], X1, fma(R[5], X2, R[10])));
    const double zc = fma(R[6], X0, fma(R[7], X1, fma(R[8], X2, R[11])));
    const double e0 = B.obs_u[i] - (cam.fx * xc / zc + cam.cx), e1 = B.obs_v[i] - (cam.fy * yc / zc + cam.cy);
    const bool in = sqrt(e0 * e0 + e1 * e1) <= thr2;
    if (B.mask_out) B.mask_out[(size_t)h * B.n_match + i] = in ? 1 : 0;
    cnt += in ? 1 : 0;
  }
  const double total = block_sum((double)cnt, sRed);
  if (tid == 0) {
    if (B.inliers_out) B.inliers_out[h] = (int)total;
    if (B.vel_out) for (int k = 0; k < 6; ++k) B.vel_out[6 * h + k] = sV[cur][k];
  }
}

}  // namespace gpba
