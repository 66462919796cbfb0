// gpba_pose.cuh -- pose-only GP optimisation: Optimizer::PoseGPOptimizationFromeLastFrame
// (src/Optimizer.cc:369-686) for a batch of independent frames, one persistent CTA per frame.
//
// The g2o graph of one frame has two VertexPoseVel (previous frame, fixed or not; current frame) and no landmark
// vertex: EdgeMonoGPOnlyPose / EdgeMonoOnlyPose / EdgeStereoOnlyPose carry the world point as a constant
// (src/G2oTypes.cc:120-223; pose Jacobians identical to the BA edges), so the system is the 12 or 24 dimensional
// Hpp and there is no Schur complement.  The whole schedule -- 4 x (initializeOptimization(0) + optimize(10)) with the
// float-typed chi2 tests in between -- runs inside one kernel: a launch-per-stage design would spend ~40 launches on
// every one of the ~16 LM iterations of a frame that carries a few thousand matches.
//
// Per evaluation of a state the CTA computes one record row per camera (interpolated camera pose + 6x24 chain matrix,
// K0) while another warp evaluates the Gaussian-process prior, then sweeps the matches once per camera (K1 / K2b of the
// BA path: robust chi2 and, in the same pass, the per-record S_r = sum w J1^T J1, g_r).  An LM trial is evaluated
// speculatively WITH its linearisation: when the trial is accepted (the common case) the next outer iteration starts
// from data that is already there -- one sweep per LM iteration instead of three.  H = sum M_r^T S_r M_r + prior +
// velocity edges; (H + lambda I) x = b by Cholesky in shared memory.  The LM controller follows
// optimization_algorithm_levenberg.cpp:61-194 (default lambda = tau * max diag, :171-185).
#pragma once
#include "gpba_kernels.cuh"

namespace gpba {

#define GPBA_POSE_MAX_CAM 8
#define GPBA_POSE_THREADS 256

struct PoseBatchView {
  int n_cam, n_frames;
  const CamConst* cam;
  const double *prev_pose, *prev_vel, *prev_time, *cur_pose, *cur_vel, *cur_time, *cam_time;
  const uint8_t* prev_fixed;
  const int64_t* obs_begin;
  const double *obs_u, *obs_v, *obs_ur, *obs_w, *obs_xw;
  const int* obs_cam;
  const uint8_t* obs_flags;
  uint8_t* level;        // [n_obs] 0 active / 1 outlier (in: mvbOutlier, out: final flags)
  uint8_t* kernel_off;   // [n_obs] scratch
  double* chi2;          // [n_obs] stored edge chi2 (the edge's _error survives pop(): stale-error quirk)
  double *out_cur_pose, *out_cur_vel, *out_prev_pose, *out_prev_vel;
  int* out_inliers;
  gpba_lm_trace* traces; // [n_frames][4] or null
};

// (H + lambda I) x = b for n <= 24 in shared memory: Cholesky by one warp.  Returns false on a non-positive pivot
// (LinearSolverDense: !_cholesky.isPositive(), linear_solver_dense.h:108-112).
GPBA_D bool pose_solve(int n, const double* __restrict__ H, const double* __restrict__ b, double lambda, double* __restrict__ L,
                       double* __restrict__ x, int lane) {
  for (int j = lane; j < n * n; j += 32) L[j] = H[j] + ((j / n == j % n) ? lambda : 0.0);
  __syncwarp();
  bool ok = true;
  for (int j = 0; j < n; ++j) {
    const double d = L[j * n + j];
    if (!(d > 0.0)) ok = false;
    const double l = sqrt(d);
    __syncwarp();
    if (lane == 0) L[j * n + j] = l;
    for (int i = j + 1 + lane; i < n; i += 32) L[i * n + j] /= l;
    __syncwarp();
    for (int q = lane; q < (n - j - 1) * (n - j - 1); q += 32) {
      const int r = j + 1 + q / (n - j - 1), c = j + 1 + q % (n - j - 1);
      if (c <= r) L[r * n + c] = fma(-L[r * n + j], L[c * n + j], L[r * n + c]);
    }
    __syncwarp();
  }
  if (lane == 0) {   // forward / backward substitution: 24 unknowns
    double y[24];
    for (int i = 0; i < n; ++i) { double s = b[i]; for (int k = 0; k < i; ++k) s -= L[i * n + k] * y[k]; y[i] = s / L[i * n + i]; }
    for (int i = n - 1; i >= 0; --i) { double s = y[i]; for (int k = i + 1; k < n; ++k) s -= L[k * n + i] * x[k]; x[i] = s / L[i * n + i]; }
  }
  __syncwarp();
  return ok;
}

__global__ void __launch_bounds__(GPBA_POSE_THREADS) k_pose_only(PoseBatchView B, DevView V) {
  __shared__ __align__(16) double sRec[GPBA_POSE_MAX_CAM][GPBA_REC_STRIDE];
  __shared__ double sState[2][26];          // [buffer][prev pose 7 | prev vel 6 | cur pose 7 | cur vel 6]
  __shared__ double sS[GPBA_POSE_MAX_CAM][28];
  __shared__ double sPart[GPBA_POSE_THREADS / 32][28];
  __shared__ double sH[576], sL[576], sb[24], sx[24];
  __shared__ double sJi[144], sJj[144], sOJi[144], sOJj[144], sE[12], sOe[12];
  __shared__ double sRed[32];
  __shared__ double sScal[8];               // [0] chi2 of the evaluation, [2] prior + velocity chi2, [3] lambda, [4] bad count
  __shared__ int sFlag[4];                  // [0] solve ok, [1] continue trial loop, [2] outer result, [3] accepted
  const int f = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t ob = B.obs_begin[f], oe = B.obs_begin[f + 1];
  const bool fix1 = B.prev_fixed[f] != 0;
  const int n = fix1 ? 12 : 24, o1 = fix1 ? -1 : 0, o2 = fix1 ? 0 : 12;
  const int n_cam = B.n_cam;
  const double t1 = B.prev_time[f], t2 = B.cur_time[f];
  const double dt = t2 - t1, dt2 = dt * dt, dt3 = dt2 * dt;
  const double o11 = 12.0 / dt3, o12 = -6.0 / dt2, o22 = 4.0 / dt;   // QiInv(dt) (GaussianProcess.h:31-41)
  if (tid < 7) { sState[0][tid] = B.prev_pose[7 * f + tid]; sState[0][13 + tid] = B.cur_pose[7 * f + tid]; }
  if (tid < 6) { sState[0][7 + tid] = B.prev_vel[6 * f + tid]; sState[0][20 + tid] = B.cur_vel[6 * f + tid]; }
  for (int64_t i = ob + tid; i < oe; i += GPBA_POSE_THREADS) { B.level[i] = (B.obs_flags[i] & 0x2u) ? 1 : 0; B.kernel_off[i] = 0; B.chi2[i] = 0.0; }
  __syncthreads();
  int cur = 0;

  // ---- record rows of state buffer `s` (threads 0 .. n_cam-1) and, on thread 32, the prior + velocity edges: chi2 ->
  // sScal[2], Ji, Jj, e, Omega e (G2oTypes.cc:100-118).  The two run on different warps.
  auto records_and_prior = [&](int s, bool full, bool with_prior) {
    const double* st = sState[s];
    if (tid < n_cam) {
      const bool gp = tid != n_cam - 1;
      if (full) record_row<true, 24>(gp ? st : nullptr, gp ? st + 7 : nullptr, t1, st + 13, st + 20, t2, B.cam_time[(size_t)f * n_cam + tid], B.cam[tid], sRec[tid]);
      else record_row<false, 24>(gp ? st : nullptr, gp ? st + 7 : nullptr, t1, st + 13, st + 20, t2, B.cam_time[(size_t)f * n_cam + tid], B.cam[tid], sRec[tid]);
    }
    if (with_prior && tid == 32) {
      const SE3 T1 = load_se3(st), T2 = load_se3(st + 13);
      const V6 v1 = load_v6(st + 7), v2 = load_v6(st + 20);
      const SE3 T = se3_mul(se3_inv(T1), T2);
      const V6 xi = se3_log(T);
      const M6 K = RightJacobianPose3Inv(xi);
      const V6 Kv2 = mul(K, v2);
      for (int i = 0; i < 6; ++i) { sE[i] = xi[i] - dt * v1[i]; sE[6 + i] = Kv2[i] - v1[i]; }
      const M6 a = se3Adj(v2);
      const M6 A = scale(-1.0, mul(K, se3_Adj(se3_inv(T))));
      const M6 haA = mul(scale(-0.5, a), A);
      const M6 haK = mul(scale(-0.5, a), K);
      for (int j = 0; j < 144; ++j) { sJi[j] = 0.0; sJj[j] = 0.0; }
      for (int r = 0; r < 6; ++r)
        for (int c = 0; c < 6; ++c) {
          sJi[r * 12 + c] = A(r, c); sJi[(6 + r) * 12 + c] = haA(r, c);
          sJj[r * 12 + c] = K(r, c); sJj[(6 + r) * 12 + c] = haK(r, c); sJj[(6 + r) * 12 + 6 + c] = K(r, c);
        }
      for (int r = 0; r < 6; ++r) { sJi[r * 12 + 6 + r] = -dt; sJi[(6 + r) * 12 + 6 + r] = -1.0; }
      double c2 = 0.0;
      for (int i = 0; i < 6; ++i) {
        sOe[i] = V.qc_inv[i] * o11 * sE[i] + V.qc_inv[i] * o12 * sE[6 + i];
        sOe[6 + i] = V.qc_inv[i] * o12 * sE[i] + V.qc_inv[i] * o22 * sE[6 + i];
      }
      for (int i = 0; i < 12; ++i) c2 += sE[i] * sOe[i];
      if (!fix1) c2 += st[7 + 2] * V.qc_inv[2] * st[7 + 2];   // EdgeVelocity on v1 (inactive when v1 is fixed)
      c2 += st[20 + 2] * V.qc_inv[2] * st[20 + 2];            // EdgeVelocity on v2
      sScal[2] = c2;
    }
    __syncthreads();
  };

  // ---- computeActiveErrors + activeRobustChi2 at state buffer `s` -> sScal[0], stored edge chi2 of the active edges,
  // and in the same sweep the linearisation of that state: S_r (upper triangle, 21) | g_r (6) per camera -> sS
  auto evaluate = [&](int s) {
    records_and_prior(s, true, true);
    double acc = 0.0;
#pragma unroll 1
    for (int c = 0; c < n_cam; ++c) {
      double a27[27];
#pragma unroll
      for (int k = 0; k < 27; ++k) a27[k] = 0.0;
      for (int64_t i = ob + tid; i < oe; i += GPBA_POSE_THREADS) {
        if (B.level[i] || B.obs_cam[i] != c) continue;
        ObsEval<true> E;
        double J1[3][6], Jp[3][3];
        const double w = B.obs_w[i];
        eval_obs<true, true>(V, sRec[c], B.cam[c], B.obs_xw[3 * i], B.obs_xw[3 * i + 1], B.obs_xw[3 * i + 2], B.obs_u[i], B.obs_v[i],
                             (B.obs_ur && c == n_cam - 1) ? B.obs_ur[i] : -1.0, w, B.kernel_off[i] ? 0x4u : 0u, E, J1, Jp);
        B.chi2[i] = E.chi2;
        acc += E.rho;
        const double wr = E.rho1 * w;
        for (int rr = 0; rr < E.rows; ++rr) {
          int k = 0;
#pragma unroll
          for (int m = 0; m < 6; ++m) {
            const double wj = wr * J1[rr][m];
#pragma unroll
            for (int q = m; q < 6; ++q) { a27[k] = fma(wj, J1[rr][q], a27[k]); ++k; }
            a27[21 + m] = fma(-wj, E.e[rr], a27[21 + m]);
          }
        }
      }
#pragma unroll
      for (int k = 0; k < 27; ++k) {
        const double v = warp_sum(a27[k]);
        if (lane == 0) sPart[warp][k] = v;
      }
      __syncthreads();
      if (tid < 27) {
        double v = 0.0;
#pragma unroll
        for (int q = 0; q < GPBA_POSE_THREADS / 32; ++q) v += sPart[q][tid];
        sS[c][tid] = v;
      }
      __syncthreads();
    }
    const double sum = block_sum(acc, sRed);
    if (tid == 0) sScal[0] = sum + sScal[2];
    __syncthreads();
  };

  // ---- buildSystem from the linearisation left by evaluate(): sH (n x n), sb
  auto assemble = [&](int s) {
    for (int j = tid; j < 576; j += GPBA_POSE_THREADS) sH[j] = 0.0;
    if (tid < 24) sb[tid] = 0.0;
    for (int j = tid; j < 144; j += GPBA_POSE_THREADS) {   // Omega J of the prior
      const int r = j / 12, c = j % 12, i = r % 6;
      const double q = V.qc_inv[i];
      if (r < 6) { sOJi[j] = q * (o11 * sJi[i * 12 + c] + o12 * sJi[(6 + i) * 12 + c]); sOJj[j] = q * (o11 * sJj[i * 12 + c] + o12 * sJj[(6 + i) * 12 + c]); }
      else { sOJi[j] = q * (o12 * sJi[i * 12 + c] + o22 * sJi[(6 + i) * 12 + c]); sOJj[j] = q * (o12 * sJj[i * 12 + c] + o22 * sJj[(6 + i) * 12 + c]); }
    }
    __syncthreads();
    // H += M_r^T S_r M_r, b += M_r^T g_r over the free vertex slices (thread = entry of the 24 x 24 product)
    for (int j = tid; j < 576; j += GPBA_POSE_THREADS) {
      const int ap = j / 24, bq = j % 24;
      const int oa = ap < 12 ? o1 : o2, obq = bq < 12 ? o1 : o2;
      if (oa < 0 || obq < 0) continue;
      double hacc = 0.0;
      for (int c = 0; c < n_cam; ++c) {
        const double* M = sRec[c] + GPBA_REC_M;
        double sacc = 0.0;
        for (int m = 0; m < 6; ++m) {
          double tq = 0.0;
          for (int q = 0; q < 6; ++q) {
            const int lo = m < q ? m : q, hi = m < q ? q : m;
            tq = fma(sS[c][lo * 6 - lo * (lo - 1) / 2 + (hi - lo)], M[q * 24 + bq], tq);   // S is stored as its upper triangle
          }
          sacc = fma(M[m * 24 + ap], tq, sacc);
        }
        hacc += sacc;
      }
      sH[(oa + ap % 12) * n + obq + bq % 12] = hacc;
    }
    if (tid >= 64 && tid < 88) {
      const int a = tid - 64;
      const int oa = a < 12 ? o1 : o2;
      if (oa >= 0) {
        double bacc = 0.0;
        for (int c = 0; c < n_cam; ++c) {
          const double* M = sRec[c] + GPBA_REC_M;
          double sacc = 0.0;
          for (int m = 0; m < 6; ++m) sacc = fma(M[m * 24 + a], sS[c][21 + m], sacc);
          bacc += sacc;
        }
        sb[oa + a % 12] = bacc;
      }
    }
    __syncthreads();
    // EdgeGaussianPrior (no kernel)
    for (int j = tid; j < 144; j += GPBA_POSE_THREADS) {
      const int r = j / 12, c = j % 12;
      double sii = 0.0, sjj = 0.0, sij = 0.0;
      for (int k = 0; k < 12; ++k) {
        sii = fma(sJi[k * 12 + r], sOJi[k * 12 + c], sii);
        sjj = fma(sJj[k * 12 + r], sOJj[k * 12 + c], sjj);
        sij = fma(sJi[k * 12 + r], sOJj[k * 12 + c], sij);
      }
      sH[(o2 + r) * n + o2 + c] += sjj;
      if (o1 >= 0) { sH[(o1 + r) * n + o1 + c] += sii; sH[(o1 + r) * n + o2 + c] += sij; sH[(o2 + c) * n + o1 + r] += sij; }
    }
    if (tid >= 160 && tid < 184) {
      const int a = tid - 160, r = a % 12;
      const double* J = a < 12 ? sJi : sJj;
      const int oa = a < 12 ? o1 : o2;
      if (oa >= 0) {
        double sacc = 0.0;
        for (int k = 0; k < 12; ++k) sacc = fma(J[k * 12 + r], -sOe[k], sacc);
        sb[oa + r] += sacc;
      }
    }
    __syncthreads();
    if (tid == 0) {   // EdgeVelocity (G2oTypes.h:496-519)
      const double* st = sState[s];
      if (o1 >= 0) { sH[(o1 + 8) * n + o1 + 8] += V.qc_inv[2]; sb[o1 + 8] -= V.qc_inv[2] * st[7 + 2]; }
      sH[(o2 + 8) * n + o2 + 8] += V.qc_inv[2]; sb[o2 + 8] -= V.qc_inv[2] * st[20 + 2];
    }
    __syncthreads();
  };

  const float chi2Mono[4] = {5.991f, 5.991f, 5.991f, 5.991f};
  const float chi2Stereo[4] = {15.6f, 9.8f, 7.815f, 7.815f};
  int nBadEdges = 0;
#pragma unroll 1
  for (int rnd = 0; rnd < GPBA_POSE_ROUNDS; ++rnd) {
    gpba_lm_trace* tr = B.traces ? B.traces + (size_t)f * GPBA_POSE_ROUNDS + rnd : nullptr;
    if (tr && tid == 0) { tr->n_iters = 0; tr->result = GPBA_RESULT_OK; tr->total_trials = 0; tr->cg_iterations = 0; tr->last_trial_chi2 = 0.0; }
    // ---------------- optimize(10)
    double lambda = 0.0, ni = 2.0;   // thread 0 only
    int nBad = 0, result = GPBA_RESULT_OK, cj = 0;
    bool have_eval = false;          // sScal[0], sS, sRec, prior data describe sState[cur]
#pragma unroll 1
    for (int it = 0; it < 10; ++it) {
      if (!have_eval) evaluate(cur);
      double currentChi = sScal[0], tempChi = currentChi;
      const double iniChi = currentChi;
      assemble(cur);
      if (it == 0 && tid == 0) {
        double mx = 0.0;
        for (int j = 0; j < n; ++j) mx = fmax(mx, fabs(sH[j * n + j]));
        lambda = 1e-5 * mx; ni = 2.0; nBad = 0;
      }
      double rho = 0.0;
      int qmax = 0;
#pragma unroll 1
      for (;;) {
        if (tid == 0) sScal[3] = lambda;
        __syncthreads();
        if (warp == 0) {
          const bool ok = pose_solve(n, sH, sb, sScal[3], sL, sx, lane);
          if (lane == 0) sFlag[0] = ok ? 1 : 0;
          // trial state = state (+) x  (PoseVelocity::Update, G2oTypes.cc:41-46): lane 0 the current frame, lane 1 the previous
          const double* st = sState[cur];
          double* tn = sState[1 - cur];
          if (lane == 0) {
            store_se3(se3_mul(load_se3(st + 13), se3_exp(load_v6(sx + o2))), tn + 13);
            for (int i = 0; i < 6; ++i) tn[20 + i] = st[20 + i] + sx[o2 + 6 + i];
          } else if (lane == 1) {
            if (o1 >= 0) {
              store_se3(se3_mul(load_se3(st), se3_exp(load_v6(sx + o1))), tn);
              for (int i = 0; i < 6; ++i) tn[7 + i] = st[7 + i] + sx[o1 + 6 + i];
            } else {
              for (int i = 0; i < 13; ++i) tn[i] = st[i];
            }
          }
        }
        __syncthreads();
        evaluate(1 - cur);
        if (tid == 0) {
          tempChi = sFlag[0] ? sScal[0] : 1.7976931348623157e308;
          double sc = 0.0;
          for (int j = 0; j < n; ++j) sc += sx[j] * (lambda * sx[j] + sb[j]);
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
        have_eval = sFlag[3] != 0;       // accepted: the trial's evaluation is the next iteration's starting point
        if (sFlag[3]) cur = 1 - cur;     // discardTop(): the trial buffer becomes the estimate
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
    // ---------------- re-flagging (Optimizer.cc:571-664): float chi2; inactive edges are re-evaluated at the estimate,
    // active ones keep the error of the last evaluated trial
    records_and_prior(cur, false, false);
    const float chi2close = (float)(1.5 * chi2Mono[rnd]);
    int bad_local = 0;
    const double* st = sState[cur];
    for (int64_t i = ob + tid; i < oe; i += GPBA_POSE_THREADS) {
      const int c = B.obs_cam[i];
      const double ur = (B.obs_ur && c == n_cam - 1) ? B.obs_ur[i] : -1.0;
      double c2 = B.chi2[i];
      if (B.level[i]) {
        ObsEval<true> E;
        eval_obs<true, false>(V, sRec[c], B.cam[c], B.obs_xw[3 * i], B.obs_xw[3 * i + 1], B.obs_xw[3 * i + 2], B.obs_u[i], B.obs_v[i], ur,
                              B.obs_w[i], 0x4u, E, nullptr, nullptr);
        c2 = E.chi2;
        B.chi2[i] = c2;
      }
      const float chi2 = (float)c2;
      bool bad;
      if (ur >= 0.0) {
        bad = chi2 > chi2Stereo[rnd];
      } else {
        const CamConst& cam = B.cam[c];
        SE3 Tbc;
        Tbc.q.x = cam.qbc[0]; Tbc.q.y = cam.qbc[1]; Tbc.q.z = cam.qbc[2]; Tbc.q.w = cam.qbc[3];
        Tbc.t = v3(cam.tbc[0], cam.tbc[1], cam.tbc[2]);
        const V3 X = v3(B.obs_xw[3 * i], B.obs_xw[3 * i + 1], B.obs_xw[3 * i + 2]);
        bool pos = se3_act(se3_inv(se3_mul(load_se3(st + 13), Tbc)), X)[2] > 0;
        if (c != n_cam - 1) pos = (se3_act(se3_inv(se3_mul(load_se3(st), Tbc)), X)[2] > 0) && pos;
        const bool close = B.obs_flags[i] & 0x1u;
        bad = (chi2 > chi2Mono[rnd] && !close) || (close && chi2 > chi2close) || !pos;
      }
      B.level[i] = bad ? 1 : 0;
      bad_local += bad ? 1 : 0;
      if (rnd == 2) B.kernel_off[i] = 1;
    }
    const double nb = block_sum((double)bad_local, sRed);
    if (tid == 0) sScal[4] = nb;
    __syncthreads();
    nBadEdges = (int)sScal[4];
    if ((oe - ob) + 3 < 10) break;   // optimizer.edges().size() < 10
  }
  if (tid < 7) {
    if (B.out_cur_pose) B.out_cur_pose[7 * f + tid] = sState[cur][13 + tid];
    if (B.out_prev_pose) B.out_prev_pose[7 * f + tid] = sState[cur][tid];
  }
  if (tid < 6) {
    if (B.out_cur_vel) B.out_cur_vel[6 * f + tid] = sState[cur][20 + tid];
    if (B.out_prev_vel) B.out_prev_vel[6 * f + tid] = sState[cur][7 + tid];
  }
  if (tid == 0 && B.out_inliers) B.out_inliers[f] = (int)(oe - ob) - nBadEdges;
}

}  // namespace gpba
