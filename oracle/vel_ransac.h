// oracle/vel_ransac.h -- TEST INFRASTRUCTURE ONLY (included by gpba_oracle.cc).
// CPU restatement of Optimizer::OptimizeVel (src/Optimizer.cc:2364-2447) called once per hypothesis as Tracking::MCRansac
// does (src/Tracking.cc:1939-2002): VertexVel (include/G2oTypes.h:128-145), EdgeVelReproj (G2oTypes.h:521-547,
// src/G2oTypes.cc:497-510), Huber delta 5.991, BlockSolverX + LinearSolverDense under g2o's LM with the default lambda
// (optimization_algorithm_levenberg.cpp:61-194), 40 iterations, then |e| <= threshold over all edges.
// PARITY: EdgeVelReproj (error, Jacobian) is pinned against the reference's own G2oTypes.cc (oracle/_ref,
// tests/test_ref_pin.py), and every hypothesis against the reference's real OptimizeVel graph / solver (ref_g2o_optimize_vel,
// tests/test_whole_path_reference.py: same winner, inlier counts, masks and velocities).
#pragma once
#include <vector>
#include <cmath>
#include <limits>
#include <cstring>

namespace ora {

struct VelHypothesis {
  const gpba_vel_batch* B;
  SE3 Tinv;                     // T^-1 with T = pF2->GetPoseW()
  std::vector<SE3> Tcb;
  std::vector<Pinhole> cams;
  V6 v;
  Huber hub;

  explicit VelHypothesis(const gpba_vel_batch* B_) : B(B_) {
    SE3 T; T.q = {B->last_pose[0], B->last_pose[1], B->last_pose[2], B->last_pose[3]};
    T.t[0] = B->last_pose[4]; T.t[1] = B->last_pose[5]; T.t[2] = B->last_pose[6];
    Tinv = se3_inv(T);
    for (int c = 0; c < B->n_cam; ++c) {
      SE3 Tbc; Tbc.q = {B->cam_Tbc[7 * c], B->cam_Tbc[7 * c + 1], B->cam_Tbc[7 * c + 2], B->cam_Tbc[7 * c + 3]};
      Tbc.t[0] = B->cam_Tbc[7 * c + 4]; Tbc.t[1] = B->cam_Tbc[7 * c + 5]; Tbc.t[2] = B->cam_Tbc[7 * c + 6];
      Tcb.push_back(se3_inv(Tbc));
      Pinhole p; p.fx = B->cam_intr[4 * c]; p.fy = B->cam_intr[4 * c + 1]; p.cx = B->cam_intr[4 * c + 2]; p.cy = B->cam_intr[4 * c + 3];
      cams.push_back(p);
    }
    hub.setDelta(B->huber_delta);
  }
  // EdgeVelReproj::computeError: Xc = (T exp(v dt) Tbc)^-1 Xw = Tbc^-1 exp(-v dt) T^-1 Xw
  void error(int i, const V6& vel, double* e, V3* Xb_out = nullptr, SE3* Tcb1_out = nullptr) const {
    const int cam = B->obs_cam[i];
    V6 ndxi; for (int k = 0; k < 6; ++k) ndxi[k] = -(vel[k] * B->cam_dt[cam]);
    const SE3 Tcb1 = se3_mul(Tcb[cam], se3_exp(ndxi));
    V3 Xw; Xw[0] = B->obs_xw[3 * i]; Xw[1] = B->obs_xw[3 * i + 1]; Xw[2] = B->obs_xw[3 * i + 2];
    const V3 Xb = se3_act(Tinv, Xw);
    const V3 Xc = se3_act(Tcb1, Xb);
    double uv[2];
    cams[cam].project(Xc, uv);
    e[0] = B->obs_u[i] - uv[0]; e[1] = B->obs_v[i] - uv[1];
    if (Xb_out) *Xb_out = Xb;
    if (Tcb1_out) *Tcb1_out = Tcb1;
  }
  // EdgeVelReproj::linearizeOplus (G2oTypes.cc:497-510): J = -proj_jac * (-Tcb1 * CircleDot(Xb) * Jr(-dxi) * dt)[0:3]
  void jacobian(int i, const V6& vel, double J[2][6]) const {
    const int cam = B->obs_cam[i];
    const double dt = B->cam_dt[cam];
    double e[2]; V3 Xb; SE3 Tcb1;
    error(i, vel, e, &Xb, &Tcb1);
    const V3 Xc = se3_act(Tcb1, Xb);
    const Mat<2, 3> P = cams[cam].projectJac(Xc);
    const M3 R = quat_to_R(Tcb1.q);
    const M3 Xh = hat(Xb);
    Mat<3, 6> D;   // R [I, -Xb^]
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) {
        D(r, c) = R(r, c);
        double s = 0; for (int k = 0; k < 3; ++k) s += R(r, k) * Xh(k, c);
        D(r, 3 + c) = -s;
      }
    V6 ndxi; for (int k = 0; k < 6; ++k) ndxi[k] = -(vel[k] * dt);
    const M6 Jr = RightJacobianPose3(ndxi);
    for (int r = 0; r < 2; ++r)
      for (int c = 0; c < 6; ++c) {
        double s = 0;
        for (int a = 0; a < 3; ++a) { double d = 0; for (int k = 0; k < 6; ++k) d += D(a, k) * Jr(k, c); s += P(r, a) * d; }
        J[r][c] = s * dt;   // -proj_jac * (-(...)) * dt
      }
  }
  double active_chi2(const int32_t* set, int ns, const V6& vel) const {
    double sum = 0;
    for (int k = 0; k < ns; ++k) {
      double e[2]; error(set[k], vel, e);
      const double w = B->obs_inv_sigma2[set[k]];
      double rho[3]; hub.robustify(e[0] * w * e[0] + e[1] * w * e[1], rho);
      sum += rho[0];
    }
    return sum;
  }
  // one OptimizeVel call; returns the inlier count
  int run(const int32_t* set, double* vel_out, uint8_t* mask, gpba_lm_trace* tr) {
    const int ns = B->set_size, n = 6;
    for (int k = 0; k < 6; ++k) v[k] = B->vel_init[k];
    if (tr) { std::memset(tr, 0, sizeof(*tr)); tr->result = GPBA_RESULT_OK; }
    double lambda = 0, ni = 2; int nBad = 0, cj = 0, result = GPBA_RESULT_OK;
    std::vector<double> H(36), b(6), x(6, 0.0);
    for (int it = 0; it < B->iterations && result == GPBA_RESULT_OK; ++it, ++cj) {
      double currentChi = active_chi2(set, ns, v), tempChi = currentChi;
      const double iniChi = currentChi;
      std::fill(H.begin(), H.end(), 0.0); std::fill(b.begin(), b.end(), 0.0);
      for (int k = 0; k < ns; ++k) {   // BaseUnaryEdge::constructQuadraticForm (base_unary_edge.hpp:43-72)
        const int i = set[k];
        double e[2], J[2][6]; error(i, v, e); jacobian(i, v, J);
        const double w = B->obs_inv_sigma2[i];
        double rho[3]; hub.robustify(e[0] * w * e[0] + e[1] * w * e[1], rho);
        for (int r = 0; r < 2; ++r)
          for (int a = 0; a < 6; ++a) {
            const double jo = J[r][a] * (rho[1] * w);
            b[a] -= jo * e[r];
            for (int c = 0; c < 6; ++c) H[a * 6 + c] += jo * J[r][c];
          }
      }
      if (it == 0) {
        double mx = 0; for (int j = 0; j < n; ++j) mx = std::max(mx, std::fabs(H[j * 6 + j]));
        lambda = 1e-5 * mx; ni = 2; nBad = 0;
      }
      double rho = 0; int qmax = 0;
      do {
        const V6 backup = v;
        std::vector<double> A(H);
        for (int j = 0; j < n; ++j) A[j * 6 + j] += lambda;
        const bool ok = ldlt_dense(n, A, b.data(), x.data());
        for (int j = 0; j < 6; ++j) v[j] += x[j];   // VertexVel::oplusImpl
        tempChi = active_chi2(set, ns, v);
        if (!ok) tempChi = std::numeric_limits<double>::max();
        double scale = 0; for (int j = 0; j < n; ++j) scale += x[j] * (lambda * x[j] + b[j]);
        rho = (currentChi - tempChi) / (scale + 1e-3);
        if (rho > 0 && std::isfinite(tempChi)) {
          double alpha = 1. - std::pow((2 * rho - 1), 3);
          alpha = std::min(alpha, 2. / 3.);
          lambda *= std::max(1. / 3., alpha);
          ni = 2; currentChi = tempChi;
        } else {
          lambda *= ni; ni *= 2; v = backup;
        }
        ++qmax;
      } while (rho < 0 && qmax < 10);
      if (tr && it < GPBA_MAX_ITERS) {
        tr->levenberg_iterations[it] = qmax; tr->chi2_before[it] = iniChi; tr->chi2_after[it] = currentChi;
        tr->lambda[it] = lambda; tr->total_trials += qmax; tr->last_trial_chi2 = tempChi;
      }
      if (qmax == 10 || rho == 0) result = GPBA_TERMINATE;
      else { if ((iniChi - currentChi) * 1e3 < iniChi) nBad++; else nBad = 0; if (nBad >= 3) result = GPBA_TERMINATE; }
    }
    if (tr) { tr->n_iters = cj; tr->result = result; }
    int inl = 0;
    for (int i = 0; i < B->n_match; ++i) {   // Optimizer.cc:2425-2440
      double e[2]; error(i, v, e);
      const bool in = std::sqrt(e[0] * e[0] + e[1] * e[1]) <= B->threshold;
      if (mask) mask[i] = in ? 1 : 0;
      inl += in;
    }
    if (vel_out) for (int k = 0; k < 6; ++k) vel_out[k] = v[k];
    return inl;
  }
};

}  // namespace ora
