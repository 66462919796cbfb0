// oracle/pose_only.h -- TEST INFRASTRUCTURE ONLY (included by gpba_oracle.cc).
// CPU restatement of Optimizer::PoseGPOptimizationFromeLastFrame (src/Optimizer.cc:369-686): a 2-vertex g2o graph
// (VertexPoseVel of the previous frame, fixed or not, and of the current frame), EdgeMonoGPOnlyPose / EdgeMonoOnlyPose /
// EdgeStereoOnlyPose against constant world points (src/G2oTypes.cc:120-223: same pose Jacobians as the BA edges, no
// point vertex), EdgeGaussianPrior without kernel, EdgeVelocity on both vertices, BlockSolverX + LinearSolverDense
// (no marginalized vertex => no Schur complement, optimization_algorithm_with_hessian.cpp:50-73) under g2o's LM
// (optimization_algorithm_levenberg.cpp:61-194) with the default lambda (computeLambdaInit: tau * max diag, :171-185),
// four rounds with float-typed chi2 tests.  PARITY: the edges (EdgeMonoGPOnlyPose, EdgeMonoOnlyPose,
// EdgeStereoOnlyPose, EdgeGaussianPrior, EdgeVelocity) and the update are pinned against the reference's own G2oTypes.cc
// (oracle/_ref, tests/test_ref_pin.py), and the whole function against the reference's real graph / solver run with the four rounds
// around it (oracle/ref_g2o_run.cc ref_g2o_pose_optimize, tests/test_whole_path_reference.py: every match classified identically, same states).
#pragma once
#include <vector>
#include <cmath>
#include <limits>
#include <cstring>

namespace ora {

inline SE3 load_se3(const double* p) { SE3 T; T.q = {p[0], p[1], p[2], p[3]}; T.t[0] = p[4]; T.t[1] = p[5]; T.t[2] = p[6]; return T; }
inline V6 load_v6(const double* p) { V6 v; for (int i = 0; i < 6; ++i) v[i] = p[i]; return v; }

struct PoseOnlyFrame {
  const gpba_pose_batch* B;
  int f;
  GaussianProcess G;
  std::vector<SE3> Tbc;
  std::vector<Pinhole> cams;
  KfState s1, s2;
  bool fix1;
  int64_t ob, oe;
  std::vector<uint8_t> level, kernel_off;   // per match
  std::vector<double> err;                  // stored _error (3 per match), updated by compute_errors on ACTIVE edges
  Huber hub_mono, hub_stereo;
  int n;                                    // free dimension: 12 or 24
  std::vector<double> H, b, x;

  PoseOnlyFrame(const gpba_pose_batch* B_, int f_) : B(B_), f(f_) {
    G.set_diag(B->qc);
    for (int c = 0; c < B->n_cam; ++c) {
      Tbc.push_back(load_se3(B->cam_Tbc + 7 * c));
      Pinhole p; p.fx = B->cam_intr[4 * c]; p.fy = B->cam_intr[4 * c + 1]; p.cx = B->cam_intr[4 * c + 2]; p.cy = B->cam_intr[4 * c + 3];
      cams.push_back(p);
    }
    s1.Twb = load_se3(B->prev_pose + 7 * f); s1.vel = load_v6(B->prev_vel + 6 * f); s1.time = B->prev_time[f];
    s2.Twb = load_se3(B->cur_pose + 7 * f); s2.vel = load_v6(B->cur_vel + 6 * f); s2.time = B->cur_time[f];
    fix1 = B->prev_fixed[f] != 0;
    ob = B->obs_begin[f]; oe = B->obs_begin[f + 1];
    level.resize(oe - ob); kernel_off.assign(oe - ob, 0); err.assign(3 * (oe - ob), 0.0);
    for (int64_t i = ob; i < oe; ++i) level[i - ob] = (B->obs_flags[i] & GPBA_OBS_LEVEL1) ? 1 : 0;
    hub_mono.setDelta(B->huber_mono); hub_stereo.setDelta(B->huber_stereo);
    n = fix1 ? 12 : 24;
  }
  bool is_gp(int64_t i) const { return B->obs_cam[i] != B->n_cam - 1; }
  int dim(int64_t i) const { return (!is_gp(i) && B->obs_ur && B->obs_ur[i] >= 0) ? 3 : 2; }
  void edge_error(int64_t i, const KfState& a, const KfState& c, double* e) const {
    const int cam = B->obs_cam[i];
    const double obs[3] = {B->obs_u[i], B->obs_v[i], B->obs_ur ? B->obs_ur[i] : -1.0};
    V3 Xw; Xw[0] = B->obs_xw[3 * i]; Xw[1] = B->obs_xw[3 * i + 1]; Xw[2] = B->obs_xw[3 * i + 2];
    reproj_error(G, is_gp(i), dim(i), &a, c, B->cam_time[(size_t)f * B->n_cam + cam], Tbc[cam], cams[cam], B->bf, Xw, obs, e);
  }
  double edge_chi2(int64_t i) const {
    const double* e = &err[3 * (i - ob)];
    double s = 0;
    for (int r = 0; r < dim(i); ++r) s += e[r] * (B->obs_inv_sigma2[i] * e[r]);
    return s;
  }
  // computeActiveErrors + activeRobustChi2 (sparse_optimizer.cpp:61-114)
  double compute_errors() {
    double sum = 0;
    for (int64_t i = ob; i < oe; ++i) {
      if (level[i - ob]) continue;
      edge_error(i, s1, s2, &err[3 * (i - ob)]);
      const double c2 = edge_chi2(i);
      if (kernel_off[i - ob]) sum += c2;
      else { double rho[3]; (dim(i) == 3 ? hub_stereo : hub_mono).robustify(c2, rho); sum += rho[0]; }
    }
    double e12[12];
    prior_error(s1, s2, e12);
    const M12 Om = G.QiInv(s2.time - s1.time);
    for (int r = 0; r < 12; ++r) for (int c = 0; c < 12; ++c) sum += e12[r] * Om(r, c) * e12[c];
    const double qv = G.QcInv(2, 2);
    if (!fix1) sum += s1.vel[2] * qv * s1.vel[2];   // EdgeVelocity on v1 is inactive when v1 is fixed (allVerticesFixed)
    sum += s2.vel[2] * qv * s2.vel[2];
    return sum;
  }
  // BlockSolver::buildSystem without landmarks: H (n x n, row-major, full), b
  void build_system() {
    H.assign((size_t)n * n, 0.0); b.assign(n, 0.0);
    const int o1 = fix1 ? -1 : 0, o2 = fix1 ? 0 : 12;
    auto add = [&](int oa, const double* Ja, int ob_, const double* Jb, int rows, const double* Om /* rows x rows */, const double* e, double w) {
      // H_ab += Ja^T (w Om) Jb ; b_a += -Ja^T (w Om) e   (base_binary_edge.hpp:55-120)
      for (int r = 0; r < rows; ++r)
        for (int c = 0; c < rows; ++c) {
          const double o = w * Om[r * rows + c];
          if (o == 0.0) continue;
          if (oa >= 0) for (int i = 0; i < 12; ++i) {
            const double jo = Ja[r * 12 + i] * o;
            b[oa + i] -= jo * e[c];
            for (int j = 0; j < 12; ++j) H[(size_t)(oa + i) * n + oa + j] += jo * Ja[c * 12 + j];
            if (ob_ >= 0 && Jb) for (int j = 0; j < 12; ++j) { H[(size_t)(oa + i) * n + ob_ + j] += jo * Jb[c * 12 + j]; H[(size_t)(ob_ + j) * n + oa + i] += jo * Jb[c * 12 + j]; }
          }
          if (ob_ >= 0 && Jb) for (int i = 0; i < 12; ++i) {
            const double jo = Jb[r * 12 + i] * o;
            b[ob_ + i] -= jo * e[c];
            for (int j = 0; j < 12; ++j) H[(size_t)(ob_ + i) * n + ob_ + j] += jo * Jb[c * 12 + j];
          }
        }
    };
    for (int64_t i = ob; i < oe; ++i) {
      if (level[i - ob]) continue;
      const int cam = B->obs_cam[i], d = dim(i);
      V3 Xw; Xw[0] = B->obs_xw[3 * i]; Xw[1] = B->obs_xw[3 * i + 1]; Xw[2] = B->obs_xw[3 * i + 2];
      double J1[36] = {0}, J2[36] = {0}, Jp[9];
      reproj_jacobian(G, is_gp(i), d, &s1, s2, B->cam_time[(size_t)f * B->n_cam + cam], Tbc[cam], cams[cam], B->bf, Xw, J1, J2, Jp);
      const double c2 = edge_chi2(i);
      double rho[3] = {c2, 1.0, 0.0};
      if (!kernel_off[i - ob]) (d == 3 ? hub_stereo : hub_mono).robustify(c2, rho);
      double Om[9] = {0};
      for (int r = 0; r < d; ++r) Om[r * d + r] = B->obs_inv_sigma2[i];
      if (is_gp(i)) add(o1, J1, o2, J2, d, Om, &err[3 * (i - ob)], rho[1]);
      else add(o2, J2, -1, nullptr, d, Om, &err[3 * (i - ob)], rho[1]);
    }
    {  // EdgeGaussianPrior (no kernel)
      double e12[12];
      prior_error(s1, s2, e12);
      M12 Ji, Jj;
      prior_jacobian(s1, s2, &Ji, &Jj);
      const M12 Om = G.QiInv(s2.time - s1.time);
      double ji[144], jj[144], om[144];
      for (int r = 0; r < 12; ++r) for (int c = 0; c < 12; ++c) { ji[r * 12 + c] = Ji(r, c); jj[r * 12 + c] = Jj(r, c); om[r * 12 + c] = Om(r, c); }
      if (o1 >= 0) add(o1, ji, o2, jj, 12, om, e12, 1.0);
      else add(o2, jj, -1, nullptr, 12, om, e12, 1.0);
    }
    const double qv = G.QcInv(2, 2);
    if (o1 >= 0) { H[(size_t)(o1 + 8) * n + o1 + 8] += qv; b[o1 + 8] -= qv * s1.vel[2]; }
    H[(size_t)(o2 + 8) * n + o2 + 8] += qv; b[o2 + 8] -= qv * s2.vel[2];
  }
  static void oplus(KfState& s, const double* d) {  // PoseVelocity::Update (G2oTypes.cc:41-46)
    V6 xi; for (int i = 0; i < 6; ++i) xi[i] = d[i];
    s.Twb = se3_mul(s.Twb, se3_exp(xi));
    for (int i = 0; i < 6; ++i) s.vel[i] += d[6 + i];
  }
  // SparseOptimizer::optimize(iters) with OptimizationAlgorithmLevenberg
  void optimize(int iters, gpba_lm_trace* tr) {
    if (tr) { std::memset(tr, 0, sizeof(*tr)); tr->result = GPBA_RESULT_OK; }
    double lambda = 0, ni = 2; int nBad = 0;
    int cj = 0, result = GPBA_RESULT_OK;
    x.assign(n, 0.0);
    for (int it = 0; it < iters && result == GPBA_RESULT_OK; ++it, ++cj) {
      double currentChi = compute_errors(), tempChi = currentChi;
      const double iniChi = currentChi;
      build_system();
      if (it == 0) {
        double mx = 0;
        for (int j = 0; j < n; ++j) mx = std::max(mx, std::fabs(H[(size_t)j * n + j]));
        lambda = 1e-5 * mx; ni = 2; nBad = 0;
      }
      double rho = 0; int qmax = 0;
      do {
        const KfState b1 = s1, b2 = s2;                 // push()
        std::vector<double> A(H);
        for (int j = 0; j < n; ++j) A[(size_t)j * n + j] += lambda;
        const bool ok2 = ldlt_dense(n, A, b.data(), x.data());
        if (!fix1) { oplus(s1, x.data()); oplus(s2, x.data() + 12); } else oplus(s2, x.data());
        tempChi = compute_errors();
        if (!ok2) tempChi = std::numeric_limits<double>::max();
        double scale = 0;
        for (int j = 0; j < n; ++j) scale += x[j] * (lambda * x[j] + b[j]);
        rho = (currentChi - tempChi) / (scale + 1e-3);
        if (rho > 0 && std::isfinite(tempChi)) {
          double alpha = 1. - std::pow((2 * rho - 1), 3);
          alpha = std::min(alpha, 2. / 3.);
          lambda *= std::max(1. / 3., alpha);
          ni = 2; currentChi = tempChi;
        } else {
          lambda *= ni; ni *= 2;
          s1 = b1; s2 = b2;                             // pop(): estimates back, stored errors stay (stale-error quirk)
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
  }
  // the four rounds of Optimizer.cc:545-670; returns nInitialCorrespondences - nBad
  int run(gpba_lm_trace* traces, uint8_t* outlier_out) {
    const float chi2Mono[4] = {5.991f, 5.991f, 5.991f, 5.991f};
    const float chi2Stereo[4] = {15.6f, 9.8f, 7.815f, 7.815f};
    int nBad = 0;
    for (int it = 0; it < 4; ++it) {
      optimize(10, traces ? traces + it : nullptr);
      nBad = 0;
      const float chi2close = (float)(1.5 * chi2Mono[it]);
      for (int64_t i = ob; i < oe; ++i) {
        const int cam = B->obs_cam[i];
        if (level[i - ob]) edge_error(i, s1, s2, &err[3 * (i - ob)]);   // "if (mvbOutlier[idx]) e->computeError()"
        const float chi2 = (float)edge_chi2(i);
        bool bad;
        V3 Xw; Xw[0] = B->obs_xw[3 * i]; Xw[1] = B->obs_xw[3 * i + 1]; Xw[2] = B->obs_xw[3 * i + 2];
        if (dim(i) == 3) bad = chi2 > chi2Stereo[it];
        else {
          const bool bclose = B->obs_flags[i] & GPBA_OBS_CLOSE;
          const bool pos = is_gp(i) ? (depth_positive(s1.Twb, Tbc[cam], Xw) && depth_positive(s2.Twb, Tbc[cam], Xw))
                                    : depth_positive(s2.Twb, Tbc[B->n_cam - 1], Xw);
          bad = (chi2 > chi2Mono[it] && !bclose) || (bclose && chi2 > chi2close) || !pos;
        }
        level[i - ob] = bad ? 1 : 0;
        nBad += bad;
        if (it == 2) kernel_off[i - ob] = 1;
      }
      if ((oe - ob) + 3 < 10) break;   // optimizer.edges().size() < 10
    }
    if (outlier_out) for (int64_t i = ob; i < oe; ++i) outlier_out[i] = level[i - ob];
    return (int)(oe - ob) - nBad;
  }
};

}  // namespace ora
