// pose_graph.h -- CPU restatement (TEST INFRASTRUCTURE ONLY; g2o::Sim3 exp/log/product/inverse, the EdgeSim3 error
// and the vertex update are pinned against the reference's own g2o/types/sim3.h compiled into oracle/_ref, tests/test_ref_pin.py;
// the whole optimisation against the reference's real VertexSim3Expmap / EdgeSim3 / BlockSolver_7_3 / LM, tests/test_whole_path_reference.py,
// within the reproducibility band of g2o's numeric Jacobians) of the
// optimisation inside Optimizer::OptimizeEssentialGraph (src/Optimizer.cc:1434-1717): a g2o graph of VertexSim3Expmap
// (Thirdparty/g2o/g2o/types/types_seven_dof_expmap.h:48-96) and EdgeSim3 (:99-126) with identity information, solved by
// BlockSolver_7_3 + LinearSolverEigen + Levenberg-Marquardt with lambda_0 = 1e-16 for 20 iterations (:1441-1448, 1665-1668).
//   g2o::Sim3 (types/sim3.h:41-297)            exp (ctor from Vector7d), log, inverse, product, map
//   EdgeSim3::computeError (:106-114)          e = Log(C * S_i * S_j^-1), vertex 0 = i, vertex 1 = j
//   EdgeSim3 has no linearizeOplus             => BaseBinaryEdge::linearizeOplus, NUMERIC central differences with
//                                                 delta = 1e-9 through oplus (core/base_binary_edge.hpp:131-200)
//   VertexSim3Expmap::oplusImpl (:60-69)       S <- Sim3(update) * S, update[6] = 0 when _fix_scale
//   BaseBinaryEdge::constructQuadraticForm     core/base_binary_edge.hpp:55-120 (no robust kernel)
//   BlockSolver::solve without Schur           core/block_solver.hpp:353-365 (all vertices non-marginalized, :1497)
//   OptimizationAlgorithmLevenberg             core/optimization_algorithm_levenberg.cpp:61-194 (as in gpba_oracle.cc)
//   point correction                           src/Optimizer.cc:1687-1712: P <- S_wr' (S_rw P)
// Eigen pieces restated from their published algorithms: Quaterniond(Matrix3d) (Geometry/Quaternion.h, the trace / largest
// diagonal branches), toRotationMatrix, quaternion product (NOT re-normalised), _transformVector, 3x3 LU solve.
#pragma once
#include <cmath>
#include <cstring>
#include <limits>
#include <vector>

#include "../include/gpba.h"
#include "lie.h"

namespace ora {

struct Sim3 { Quat r; V3 t; double s; };

inline Quat quat_from_R_eigen(const M3& m) {  // Eigen::internal::quaternionbase_assign_impl<Other,3,3>::run
  Quat q;
  double t = m(0, 0) + m(1, 1) + m(2, 2);
  if (t > 0.0) {
    t = std::sqrt(t + 1.0);
    q.w = 0.5 * t;
    t = 0.5 / t;
    q.x = (m(2, 1) - m(1, 2)) * t;
    q.y = (m(0, 2) - m(2, 0)) * t;
    q.z = (m(1, 0) - m(0, 1)) * t;
  } else {
    int i = 0;
    if (m(1, 1) > m(0, 0)) i = 1;
    if (m(2, 2) > m(i, i)) i = 2;
    int j = (i + 1) % 3, k = (j + 1) % 3;
    t = std::sqrt(m(i, i) - m(j, j) - m(k, k) + 1.0);
    double v[3];
    v[i] = 0.5 * t;
    t = 0.5 / t;
    q.w = (m(k, j) - m(j, k)) * t;
    v[j] = (m(j, i) + m(i, j)) * t;
    v[k] = (m(k, i) + m(i, k)) * t;
    q.x = v[0]; q.y = v[1]; q.z = v[2];
  }
  return q;
}
inline Quat quat_mul_raw(const Quat& a, const Quat& b) {  // Eigen quaternion product: no normalisation
  return Quat{a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y, a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z,
              a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x, a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z};
}
inline V3 quat_rot_eigen(const Quat& q, const V3& v) {  // QuaternionBase::_transformVector
  V3 qv; qv[0] = q.x; qv[1] = q.y; qv[2] = q.z;
  V3 uv = cross3(qv, v);
  uv = uv + uv;
  return v + q.w * uv + cross3(qv, uv);
}
inline M3 skew_g2o(const V3& v) { return hat(v); }  // se3_ops.hpp:27-38 == SO3::hat
inline V3 deltaR(const M3& R) { V3 v; v[0] = R(2, 1) - R(1, 2); v[1] = R(0, 2) - R(2, 0); v[2] = R(1, 0) - R(0, 1); return v; }

inline Sim3 sim3_exp(const double* u) {  // Sim3(const Vector7d&), sim3.h:63-134: [omega; upsilon; sigma]
  V3 omega, ups;
  for (int i = 0; i < 3; ++i) { omega[i] = u[i]; ups[i] = u[3 + i]; }
  const double sigma = u[6];
  const double theta = std::sqrt(omega[0] * omega[0] + omega[1] * omega[1] + omega[2] * omega[2]);
  const M3 Om = skew_g2o(omega), Om2 = Om * Om, I = M3::Identity();
  Sim3 S;
  S.s = std::exp(sigma);
  M3 R;
  const double eps = 0.00001;
  double A, B, C;
  if (std::fabs(sigma) < eps) {
    C = 1;
    if (theta < eps) { A = 1. / 2.; B = 1. / 6.; R = I + Om + Om * Om; }
    else {
      const double theta2 = theta * theta;
      A = (1 - std::cos(theta)) / theta2;
      B = (theta - std::sin(theta)) / (theta2 * theta);
      R = I + Om * (std::sin(theta) / theta) + Om2 * ((1 - std::cos(theta)) / (theta * theta));
    }
  } else {
    C = (S.s - 1) / sigma;
    if (theta < eps) {
      const double sigma2 = sigma * sigma;
      A = ((sigma - 1) * S.s + 1) / sigma2;
      B = ((0.5 * sigma2 - sigma + 1) * S.s) / (sigma2 * sigma);
      R = I + Om + Om2;
    } else {
      R = I + Om * (std::sin(theta) / theta) + Om2 * ((1 - std::cos(theta)) / (theta * theta));
      const double a = S.s * std::sin(theta), b = S.s * std::cos(theta);
      const double theta2 = theta * theta, sigma2 = sigma * sigma, c = theta2 + sigma2;
      A = (a * sigma + (1 - b) * theta) / (theta * c);
      B = (C - ((b - 1) * sigma + a * theta) / c) * 1. / theta2;
    }
  }
  S.r = quat_from_R_eigen(R);
  const M3 W = Om * A + Om2 * B + I * C;
  S.t = W * ups;
  return S;
}
inline V3 solve3(const M3& A, const V3& b) {  // W.lu().solve(t): 3 x 3, pivoted elimination (the system is well conditioned)
  double m[3][4];
  for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) m[r][c] = A(r, c); m[r][3] = b[r]; }
  for (int k = 0; k < 3; ++k) {
    int p = k;
    for (int r = k + 1; r < 3; ++r) if (std::fabs(m[r][k]) > std::fabs(m[p][k])) p = r;
    if (p != k) for (int c = 0; c < 4; ++c) std::swap(m[k][c], m[p][c]);
    for (int r = k + 1; r < 3; ++r) { const double f = m[r][k] / m[k][k]; for (int c = k; c < 4; ++c) m[r][c] -= f * m[k][c]; }
  }
  V3 x;
  for (int r = 2; r >= 0; --r) { double s = m[r][3]; for (int c = r + 1; c < 3; ++c) s -= m[r][c] * x[c]; x[r] = s / m[r][r]; }
  return x;
}
inline void sim3_log(const Sim3& S, double* res) {  // Sim3::log, sim3.h:140-216
  const double sigma = std::log(S.s);
  const M3 R = quat_to_R(S.r);
  const double d = 0.5 * (R(0, 0) + R(1, 1) + R(2, 2) - 1);
  const double eps = 0.00001;
  const M3 I = M3::Identity();
  V3 omega;
  M3 Om;
  double A, B, C;
  if (std::fabs(sigma) < eps) {
    C = 1;
    if (d > 1 - eps) { omega = 0.5 * deltaR(R); Om = skew_g2o(omega); A = 1. / 2.; B = 1. / 6.; }
    else {
      const double theta = std::acos(d), theta2 = theta * theta;
      omega = (theta / (2 * std::sqrt(1 - d * d))) * deltaR(R);
      Om = skew_g2o(omega);
      A = (1 - std::cos(theta)) / theta2;
      B = (theta - std::sin(theta)) / (theta2 * theta);
    }
  } else {
    C = (S.s - 1) / sigma;
    if (d > 1 - eps) {
      const double sigma2 = sigma * sigma;
      omega = 0.5 * deltaR(R);
      Om = skew_g2o(omega);
      A = ((sigma - 1) * S.s + 1) / sigma2;
      B = ((0.5 * sigma2 - sigma + 1) * S.s) / (sigma2 * sigma);
    } else {
      const double theta = std::acos(d);
      omega = (theta / (2 * std::sqrt(1 - d * d))) * deltaR(R);
      Om = skew_g2o(omega);
      const double theta2 = theta * theta;
      const double a = S.s * std::sin(theta), b = S.s * std::cos(theta), c = theta2 + sigma * sigma;
      A = (a * sigma + (1 - b) * theta) / (theta * c);
      B = (C - ((b - 1) * sigma + a * theta) / c) * 1. / theta2;
    }
  }
  const M3 W = Om * A + (Om * Om) * B + I * C;
  const V3 ups = solve3(W, S.t);
  for (int i = 0; i < 3; ++i) { res[i] = omega[i]; res[3 + i] = ups[i]; }
  res[6] = sigma;
}
inline Sim3 sim3_inv(const Sim3& S) {  // sim3.h:219-222
  Sim3 r;
  r.r = Quat{-S.r.x, -S.r.y, -S.r.z, S.r.w};
  r.t = quat_rot_eigen(r.r, (-1. / S.s) * S.t);
  r.s = 1. / S.s;
  return r;
}
inline Sim3 sim3_mul(const Sim3& a, const Sim3& b) {  // sim3.h:255-261
  Sim3 r;
  r.r = quat_mul_raw(a.r, b.r);
  r.t = a.s * quat_rot_eigen(a.r, b.t) + a.t;
  r.s = a.s * b.s;
  return r;
}
inline V3 sim3_map(const Sim3& S, const V3& p) { return S.s * quat_rot_eigen(S.r, p) + S.t; }  // sim3.h:136-138
inline Sim3 sim3_from8(const double* p) { Sim3 S; S.r = Quat{p[0], p[1], p[2], p[3]}; S.t[0] = p[4]; S.t[1] = p[5]; S.t[2] = p[6]; S.s = p[7]; return S; }
inline void sim3_to8(const Sim3& S, double* p) { p[0] = S.r.x; p[1] = S.r.y; p[2] = S.r.z; p[3] = S.r.w; p[4] = S.t[0]; p[5] = S.t[1]; p[6] = S.t[2]; p[7] = S.s; }

inline void edge_sim3_error(const Sim3& C, const Sim3& Si, const Sim3& Sj, double* e7) {  // EdgeSim3::computeError
  sim3_log(sim3_mul(sim3_mul(C, Si), sim3_inv(Sj)), e7);
}
inline Sim3 sim3_oplus(const Sim3& S, const double* upd, bool fix_scale) {  // VertexSim3Expmap::oplusImpl
  double u[7];
  for (int i = 0; i < 7; ++i) u[i] = upd[i];
  if (fix_scale) u[6] = 0;
  return sim3_mul(sim3_exp(u), S);
}

struct PoseGraphOracle {
  int n = 0;
  std::vector<Sim3> S;
  std::vector<uint8_t> fixed;
  bool fix_scale = false;
  int64_t n_edge = 0;
  std::vector<int> ei, ej;
  std::vector<Sim3> meas;
  double lambda_init = 1e-16;
  // structure
  std::vector<int> h;   // hessian index or -1
  int np = 0;
  std::vector<double> err;  // 7 per edge

  void load(const gpba_pose_graph* g) {
    n = g->n_kf; n_edge = g->n_edge; fix_scale = g->fix_scale != 0; lambda_init = g->lambda_init;
    S.resize(n); fixed.assign(g->fixed, g->fixed + n);
    for (int i = 0; i < n; ++i) S[i] = sim3_from8(g->sim3 + 8 * i);
    ei.assign(g->edge_i, g->edge_i + n_edge); ej.assign(g->edge_j, g->edge_j + n_edge);
    meas.resize(n_edge);
    for (int64_t k = 0; k < n_edge; ++k) meas[k] = sim3_from8(g->edge_meas + 8 * k);
    err.assign((size_t)n_edge * 7, 0.0);
  }
  bool edge_active(int64_t k) const { return !(fixed[ei[k]] && fixed[ej[k]]); }
  void compute_errors() { for (int64_t k = 0; k < n_edge; ++k) if (edge_active(k)) edge_sim3_error(meas[k], S[ei[k]], S[ej[k]], &err[(size_t)k * 7]); }
  double chi2() const {
    double c = 0;
    for (int64_t k = 0; k < n_edge; ++k) if (edge_active(k)) for (int d = 0; d < 7; ++d) c += err[(size_t)k * 7 + d] * err[(size_t)k * 7 + d];
    return c;
  }
  // numeric Jacobian of edge k wrt vertex `which` (0: i, 1: j), 7 x 7 row-major (base_binary_edge.hpp:147-195)
  void numeric_jacobian(int64_t k, int which, double* J) const {
    const double delta = 1e-9, scalar = 1.0 / (2 * delta);
    double add[7] = {0, 0, 0, 0, 0, 0, 0}, ep[7], em[7];
    for (int d = 0; d < 7; ++d) {
      add[d] = delta;
      Sim3 Sp = sim3_oplus(S[which ? ej[k] : ei[k]], add, fix_scale);
      edge_sim3_error(meas[k], which ? S[ei[k]] : Sp, which ? Sp : S[ej[k]], ep);
      add[d] = -delta;
      Sim3 Sm = sim3_oplus(S[which ? ej[k] : ei[k]], add, fix_scale);
      edge_sim3_error(meas[k], which ? S[ei[k]] : Sm, which ? Sm : S[ej[k]], em);
      add[d] = 0.0;
      for (int r = 0; r < 7; ++r) J[r * 7 + d] = scalar * (ep[r] - em[r]);
    }
  }
  void build_structure() {
    // a vertex is active iff it has an active edge; free vertices in ascending id = Hessian order
    std::vector<char> act(n, 0);
    for (int64_t k = 0; k < n_edge; ++k) if (edge_active(k)) { act[ei[k]] = 1; act[ej[k]] = 1; }
    h.assign(n, -1); np = 0;
    for (int i = 0; i < n; ++i) if (act[i] && !fixed[i]) h[i] = np++;
  }
  // dense system H (7 np)^2, b
  void build_system(std::vector<double>& H, std::vector<double>& b) const {
    const int N = 7 * np;
    H.assign((size_t)N * N, 0.0); b.assign(N, 0.0);
    double Ji[49], Jj[49];
    for (int64_t k = 0; k < n_edge; ++k) {
      if (!edge_active(k)) continue;
      const int hi = h[ei[k]], hj = h[ej[k]];
      const double* e = &err[(size_t)k * 7];
      if (hi >= 0) numeric_jacobian(k, 0, Ji);
      if (hj >= 0) numeric_jacobian(k, 1, Jj);
      auto add_block = [&](int ha, const double* A, int hb, const double* B) {  // H[ha, hb] += A^T B
        for (int r = 0; r < 7; ++r)
          for (int c = 0; c < 7; ++c) {
            double s = 0;
            for (int d = 0; d < 7; ++d) s += A[d * 7 + r] * B[d * 7 + c];
            H[(size_t)(7 * ha + r) * N + 7 * hb + c] += s;
          }
      };
      auto add_b = [&](int ha, const double* A) {
        for (int r = 0; r < 7; ++r) { double s = 0; for (int d = 0; d < 7; ++d) s += A[d * 7 + r] * (-e[d]); b[7 * ha + r] += s; }
      };
      if (hi >= 0) { add_block(hi, Ji, hi, Ji); add_b(hi, Ji); }
      if (hj >= 0) { add_block(hj, Jj, hj, Jj); add_b(hj, Jj); }
      if (hi >= 0 && hj >= 0) { add_block(hi, Ji, hj, Jj); add_block(hj, Jj, hi, Ji); }
    }
  }
  int optimize(int iterations, const gpba_lm_params& P, gpba_lm_trace* tr, bool (*ldlt)(int, std::vector<double>&, const double*, double*)) {
    if (tr) { std::memset(tr, 0, sizeof(*tr)); tr->result = GPBA_RESULT_OK; }
    build_structure();
    double lambda = -1, ni = 2;
    int nBad = 0, cj = 0, result = GPBA_RESULT_OK;
    std::vector<double> H, b, Hl, x;
    bool ok = true;
    for (int it = 0; it < iterations && ok; ++it) {
      compute_errors();
      double currentChi = chi2(), tempChi = currentChi;
      const double iniChi = currentChi;
      build_system(H, b);
      const int N = 7 * np;
      if (it == 0) {
        if (lambda_init > 0) lambda = lambda_init;
        else { double mx = 0; for (int i = 0; i < N; ++i) mx = std::max(mx, std::fabs(H[(size_t)i * N + i])); lambda = P.tau * mx; }
        ni = 2; nBad = 0;
      }
      double rho = 0;
      int qmax = 0;
      do {
        std::vector<Sim3> backup = S;   // push
        Hl = H;
        for (int i = 0; i < N; ++i) Hl[(size_t)i * N + i] += lambda;
        x.assign(N, 0.0);
        const bool ok2 = N == 0 ? true : ldlt(N, Hl, b.data(), x.data());
        for (int i = 0; i < n; ++i) if (h[i] >= 0) S[i] = sim3_oplus(S[i], &x[(size_t)7 * h[i]], fix_scale);
        compute_errors();
        tempChi = chi2();
        if (!ok2) tempChi = std::numeric_limits<double>::max();
        rho = currentChi - tempChi;
        double scale = 0;
        for (int i = 0; i < N; ++i) scale += x[i] * (lambda * x[i] + b[i]);
        scale += 1e-3;
        rho /= scale;
        if (rho > 0 && std::isfinite(tempChi)) {
          double alpha = 1. - std::pow(2 * rho - 1, 3);
          alpha = std::min(alpha, P.good_step_upper);
          lambda *= std::max(P.good_step_lower, alpha);
          ni = 2;
          currentChi = tempChi;   // discardTop
        } else {
          lambda *= ni; ni *= 2;
          S = backup;             // pop
        }
        qmax++;
      } while (rho < 0 && qmax < P.max_trials_after_failure);
      if (tr && it < GPBA_MAX_ITERS) {
        tr->levenberg_iterations[it] = qmax; tr->chi2_before[it] = iniChi; tr->chi2_after[it] = currentChi; tr->lambda[it] = lambda;
        tr->total_trials += qmax; tr->last_trial_chi2 = tempChi;
      }
      ++cj;
      if (qmax == P.max_trials_after_failure || rho == 0) { result = GPBA_TERMINATE; ok = false; }
      else {
        if ((iniChi - currentChi) * 1e3 < iniChi) nBad++; else nBad = 0;
        if (nBad >= 3) { result = GPBA_TERMINATE; ok = false; }
      }
    }
    if (tr) { tr->n_iters = cj; tr->result = result; }
    return cj;
  }
};

}  // namespace ora
