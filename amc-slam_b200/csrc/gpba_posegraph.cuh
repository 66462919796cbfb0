// gpba_posegraph.cuh -- essential-graph optimisation: the solve inside Optimizer::OptimizeEssentialGraph
// (src/Optimizer.cc:1434-1717) on the device.
//   g2o::Sim3 (Thirdparty/g2o/g2o/types/sim3.h:41-297)             sim3_exp / sim3_log / sim3_mul / sim3_inv / sim3_map
//   VertexSim3Expmap::oplusImpl (types_seven_dof_expmap.h:60-69)   S <- Sim3(update) S, update[6] = 0 when _fix_scale
//   EdgeSim3::computeError (:106-114)                              e = Log(C S_i S_j^-1)
//   BaseBinaryEdge::linearizeOplus (core/base_binary_edge.hpp:131-200): EdgeSim3 has no analytic Jacobian, g2o differentiates
//                                   numerically, central differences with delta = 1e-9 through oplus -- and so does K_lin:
//                                   the LM trajectory of the reference is the trajectory of THIS Jacobian
//   BaseBinaryEdge::constructQuadraticForm (:55-120), information = I_7, no robust kernel (Optimizer.cc:1505, 1535-1541)
// A pose graph has no landmarks, so there is no Schur complement: the Hessian blocks go straight into the tile Cholesky of
// gpba_chol.cuh (a keyframe occupies a 12-slot of the reduced-system layout: seven dimensions of the Sim(3) tangent in front
// of five padding dimensions with a unit diagonal), lambda on the real diagonal as BlockSolver::setLambda does.
// Every edge is one thread: 28 error evaluations (2 vertices x 7 dimensions x +-delta) of ~150 flops each.
#pragma once
#include "gpba_kernels.cuh"

namespace gpba {

struct Sim3 { Quat r; V3 t; double s; };

GPBA_HD Quat quat_from_R_eigen(const M3& m) {  // Eigen quaternionbase_assign_impl<Other,3,3>::run (Geometry/Quaternion.h)
  Quat q;
  double t = m(0, 0) + m(1, 1) + m(2, 2);
  if (t > 0.0) {
    t = sqrt(t + 1.0);
    q.w = 0.5 * t;
    t = 0.5 / t;
    q.x = (m(2, 1) - m(1, 2)) * t;
    q.y = (m(0, 2) - m(2, 0)) * t;
    q.z = (m(1, 0) - m(0, 1)) * t;
  } else {
    int i = 0;
    if (m(1, 1) > m(0, 0)) i = 1;
    if (m(2, 2) > m(i, i)) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    t = sqrt(m(i, i) - m(j, j) - m(k, k) + 1.0);
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
GPBA_HD Quat quat_mul_raw(const Quat& a, const Quat& b) {  // Eigen's quaternion product does not re-normalise
  Quat r;
  r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
  r.y = a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z;
  r.z = a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x;
  r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
  return r;
}
GPBA_HD V3 quat_rot_eigen(const Quat& q, const V3& v) {  // QuaternionBase::_transformVector
  const V3 qv = v3(q.x, q.y, q.z);
  V3 uv = cross3(qv, v);
  uv = add(uv, uv);
  return add(add(v, scale(q.w, uv)), cross3(qv, uv));
}
GPBA_HD V3 deltaR(const M3& R) { return v3(R(2, 1) - R(1, 2), R(0, 2) - R(2, 0), R(1, 0) - R(0, 1)); }  // se3_ops.hpp:40-47

// A, B, C of W = A Omega + B Omega^2 + C I (sim3.h:80-128 and :160-206 are the same four branches)
GPBA_HD void sim3_abc(double sigma, double s, double theta, bool small_angle, double& A, double& B, double& C) {
  const double eps = 0.00001;
  if (fabs(sigma) < eps) {
    C = 1;
    if (small_angle) { A = 1. / 2.; B = 1. / 6.; }
    else {
      const double theta2 = theta * theta;
      A = (1 - cos(theta)) / theta2;
      B = (theta - sin(theta)) / (theta2 * theta);
    }
  } else {
    C = (s - 1) / sigma;
    if (small_angle) {
      const double sigma2 = sigma * sigma;
      A = ((sigma - 1) * s + 1) / sigma2;
      B = ((0.5 * sigma2 - sigma + 1) * s) / (sigma2 * sigma);
    } else {
      const double a = s * sin(theta), b = s * cos(theta);
      const double theta2 = theta * theta, c = theta2 + sigma * sigma;
      A = (a * sigma + (1 - b) * theta) / (theta * c);
      B = (C - ((b - 1) * sigma + a * theta) / c) * 1. / theta2;
    }
  }
}
GPBA_HD Sim3 sim3_exp(const double* u) {  // Sim3(const Vector7d&): [omega; upsilon; sigma]
  const V3 omega = v3(u[0], u[1], u[2]), ups = v3(u[3], u[4], u[5]);
  const double sigma = u[6];
  const double theta = sqrt(omega[0] * omega[0] + omega[1] * omega[1] + omega[2] * omega[2]);
  const M3 Om = hat(omega), Om2 = mul(Om, Om), I = eye<3>();
  Sim3 S;
  S.s = exp(sigma);
  const double eps = 0.00001;
  const bool small_angle = theta < eps;
  double A, B, C;
  sim3_abc(sigma, S.s, theta, small_angle, A, B, C);
  M3 R;
  if (small_angle) R = add(add(I, Om), Om2);
  else R = add(add(I, scale(sin(theta) / theta, Om)), scale((1 - cos(theta)) / (theta * theta), Om2));
  S.r = quat_from_R_eigen(R);
  const M3 W = add(add(scale(A, Om), scale(B, Om2)), scale(C, I));
  S.t = mul(W, ups);
  return S;
}
GPBA_HD V3 solve3(const M3& A, const V3& b) {  // W.lu().solve(t): pivoted elimination
  double m[3][4];
  for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) m[r][c] = A(r, c); m[r][3] = b[r]; }
  for (int k = 0; k < 3; ++k) {
    int p = k;
    for (int r = k + 1; r < 3; ++r) if (fabs(m[r][k]) > fabs(m[p][k])) p = r;
    if (p != k) for (int c = 0; c < 4; ++c) { const double t = m[k][c]; m[k][c] = m[p][c]; m[p][c] = t; }
    for (int r = k + 1; r < 3; ++r) { const double f = m[r][k] / m[k][k]; for (int c = k; c < 4; ++c) m[r][c] -= f * m[k][c]; }
  }
  V3 x;
  for (int r = 2; r >= 0; --r) { double s = m[r][3]; for (int c = r + 1; c < 3; ++c) s -= m[r][c] * x[c]; x[r] = s / m[r][r]; }
  return x;
}
GPBA_HD void sim3_log(const Sim3& S, double* res) {  // Sim3::log
  const double sigma = log(S.s);
  const M3 R = quat_to_R(S.r);
  const double d = 0.5 * (R(0, 0) + R(1, 1) + R(2, 2) - 1);
  const double eps = 0.00001;
  const bool small_angle = d > 1 - eps;
  V3 omega;
  double theta = 0.0;
  if (small_angle) omega = scale(0.5, deltaR(R));
  else {
    theta = acos(d);
    omega = scale(theta / (2 * sqrt(1 - d * d)), deltaR(R));
  }
  double A, B, C;
  sim3_abc(sigma, S.s, theta, small_angle, A, B, C);
  const M3 Om = hat(omega);
  const M3 W = add(add(scale(A, Om), scale(B, mul(Om, Om))), scale(C, eye<3>()));
  const V3 ups = solve3(W, S.t);
  for (int i = 0; i < 3; ++i) { res[i] = omega[i]; res[3 + i] = ups[i]; }
  res[6] = sigma;
}
GPBA_HD Sim3 sim3_inv(const Sim3& S) {
  Sim3 r;
  r.r.x = -S.r.x; r.r.y = -S.r.y; r.r.z = -S.r.z; r.r.w = S.r.w;
  r.t = quat_rot_eigen(r.r, scale(-1. / S.s, S.t));
  r.s = 1. / S.s;
  return r;
}
GPBA_HD Sim3 sim3_mul(const Sim3& a, const Sim3& b) {
  Sim3 r;
  r.r = quat_mul_raw(a.r, b.r);
  r.t = add(scale(a.s, quat_rot_eigen(a.r, b.t)), a.t);
  r.s = a.s * b.s;
  return r;
}
GPBA_HD V3 sim3_map(const Sim3& S, const V3& p) { return add(scale(S.s, quat_rot_eigen(S.r, p)), S.t); }
GPBA_HD Sim3 load_sim3(const double* p) {
  Sim3 S;
  S.r.x = p[0]; S.r.y = p[1]; S.r.z = p[2]; S.r.w = p[3];
  S.t = v3(p[4], p[5], p[6]);
  S.s = p[7];
  return S;
}
GPBA_HD void store_sim3(const Sim3& S, double* p) {
  p[0] = S.r.x; p[1] = S.r.y; p[2] = S.r.z; p[3] = S.r.w; p[4] = S.t[0]; p[5] = S.t[1]; p[6] = S.t[2]; p[7] = S.s;
}
GPBA_HD void edge_sim3_error(const Sim3& C, const Sim3& Si, const Sim3& Sj, double* e7) {
  sim3_log(sim3_mul(sim3_mul(C, Si), sim3_inv(Sj)), e7);
}
GPBA_HD Sim3 sim3_oplus(const Sim3& S, const double* upd, bool fix_scale) {
  double u[7];
  for (int i = 0; i < 7; ++i) u[i] = upd[i];
  if (fix_scale) u[6] = 0;
  return sim3_mul(sim3_exp(u), S);
}

struct PgView {
  int n_kf, fix_scale;
  int64_t n_edge;
  const int* ei; const int* ej;
  const double* meas;             // [n_edge][8]
  const unsigned char* fixed;
  const int* h;                   // [n_kf] Hessian index or -1
  const int* blk_ii; const int* blk_jj; const int* blk_ij;   // Hessian block of every edge (-1: none; bit 30 of blk_ij: transposed)
};

// computeActiveErrors + activeRobustChi2: chi2 = e^T e per active edge, block partial sums
__global__ void __launch_bounds__(128) k_pg_errors(PgView G, const double* __restrict__ S, double* __restrict__ partial) {
  __shared__ double red[32];
  double acc = 0.0;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < G.n_edge; k += (int64_t)gridDim.x * blockDim.x) {
    const int i = G.ei[k], j = G.ej[k];
    if (G.fixed[i] && G.fixed[j]) continue;   // allVerticesFixed: inactive
    double e[7];
    edge_sim3_error(load_sim3(G.meas + 8 * k), load_sim3(S + 8 * i), load_sim3(S + 8 * j), e);
    for (int d = 0; d < 7; ++d) acc = fma(e[d], e[d], acc);
  }
  const double s = block_sum(acc, red);
  if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

// buildSystem: numeric Jacobians + quadratic form of every active edge, accumulated into the 12 x 12 padded blocks
__global__ void __launch_bounds__(64) k_pg_linearize(PgView G, const double* __restrict__ S, double* __restrict__ hpp, double* __restrict__ bp) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= G.n_edge) return;
  const int i = G.ei[k], j = G.ej[k];
  const int hi = G.h[i], hj = G.h[j];
  if (hi < 0 && hj < 0) return;
  const Sim3 C = load_sim3(G.meas + 8 * k), Si = load_sim3(S + 8 * i), Sj = load_sim3(S + 8 * j);
  double e[7];
  edge_sim3_error(C, Si, Sj, e);
  const double delta = 1e-9, scalar = 1.0 / (2 * delta);
  double Ji[49], Jj[49];
  for (int which = 0; which < 2; ++which) {
    if ((which ? hj : hi) < 0) continue;
    double* J = which ? Jj : Ji;
    double add_v[7] = {0, 0, 0, 0, 0, 0, 0}, ep[7], em[7];
    for (int d = 0; d < 7; ++d) {
      add_v[d] = delta;
      const Sim3 Sp = sim3_oplus(which ? Sj : Si, add_v, G.fix_scale != 0);
      edge_sim3_error(C, which ? Si : Sp, which ? Sp : Sj, ep);
      add_v[d] = -delta;
      const Sim3 Sm = sim3_oplus(which ? Sj : Si, add_v, G.fix_scale != 0);
      edge_sim3_error(C, which ? Si : Sm, which ? Sm : Sj, em);
      add_v[d] = 0.0;
      for (int r = 0; r < 7; ++r) J[r * 7 + d] = scalar * (ep[r] - em[r]);
    }
  }
  auto add_block = [&](int blk, const double* A, const double* B, bool transposed) {   // block += A^T B
    double* H = hpp + (size_t)blk * 144;
    for (int r = 0; r < 7; ++r)
      for (int c = 0; c < 7; ++c) {
        double s = 0.0;
        for (int d = 0; d < 7; ++d) s = fma(A[d * 7 + r], B[d * 7 + c], s);
        atomicAdd(transposed ? &H[c * 12 + r] : &H[r * 12 + c], s);
      }
  };
  auto add_b = [&](int hh, const double* A) {
    for (int r = 0; r < 7; ++r) {
      double s = 0.0;
      for (int d = 0; d < 7; ++d) s = fma(A[d * 7 + r], -e[d], s);
      atomicAdd(&bp[(size_t)hh * 12 + r], s);
    }
  };
  if (hi >= 0) { add_block(G.blk_ii[k], Ji, Ji, false); add_b(hi, Ji); }
  if (hj >= 0) { add_block(G.blk_jj[k], Jj, Jj, false); add_b(hj, Jj); }
  if (hi >= 0 && hj >= 0) {
    const int b = G.blk_ij[k];
    add_block(b & 0x3fffffff, Ji, Jj, (b & 0x40000000) != 0);   // upper storage: (min, max); transposed when h_i > h_j
  }
}

// Hpp + lambda I on the seven real dimensions, unit diagonal on the padding -> the reduced-system arrays the tile Cholesky
// loads (BlockSolver::setLambda, block_solver.hpp:563-589; without landmarks Hschur is Hpp, :353-365)
__global__ void k_pg_damp(int n_blk, const int* __restrict__ diag_of_blk, const double* __restrict__ hpp, double lambda,
                          double* __restrict__ hs, int n_pose, const double* __restrict__ bp, double* __restrict__ bs) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n_blk * 144) {
    const int blk = t / 144, r = (t % 144) / 12, c = t % 12;
    double v = hpp[t];
    if (diag_of_blk[blk] >= 0 && r == c) v += r < 7 ? lambda : 1.0;
    hs[t] = v;
  }
  if (t < n_pose * 12) bs[t] = bp[t];
}

// SparseOptimizer::update + computeScale: S_new = Sim3(x) S for the free vertices, scale partial = sum x (lambda x + b)
__global__ void k_pg_update(PgView G, double lambda, const double* __restrict__ x, const double* __restrict__ bp,
                            const double* __restrict__ S_cur, double* __restrict__ S_new, double* __restrict__ pose_scale) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= G.n_kf) return;
  const int h = G.h[i];
  Sim3 S = load_sim3(S_cur + 8 * i);
  if (h >= 0) {
    const double* xv = x + (size_t)h * 12;
    S = sim3_oplus(S, xv, G.fix_scale != 0);
    double sc = 0.0;
    for (int d = 0; d < 7; ++d) sc += xv[d] * (lambda * xv[d] + bp[(size_t)h * 12 + d]);
    pose_scale[h] = sc;
  }
  store_sim3(S, S_new + 8 * i);
}

// Map point correction after the pose graph (src/Optimizer.cc:1687-1712): P <- S_wr' (S_rw P)
__global__ void k_correct_points(int64_t n_pt, const double* __restrict__ xyz, const int* __restrict__ ref_kf,
                                 const double* __restrict__ before, const double* __restrict__ after, double* __restrict__ out) {
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n_pt; p += (int64_t)gridDim.x * blockDim.x) {
    const int r = ref_kf[p];
    const Sim3 Srw = load_sim3(before + 8 * (size_t)r), Swr = sim3_inv(load_sim3(after + 8 * (size_t)r));
    const V3 q = sim3_map(Swr, sim3_map(Srw, v3(xyz[3 * p], xyz[3 * p + 1], xyz[3 * p + 2])));
    out[3 * p] = q[0]; out[3 * p + 1] = q[1]; out[3 * p + 2] = q[2];
  }
}

}  // namespace gpba
