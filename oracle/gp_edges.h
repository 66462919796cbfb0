// oracle/gp_edges.h -- TEST INFRASTRUCTURE ONLY (CPU restatement; never linked into libgpba.so).
// PARITY: PINNED against the reference's own src/GaussianProcess.cc, src/G2oTypes.cc and g2o robust_kernel_impl.cpp, compiled
// unmodified into oracle/_ref (oracle/Makefile target _ref, stand-in headers oracle/ref_shim/): QueryPose and its 6x12 blocks,
// error and every Jacobian block of every edge below, the Huber kernel -- tests/test_ref_pin.py at 1e-11 (Huber: equal).
//
// GP interpolation + every edge type of the BA path, restated per edge exactly as the reference
// evaluates them (including the redundant per-observation QueryPose and the 12x12 products):
//   include/GaussianProcess.h:20-48   Qi, QiInv, Transition
//   src/GaussianProcess.cc:5-42       QueryPose (both overloads)
//   src/CameraModels/Pinhole.cpp:35-41 project, :71-81 projectJac
//   src/G2oTypes.cc:225-239           EdgeMonoGP::computeError
//   src/G2oTypes.cc:316-367           EdgeMonoGP::linearizeOplus   (== EdgeMonoGPExtrinsic :258-314 with a fixed extrinsic)
//   src/G2oTypes.cc:369-443           EdgeStereoGP
//   src/G2oTypes.cc:445-495, include/G2oTypes.h:423-468   EdgeMono / EdgeStereo
//   src/G2oTypes.cc:100-118, include/G2oTypes.h:155-163   EdgeGaussianPrior
//   include/G2oTypes.h:496-519        EdgeVelocity
//   src/G2oTypes.cc:65-81, include/G2oTypes.h:362-370     isDepthPositive
#pragma once
#include "lie.h"

namespace ora {

typedef Mat<12, 12> M12;
typedef Mat<12, 1> V12;
typedef Mat<6, 12> M6x12;

struct GaussianProcess {
  M6 Qc, QcInv;
  void set_diag(const double* qc) {
    Qc = M6::Zero();
    for (int i = 0; i < 6; ++i) Qc(i, i) = qc[i];
    QcInv = inverse<6>(Qc);  // GaussianProcess.h:15  mQcInv(Qc.inverse())
  }
  M12 Qi(double dt) const {  // GaussianProcess.h:20-29
    M12 m = M12::Zero();
    double dt2 = dt * dt, dt3 = dt2 * dt;
    m.set_block(0, 0, (1.0 / 3.0 * dt3) * Qc);
    m.set_block(0, 6, (1.0 / 2.0 * dt2) * Qc);
    m.set_block(6, 0, (1.0 / 2.0 * dt2) * Qc);
    m.set_block(6, 6, dt * Qc);
    return m;
  }
  M12 QiInv(double dt) const {  // :31-41
    M12 m = M12::Zero();
    double dt2 = dt * dt, dt3 = dt2 * dt;
    m.set_block(0, 0, (12.0 / dt3) * QcInv);
    m.set_block(0, 6, (-6.0 / dt2) * QcInv);
    m.set_block(6, 0, (-6.0 / dt2) * QcInv);
    m.set_block(6, 6, (4.0 / dt) * QcInv);
    return m;
  }
  M12 Transition(double t1, double t2) const {  // :44-48
    M12 m = M12::Identity();
    for (int i = 0; i < 6; ++i) m(i, 6 + i) = (t2 - t1);
    return m;
  }
  // GaussianProcess.cc:23-42 (the :5-21 overload is the same arithmetic without the extra outputs)
  SE3 QueryPose(const SE3& pose1, const SE3& pose2, const V6& v1, const V6& v2, double t1, double t2, double t,
                M6x12* At1o = nullptr, M6x12* Pt1o = nullptr, SE3* dTo = nullptr, V6* xi12o = nullptr) const {
    M12 Pt = Qi(t - t1) * transpose(Transition(t, t2)) * QiInv(t2 - t1);
    M12 At = Transition(t1, t) - Pt * Transition(t1, t2);
    M6x12 At1 = At.block<6, 12>(0, 0);
    M6x12 Pt1 = Pt.block<6, 12>(0, 0);
    V12 x1 = V12::Zero(), x2;
    for (int i = 0; i < 6; ++i) x1[6 + i] = v1[i];
    SE3 dp = se3_mul(se3_inv(pose1), pose2);
    V6 xi = se3_log(dp);
    V6 jv = RightJacobianPose3Inv(xi) * v2;
    for (int i = 0; i < 6; ++i) { x2[i] = xi[i]; x2[6 + i] = jv[i]; }
    V6 arg = At1 * x1 + Pt1 * x2;
    SE3 dT = se3_exp(arg);
    if (At1o) *At1o = At1;
    if (Pt1o) *Pt1o = Pt1;
    if (dTo) *dTo = dT;
    if (xi12o) *xi12o = xi;
    return se3_mul(pose1, dT);
  }
};

struct Pinhole {
  double fx, fy, cx, cy;
  void project(const V3& p, double* uv) const {  // Pinhole.cpp:35-41
    uv[0] = fx * p[0] / p[2] + cx;
    uv[1] = fy * p[1] / p[2] + cy;
  }
  Mat<2, 3> projectJac(const V3& p) const {  // Pinhole.cpp:71-81
    Mat<2, 3> J;
    J(0, 0) = fx / p[2]; J(0, 1) = 0.0; J(0, 2) = -fx * p[0] / (p[2] * p[2]);
    J(1, 0) = 0.0; J(1, 1) = fy / p[2]; J(1, 2) = -fy * p[1] / (p[2] * p[2]);
    return J;
  }
};

struct KfState { SE3 Twb; V6 vel; double time; };

// Reprojection error of one edge. gp==true: EdgeMonoGP / EdgeStereoGP; else EdgeMono / EdgeStereo.
// dim = 2 (mono) or 3 (stereo).  err[] = obs - projection.
inline void reproj_error(const GaussianProcess& G, bool gp, int dim, const KfState* f1, const KfState& f2, double t,
                         const SE3& Tbc, const Pinhole& cam, double bf, const V3& Xw, const double* obs, double* err) {
  SE3 Twb = gp ? G.QueryPose(f1->Twb, f2.Twb, f1->vel, f2.vel, f1->time, f2.time, t) : f2.Twb;
  SE3 Twc = se3_mul(Twb, Tbc);
  V3 Xc = se3_act(se3_inv(Twc), Xw);
  double uv[2];
  cam.project(Xc, uv);
  err[0] = obs[0] - uv[0];
  err[1] = obs[1] - uv[1];
  if (dim == 3) {
    double invZ = 1 / Xc[2];
    err[2] = obs[2] - (uv[0] - bf * invZ);
  }
}

// Jacobians of one reprojection edge; rows = dim (2|3). J1kf (dim x 12) wrt vertex 0 (KF_prev, GP only),
// J2kf (dim x 12) wrt KF_cur, Jpt (dim x 3).  Row-major, leading dimension 12 / 3.
inline void reproj_jacobian(const GaussianProcess& G, bool gp, int dim, const KfState* f1, const KfState& f2, double t,
                            const SE3& Tbc, const Pinhole& cam, double bf, const V3& Xw, double* J1kf, double* J2kf,
                            double* Jpt, double* Jext = nullptr /* dim x 6: EdgeMonoGPExtrinsic's fourth block */) {
  M6x12 At1, Pt1;
  SE3 dT, Twb;
  V6 xi12;
  if (gp)
    Twb = G.QueryPose(f1->Twb, f2.Twb, f1->vel, f2.vel, f1->time, f2.time, t, &At1, &Pt1, &dT, &xi12);
  else
    Twb = f2.Twb;
  const SE3 Tcb = se3_inv(Tbc);
  const M3 Rbw = transpose(se3_R(Twb));
  const V3 Xb = se3_act(se3_inv(Twb), Xw);
  const V3 Xc = se3_act(Tcb, Xb);
  Mat<3, 3> P = Mat<3, 3>::Zero();
  Mat<2, 3> p2 = cam.projectJac(Xc);
  for (int c = 0; c < 3; ++c) { P(0, c) = p2(0, c); P(1, c) = p2(1, c); }
  if (dim == 3) {  // G2oTypes.cc:407-412
    const double inv_z2 = 1.0 / (Xc[2] * Xc[2]);
    for (int c = 0; c < 3; ++c) P(2, c) = P(0, c);
    P(2, 2) += bf * inv_z2;
  }
  const M3 Rcb = se3_R(Tcb);
  Mat<3, 6> SE3deriv;
  SE3deriv.set_block(0, 0, -Rcb);
  SE3deriv.set_block(0, 3, Rcb * hat(Xb));
  Mat<3, 6> J1 = -(P * SE3deriv);  // rows >= dim unused
  Mat<3, 3> Jp = -(P * Rcb * Rbw);
  for (int r = 0; r < dim; ++r)
    for (int c = 0; c < 3; ++c) Jpt[r * 3 + c] = Jp(r, c);
  if (Jext) {  // _jacobianOplus[3] = -proj_jac * [-I, Skew(Xc)]   (src/G2oTypes.cc:311-313)
    Mat<3, 6> SE3deriv2;
    SE3deriv2.set_block(0, 0, -M3::Identity());
    SE3deriv2.set_block(0, 3, hat(Xc));
    Mat<3, 6> Je = -(P * SE3deriv2);
    for (int r = 0; r < dim; ++r)
      for (int c = 0; c < 6; ++c) Jext[r * 6 + c] = Je(r, c);
  }
  if (!gp) {  // EdgeMono / EdgeStereo: pose block = J1, velocity block = 0
    for (int r = 0; r < dim; ++r)
      for (int c = 0; c < 12; ++c) J2kf[r * 12 + c] = c < 6 ? J1(r, c) : 0.0;
    return;
  }
  V6 dxi = se3_log(dT);
  M6 Ad_dT = se3_Adj(se3_exp(-dxi));
  M6 Jr_dxi = RightJacobianPose3(dxi);
  M6 Jr_inv_xi12 = RightJacobianPose3Inv(xi12);
  M6 ad_v2 = se3Adj(f2.vel);
  M6 ad_T12 = se3_Adj(se3_exp(xi12));
  Mat<12, 6> JinT1 = Mat<12, 6>::Zero(), JinV1 = Mat<12, 6>::Zero(), JinT2 = Mat<12, 6>::Zero(),
             JinV2 = Mat<12, 6>::Zero();
  M6 top = -(Jr_inv_xi12 * inverse<6>(ad_T12));
  JinT1.set_block(0, 0, top);
  JinT1.set_block(6, 0, (-0.5 * ad_v2) * top);
  JinV1.set_block(6, 0, M6::Identity());
  JinT2.set_block(0, 0, Jr_inv_xi12);
  JinT2.set_block(6, 0, (-0.5 * ad_v2) * Jr_inv_xi12);
  JinV2.set_block(6, 0, Jr_inv_xi12);
  Mat<3, 6> JT1 = J1 * (Jr_dxi * Pt1 * JinT1 + Ad_dT);
  Mat<3, 6> JV1 = J1 * Jr_dxi * At1 * JinV1;
  Mat<3, 12> Jj1 = J1 * Jr_dxi * Pt1;
  Mat<3, 6> JT2 = Jj1 * JinT2;
  Mat<3, 6> JV2 = Jj1 * JinV2;
  for (int r = 0; r < dim; ++r)
    for (int c = 0; c < 6; ++c) {
      J1kf[r * 12 + c] = JT1(r, c);
      J1kf[r * 12 + 6 + c] = JV1(r, c);
      J2kf[r * 12 + c] = JT2(r, c);
      J2kf[r * 12 + 6 + c] = JV2(r, c);
    }
}

// PoseVelocity::isDepthPositive at one keyframe pose (G2oTypes.cc:65-81)
inline bool depth_positive(const SE3& Twb, const SE3& Tbc, const V3& Xw) {
  V3 Xc = se3_act(se3_inv(se3_mul(Twb, Tbc)), Xw);
  return Xc[2] > 0;
}

// EdgeGaussianPrior::computeError (G2oTypes.h:155-163)
inline void prior_error(const KfState& f1, const KfState& f2, double* err) {
  V6 dxi = se3_log(se3_mul(se3_inv(f1.Twb), f2.Twb));
  V6 a = dxi - (f2.time - f1.time) * f1.vel;
  V6 b = RightJacobianPose3Inv(dxi) * f2.vel - f1.vel;
  for (int i = 0; i < 6; ++i) { err[i] = a[i]; err[6 + i] = b[i]; }
}
// EdgeGaussianPrior::linearizeOplus (G2oTypes.cc:100-118)
inline void prior_jacobian(const KfState& f1, const KfState& f2, M12* Ji, M12* Jj) {
  SE3 T = se3_mul(se3_inv(f1.Twb), f2.Twb);
  V6 xi = se3_log(T);
  M6 Jr_inv_T = RightJacobianPose3Inv(xi);
  M6 ad_v2 = se3Adj(f2.vel);
  *Ji = M12::Zero();
  *Jj = M12::Zero();
  M6 a = -(Jr_inv_T * inverse<6>(se3_Adj(T)));
  Ji->set_block(0, 0, a);
  Ji->set_block(6, 0, (-0.5 * ad_v2) * a);
  Ji->set_block(0, 6, -(f2.time - f1.time) * M6::Identity());
  Ji->set_block(6, 6, -M6::Identity());
  Jj->set_block(0, 0, Jr_inv_T);
  Jj->set_block(6, 0, (-0.5 * ad_v2) * Jr_inv_T);
  Jj->set_block(6, 6, Jr_inv_T);
}

// RobustKernelHuber with the reference's float-typed dsqr (robust_kernel_impl.h:84, .cpp:65-91)
// EdgeExtrinsicPrior (include/G2oTypes.h:470-494): e = Log(R_ini^-1 R_bc), J = [0 | RightJacobianSO3(e)^-1] wrt the
// 6-dim tangent [translation; rotation] of VertexExtrinsic (oplus: Tbc <- Tbc exp(delta), :98-100).
// RightJacobianSO3 (src/G2oTypes.cc:575-590, ORB-SLAM3's): I - W (1 - cos d) / d^2 + W^2 (d - sin d) / d^3, identity below d = 1e-5;
// `.inverse()` is Eigen's fixed-size 3 x 3 inverse (cofactors).
inline M3 RightJacobianSO3_orb(const V3& v) {
  const double d2 = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
  const double d = std::sqrt(d2);
  M3 W = hat(v);
  if (d < 1e-5) return M3::Identity();
  return M3::Identity() - W * ((1.0 - std::cos(d)) / d2) + (W * W) * ((d - std::sin(d)) / (d2 * d));
}
inline M3 inverse3_cofactor(const M3& A) {
  M3 C;
  C(0, 0) = A(1, 1) * A(2, 2) - A(1, 2) * A(2, 1); C(0, 1) = A(0, 2) * A(2, 1) - A(0, 1) * A(2, 2); C(0, 2) = A(0, 1) * A(1, 2) - A(0, 2) * A(1, 1);
  C(1, 0) = A(1, 2) * A(2, 0) - A(1, 0) * A(2, 2); C(1, 1) = A(0, 0) * A(2, 2) - A(0, 2) * A(2, 0); C(1, 2) = A(0, 2) * A(1, 0) - A(0, 0) * A(1, 2);
  C(2, 0) = A(1, 0) * A(2, 1) - A(1, 1) * A(2, 0); C(2, 1) = A(0, 1) * A(2, 0) - A(0, 0) * A(2, 1); C(2, 2) = A(0, 0) * A(1, 1) - A(0, 1) * A(1, 0);
  const double det = A(0, 0) * C(0, 0) + A(0, 1) * C(1, 0) + A(0, 2) * C(2, 0);
  return C * (1.0 / det);
}
inline V3 ext_prior_error(const Quat& q_ini_inv, const SE3& Tbc) { double theta; return so3_log(quat_mul(q_ini_inv, Tbc.q), &theta); }

struct Huber {
  double delta;
  float dsqr;
  void setDelta(double d) { dsqr = (float)(d * d); delta = d; }
  void robustify(double e, double* rho) const {
    if (e <= dsqr) {
      rho[0] = e; rho[1] = 1.; rho[2] = 0.;
    } else {
      double sqrte = std::sqrt(e);
      rho[0] = 2 * sqrte * delta - dsqr;
      rho[1] = delta / sqrte;
      rho[2] = -0.5 * rho[1] / e;
    }
  }
};

}  // namespace ora
