// TEST INFRASTRUCTURE, not product code.  C entry points over the REFERENCE'S OWN edge code: this file is linked with
// /root/reference/src/{Pose3utils,GaussianProcess,G2oTypes}.cc, compiled unmodified from where they lie against the stand-in
// headers under oracle/ref_shim/ (Eigen, Sophus, the g2o base classes and the map classes are absent from this image or need
// them; see the headers there for what each stand-in does and does not pin).  Built by `make -C oracle _ref` into
// oracle/_ref/libamc_ref_edges.so, used by tests/golden/make_golden_ref.py to mint tests/golden/ref_edges.npz and by
// tests/test_ref_pin.py to check the oracle restatement (oracle/gp_edges.h, lie.h, pose_only.h, vel_ransac.h) against what the
// reference computes: every function below calls the reference's computeError() / linearizeOplus() / QueryPose / Jacobian
// helpers and copies the result out.  Same probe signatures as the oracle's (oracle/gpba_oracle.cc, oracle_* probes):
// poses as [qx qy qz qw tx ty tz], matrices row-major, tangent order [translation, rotation].
//
// What is the reference's own arithmetic here: the GP interpolation and its 6x12 blocks (GaussianProcess.cc/.h), the SE(3)
// left/right Jacobians and their inverses, Q, se3Adj, CircleDot (Pose3utils.cc), every edge's error and Jacobian blocks, the
// PoseVelocity update and the SO(3) helpers (G2oTypes.cc/.h).  What is stand-in arithmetic: matrix products / inverses
// (Eigen), quaternion algebra and SE(3) exp/log/Adj (Sophus), the pinhole projection (Pinhole.cpp:35-41, 71-81 restated below).
#include <cstring>
#include "G2oTypes.h"
#include "Thirdparty/g2o/g2o/core/robust_kernel_impl.h"
#include "Thirdparty/g2o/g2o/types/sim3.h"

using namespace ORB_SLAM3;
typedef Eigen::Matrix<double, 6, 1> V6;

std::vector<Sophus::SE3d> ORB_SLAM3::MultiKeyFrame::mTbc;
std::vector<Sophus::SE3d> ORB_SLAM3::MultiFrame::mTbc;

namespace {

struct PinholeStandIn : GeometricCamera {   // src/CameraModels/Pinhole.cpp:35-41, 71-81
  double fx, fy, cx, cy;
  explicit PinholeStandIn(const double* k) : fx(k[0]), fy(k[1]), cx(k[2]), cy(k[3]) {}
  Eigen::Vector2d project(const Eigen::Vector3d& v) override {
    Eigen::Vector2d r;
    r[0] = fx * v[0] / v[2] + cx;
    r[1] = fy * v[1] / v[2] + cy;
    return r;
  }
  Eigen::Matrix<double, 2, 3> projectJac(const Eigen::Vector3d& v) override {
    Eigen::Matrix<double, 2, 3> J;
    J(0, 0) = fx / v[2]; J(0, 1) = 0; J(0, 2) = -fx * v[0] / (v[2] * v[2]);
    J(1, 0) = 0; J(1, 1) = fy / v[2]; J(1, 2) = -fy * v[1] / (v[2] * v[2]);
    return J;
  }
};

Sophus::SE3d from7(const double* p) {
  return Sophus::SE3d(Sophus::SO3d::fromQuaternion(p[0], p[1], p[2], p[3]), Eigen::Vector3d(p[4], p[5], p[6]));
}
void to7(const Sophus::SE3d& T, double* p) {
  p[0] = T.so3().qx(); p[1] = T.so3().qy(); p[2] = T.so3().qz(); p[3] = T.so3().qw();
  for (int i = 0; i < 3; ++i) p[4 + i] = T.translation()(i);
}
V6 v6(const double* p) { V6 v; for (int i = 0; i < 6; ++i) v(i) = p[i]; return v; }
template <class M> void out(const M& m, double* o) {
  if (!o) return;
  for (int i = 0; i < m.rows(); ++i) for (int j = 0; j < m.cols(); ++j) o[i * m.cols() + j] = m(i, j);
}
GaussianProcess make_gp(const double* qc) {
  Eigen::Matrix<double, 6, 6> Qc = Eigen::Matrix<double, 6, 6>::Zero();
  for (int i = 0; i < 6; ++i) Qc(i, i) = qc[i];
  return GaussianProcess(Qc);
}
PoseVelocity make_pv(const double* T7, const double* v, double time, double bf, GeometricCamera* cam) {
  PoseVelocity pv;
  pv.Twb = from7(T7); pv.Vel = v6(v); pv.time = time; pv.bf = bf;
  pv.vpCameras.assign(1, cam);
  return pv;
}
void set_tbc(const double* Tbc7) {
  MultiKeyFrame::mTbc.assign(1, from7(Tbc7));   // one camera: mTbc[0] == mTbc.back()
  MultiFrame::mTbc = MultiKeyFrame::mTbc;
}

}  // namespace

extern "C" {

// Pose3utils.cc: 0 LeftJacobianPose3, 1 RightJacobianPose3, 2 LeftJacobianPose3Inv, 3 RightJacobianPose3Inv, 4 se3Adj
void ref_jac_pose3(const double* xi, int which, double* out36) {
  const V6 x = v6(xi);
  Eigen::Matrix<double, 6, 6> J;
  switch (which) {
    case 0: J = LeftJacobianPose3(x); break;
    case 1: J = RightJacobianPose3(x); break;
    case 2: J = LeftJacobianPose3Inv(x); break;
    case 3: J = RightJacobianPose3Inv(x); break;
    default: J = se3Adj(x); break;
  }
  out(J, out36);
}
// Pose3utils.cc: 0 LeftJacobianPose3Q(xi) (3x3), 1 LeftJacobianRot3(xi[3:6]), 2 LeftJacobianRot3Inv(xi[3:6])
void ref_jac_small(const double* xi, int which, double* out9) {
  const V6 x = v6(xi);
  const Eigen::Vector3d om = x.tail<3>();
  Eigen::Matrix3d J;
  if (which == 0) J = LeftJacobianPose3Q(x);
  else if (which == 1) J = LeftJacobianRot3(om);
  else J = LeftJacobianRot3Inv(om);
  out(J, out9);
}
void ref_circle_dot(const double* p3, double* out24) { out(CircleDot(Eigen::Vector3d(p3[0], p3[1], p3[2])), out24); }
// G2oTypes.cc SO(3) helpers: 0 RightJacobianSO3, 1 InverseRightJacobianSO3, 2 ExpSO3, 3 Skew; LogSO3 below
void ref_so3_helper(const double* w3, int which, double* out9) {
  const Eigen::Vector3d w(w3[0], w3[1], w3[2]);
  Eigen::Matrix3d J;
  if (which == 0) J = RightJacobianSO3(w);
  else if (which == 1) J = InverseRightJacobianSO3(w);
  else if (which == 2) J = ExpSO3(w);
  else J = Skew(w);
  out(J, out9);
}
void ref_log_so3(const double* R9, double* w3) {
  Eigen::Matrix3d R;
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) R(i, j) = R9[3 * i + j];
  out(LogSO3(R), w3);
}
// GaussianProcess::QueryPose (GaussianProcess.cc:24-42) with the 6x12 blocks At1, Pt1; the plain overload (:5-22) is checked
// against it here: the return value is 1 when both overloads agree to the last bit
int ref_query_pose(const double* qc, const double* T1, const double* T2, const double* v1, const double* v2, double t1, double t2,
                   double t, double* out7, double* At1, double* Pt1, double* dT7, double* xi12) {
  GaussianProcess gp = make_gp(qc);
  Eigen::Matrix<double, 6, 12> A, P;
  Sophus::SE3d dT;
  Eigen::VectorXd xi;
  const Sophus::SE3d T = gp.QueryPose(from7(T1), from7(T2), v6(v1), v6(v2), t1, t2, t, A, P, dT, xi);
  const Sophus::SE3d Tplain = gp.QueryPose(from7(T1), from7(T2), v6(v1), v6(v2), t1, t2, t);
  to7(T, out7);
  out(A, At1); out(P, Pt1);
  if (dT7) to7(dT, dT7);
  out(xi, xi12);
  double a[7], b[7];
  to7(T, a); to7(Tplain, b);
  return std::memcmp(a, b, sizeof(a)) == 0 ? 1 : 0;
}
// GaussianProcess.h:24-54: Qi(dt), QiInv(dt), Transition(t1, t2) as 12x12
void ref_gp_matrices(const double* qc, double dt, double t1, double t2, double* Qi144, double* QiInv144, double* Phi144) {
  GaussianProcess gp = make_gp(qc);
  out(gp.Qi(dt), Qi144); out(gp.QiInv(dt), QiInv144); out(gp.Transition(t1, t2), Phi144);
}

// One landmark edge: EdgeMonoGP (gp, dim 2), EdgeStereoGP (gp, dim 3), EdgeMono (no gp, dim 2), EdgeStereo (no gp, dim 3);
// dim = 3 when obs3[2] >= 0.  error (dim), J1 / J2 (dim x 12, J1 only for GP edges), Jp (dim x 3); depth = isDepthPositive()
// where the edge has one (EdgeMonoGP, EdgeMono), else -1.
int ref_edge_eval(const double* qc, int gp_edge, const double* T1, const double* v1, double t1, const double* T2, const double* v2,
                  double t2, double t, const double* Tbc7, const double* intr, double bf, const double* Xw, const double* obs3,
                  double* err, double* J1, double* J2, double* Jp) {
  GaussianProcess gp = make_gp(qc);
  PinholeStandIn cam(intr);
  set_tbc(Tbc7);
  VertexPoseVel va, vb;
  g2o::VertexSBAPointXYZ vp;
  vp.setEstimate(Eigen::Vector3d(Xw[0], Xw[1], Xw[2]));
  vb.setEstimate(make_pv(T2, v2, t2, bf, &cam));
  const int dim = obs3[2] >= 0 ? 3 : 2;
  int depth = -1;
  if (gp_edge) {
    va.setEstimate(make_pv(T1, v1, t1, bf, &cam));
    if (dim == 2) {
      EdgeMonoGP e(0, t, &gp);
      e.setVertex(0, &va); e.setVertex(1, &vb); e.setVertex(2, &vp);
      e.setMeasurement(Eigen::Vector2d(obs3[0], obs3[1]));
      e.computeError(); e.linearizeOplus();
      out(e.error(), err); out(e.jacobianOplus()[0], J1); out(e.jacobianOplus()[1], J2); out(e.jacobianOplus()[2], Jp);
      depth = e.isDepthPositive() ? 1 : 0;
    } else {
      EdgeStereoGP e(0, t, &gp);
      e.setVertex(0, &va); e.setVertex(1, &vb); e.setVertex(2, &vp);
      e.setMeasurement(Eigen::Vector3d(obs3[0], obs3[1], obs3[2]));
      e.computeError(); e.linearizeOplus();
      out(e.error(), err); out(e.jacobianOplus()[0], J1); out(e.jacobianOplus()[1], J2); out(e.jacobianOplus()[2], Jp);
    }
  } else if (dim == 2) {
    EdgeMono e;
    e.setVertex(0, &vb); e.setVertex(1, &vp);
    e.setMeasurement(Eigen::Vector2d(obs3[0], obs3[1]));
    e.computeError(); e.linearizeOplus();
    out(e.error(), err); out(e.jacobianOplusXi(), J2); out(e.jacobianOplusXj(), Jp);
    depth = e.isDepthPositive() ? 1 : 0;
  } else {
    EdgeStereo e;
    e.setVertex(0, &vb); e.setVertex(1, &vp);
    e.setMeasurement(Eigen::Vector3d(obs3[0], obs3[1], obs3[2]));
    e.computeError(); e.linearizeOplus();
    out(e.error(), err); out(e.jacobianOplusXi(), J2); out(e.jacobianOplusXj(), Jp);
  }
  return depth;
}
// EdgeMonoGPExtrinsic (G2oTypes.cc:239-314): the extrinsic is the fourth vertex, not MultiKeyFrame::mTbc
int ref_edge_ext_eval(const double* qc, const double* T1, const double* v1, double t1, const double* T2, const double* v2, double t2,
                      double t, const double* Tbc7, const double* intr, double bf, const double* Xw, const double* obs2,
                      double* err, double* J1, double* J2, double* Jp, double* Jext) {
  GaussianProcess gp = make_gp(qc);
  PinholeStandIn cam(intr);
  MultiKeyFrame::mTbc.clear();   // any use of the static extrinsics on this path would fault
  VertexPoseVel va, vb;
  g2o::VertexSBAPointXYZ vp;
  VertexExtrinsic ve(from7(Tbc7));
  vp.setEstimate(Eigen::Vector3d(Xw[0], Xw[1], Xw[2]));
  va.setEstimate(make_pv(T1, v1, t1, bf, &cam));
  vb.setEstimate(make_pv(T2, v2, t2, bf, &cam));
  EdgeMonoGPExtrinsic e(0, t, &gp);
  e.setVertex(0, &va); e.setVertex(1, &vb); e.setVertex(2, &vp); e.setVertex(3, &ve);
  e.setMeasurement(Eigen::Vector2d(obs2[0], obs2[1]));
  e.computeError(); e.linearizeOplus();
  out(e.error(), err);
  out(e.jacobianOplus()[0], J1); out(e.jacobianOplus()[1], J2); out(e.jacobianOplus()[2], Jp); out(e.jacobianOplus()[3], Jext);
  return e.isDepthPositive() ? 1 : 0;
}
// The tracking-side edges on a fixed landmark: EdgeMonoGPOnlyPose (gp), EdgeMonoOnlyPose (no gp, dim 2), EdgeStereoOnlyPose
// (no gp, dim 3); Xw is a float vector in their constructors (G2oTypes.h:190,223,250) -- pass float-representable values
int ref_pose_edge_eval(const double* qc, int gp_edge, const double* T1, const double* v1, double t1, const double* T2,
                       const double* v2, double t2, double t, const double* Tbc7, const double* intr, double bf, const double* Xw,
                       const double* obs3, double* err, double* J1, double* J2) {
  GaussianProcess gp = make_gp(qc);
  PinholeStandIn cam(intr);
  set_tbc(Tbc7);
  VertexPoseVel va, vb;
  vb.setEstimate(make_pv(T2, v2, t2, bf, &cam));
  const Eigen::Vector3f Xf((float)Xw[0], (float)Xw[1], (float)Xw[2]);
  const int dim = obs3[2] >= 0 ? 3 : 2;
  if (gp_edge) {
    va.setEstimate(make_pv(T1, v1, t1, bf, &cam));
    EdgeMonoGPOnlyPose e(Xf, 0, t, &gp);
    e.setVertex(0, &va); e.setVertex(1, &vb);
    e.setMeasurement(Eigen::Vector2d(obs3[0], obs3[1]));
    e.computeError(); e.linearizeOplus();
    out(e.error(), err); out(e.jacobianOplusXi(), J1); out(e.jacobianOplusXj(), J2);
    return e.isDepthPositive() ? 1 : 0;
  }
  if (dim == 2) {
    EdgeMonoOnlyPose e(Xf);
    e.setVertex(0, &vb);
    e.setMeasurement(Eigen::Vector2d(obs3[0], obs3[1]));
    e.computeError(); e.linearizeOplus();
    out(e.error(), err); out(e.jacobianOplusXi(), J2);
    return e.isDepthPositive() ? 1 : 0;
  }
  EdgeStereoOnlyPose e(Xf);
  e.setVertex(0, &vb);
  e.setMeasurement(Eigen::Vector3d(obs3[0], obs3[1], obs3[2]));
  e.computeError(); e.linearizeOplus();
  out(e.error(), err); out(e.jacobianOplusXi(), J2);
  return -1;
}
// EdgeGaussianPrior (G2oTypes.h:147-184, G2oTypes.cc:96-115): error (12), Jacobians (12 x 12 each)
void ref_prior_eval(const double* T1, const double* v1, double t1, const double* T2, const double* v2, double t2, double* err12,
                    double* Ji144, double* Jj144) {
  VertexPoseVel va, vb;
  va.setEstimate(make_pv(T1, v1, t1, 0, nullptr));
  vb.setEstimate(make_pv(T2, v2, t2, 0, nullptr));
  EdgeGaussianPrior e;
  e.setVertex(0, &va); e.setVertex(1, &vb);
  e.computeError(); e.linearizeOplus();
  out(e.error(), err12); out(e.jacobianOplusXi(), Ji144); out(e.jacobianOplusXj(), Jj144);
}
// EdgeExtrinsicPrior (G2oTypes.h:470-494): error (3), full Jacobian (3 x 6)
void ref_ext_prior_eval(const double* q_ini, const double* Tbc7, double* err3, double* J18) {
  VertexExtrinsic ve(from7(Tbc7));
  EdgeExtrinsicPrior e(Sophus::SO3d::fromQuaternion(q_ini[0], q_ini[1], q_ini[2], q_ini[3]));
  e.setVertex(0, &ve);
  e.computeError(); e.linearizeOplus();
  out(e.error(), err3); out(e.jacobianOplusXi(), J18);
}
// EdgeVelocity (G2oTypes.h:496-519): error (1), Jacobian (1 x 12)
void ref_velocity_edge_eval(const double* T7, const double* v, double* err1, double* J12) {
  VertexPoseVel va;
  va.setEstimate(make_pv(T7, v, 0, 0, nullptr));
  EdgeVelocity e;
  e.setVertex(0, &va);
  e.computeError(); e.linearizeOplus();
  out(e.error(), err1); out(e.jacobianOplusXi(), J12);
}
// EdgeVelReproj (G2oTypes.h:521-547, G2oTypes.cc:497-510): error (2), Jacobian (2 x 6) wrt the body velocity
void ref_vel_edge_eval(const double* Tlast7, const double* Tbc7, const double* intr, double dt, const double* vel, const double* Xw,
                       const double* obs2, double* err2, double* J12) {
  PinholeStandIn cam(intr);
  set_tbc(Tbc7);
  MultiFrame F;
  F.mvpCamera.assign(1, &cam);
  VertexVel vv;
  vv.setEstimate(v6(vel));
  EdgeVelReproj e(from7(Tlast7), dt, Eigen::Vector3d(Xw[0], Xw[1], Xw[2]), 0, &F);
  e.setVertex(0, &vv);
  e.setMeasurement(Eigen::Vector2d(obs2[0], obs2[1]));
  e.computeError(); e.linearizeOplus();
  out(e.error(), err2); out(e.jacobianOplusXi(), J12);
}
// The vertex updates: VertexPoseVel::oplusImpl -> PoseVelocity::Update (G2oTypes.cc:41-46), VertexExtrinsic::oplusImpl
// (G2oTypes.h:99-101), VertexVel::oplusImpl (G2oTypes.h:141-143)
void ref_posevel_update(const double* T7, const double* v, const double* upd12, double* T7_out, double* v_out) {
  VertexPoseVel va;
  va.setEstimate(make_pv(T7, v, 0, 0, nullptr));
  va.oplus(upd12);
  to7(va.estimate().Twb, T7_out);
  out(va.estimate().Vel, v_out);
}
void ref_extrinsic_update(const double* Tbc7, const double* upd6, double* T7_out) {
  VertexExtrinsic ve(from7(Tbc7));
  ve.oplus(upd6);
  to7(ve.estimate(), T7_out);
}
// RobustKernelHuber (Thirdparty/g2o/g2o/core/robust_kernel_impl.cpp:65-91, the reference's own file): rho, rho', rho''
void ref_huber(double delta, double e, double* rho3) {
  g2o::RobustKernelHuber k;
  k.setDelta(delta);
  Eigen::Vector3d rho;
  k.robustify(e, rho);
  out(rho, rho3);
}
// g2o::Sim3 from the reference's own Thirdparty/g2o/g2o/types/sim3.h (exp = the Vector7d constructor :69-139, log :146-222,
// product :262-268, inverse :225-228), the arithmetic of VertexSim3Expmap / EdgeSim3 in the essential-graph optimisation.
// S8 = [qx qy qz qw tx ty tz s], tangent [omega, upsilon, sigma].
static g2o::Sim3 sim3_from8(const double* p) {
  return g2o::Sim3(Eigen::Quaterniond(p[3], p[0], p[1], p[2]), Eigen::Vector3d(p[4], p[5], p[6]), p[7]);
}
static void sim3_to8(const g2o::Sim3& S, double* p) { for (int i = 0; i < 8; ++i) p[i] = S[i]; }
static g2o::Vector7d v7(const double* p) { g2o::Vector7d v; for (int i = 0; i < 7; ++i) v[i] = p[i]; return v; }
void ref_sim3_exp(const double* u7, double* S8) { sim3_to8(g2o::Sim3(v7(u7)), S8); }
void ref_sim3_log(const double* S8, double* u7) { out(sim3_from8(S8).log(), u7); }
void ref_sim3_mul(const double* a, const double* b, double* o8) { sim3_to8(sim3_from8(a) * sim3_from8(b), o8); }
void ref_sim3_inv(const double* a, double* o8) { sim3_to8(sim3_from8(a).inverse(), o8); }
void ref_sim3_map(const double* a, const double* p3, double* o3) { out(sim3_from8(a).map(Eigen::Vector3d(p3[0], p3[1], p3[2])), o3); }
// EdgeSim3::computeError (types_seven_dof_expmap.h:111-119): log(C * S_i * S_j^-1), the composition restated on sim3.h
void ref_sim3_edge_error(const double* meas8, const double* Si8, const double* Sj8, double* err7) {
  const g2o::Sim3 C = sim3_from8(meas8);
  const g2o::Sim3 error_ = C * sim3_from8(Si8) * sim3_from8(Sj8).inverse();
  out(error_.log(), err7);
}
// VertexSim3Expmap::oplusImpl (types_seven_dof_expmap.h:60-69): S <- Sim3(update) * S, update[6] = 0 when the scale is fixed
void ref_sim3_update(const double* S8, const double* u7, int fix_scale, double* o8) {
  g2o::Vector7d update = v7(u7);
  if (fix_scale) update[6] = 0;
  sim3_to8(g2o::Sim3(update) * sim3_from8(S8), o8);
}
// stand-in arithmetic, exported so the tests can state how far the stand-in Lie layer is from the oracle's
void ref_standin_se3_exp(const double* xi, double* out7) { to7(Sophus::SE3d::exp(v6(xi)), out7); }
void ref_standin_se3_log(const double* T7, double* xi) { out(from7(T7).log(), xi); }

}  // extern "C"
