// TEST INFRASTRUCTURE, not product code.  Shared by oracle/ref_g2o_run.cc and oracle/ref_adapter_check.cc: the includes of the
// reference's real g2o / AMC-SLAM headers, the pinhole stand-in, pose conversion, the per-iteration recorder and BaGraph --
// the g2o graph of Optimizer::BundleAdjustment / LocalGPBA built from a flattened gpba_problem (see ref_g2o_run.cc).
#pragma once
#include <cstring>
#include <vector>
#include "../include/gpba.h"
#include "G2oTypes.h"
#include "Thirdparty/g2o/g2o/core/block_solver.h"
#include "Thirdparty/g2o/g2o/core/hyper_graph_action.h"
#include "Thirdparty/g2o/g2o/core/optimization_algorithm_levenberg.h"
#include "Thirdparty/g2o/g2o/core/robust_kernel_impl.h"
#include "Thirdparty/g2o/g2o/core/sparse_optimizer.h"
#include "Thirdparty/g2o/g2o/solvers/linear_solver_dense.h"
#include "Thirdparty/g2o/g2o/types/types_seven_dof_expmap.h"

using namespace ORB_SLAM3;
typedef Eigen::Matrix<double, 6, 1> V6;

#ifdef REF_G2O_DEFINE_STATICS   // exactly one translation unit per library defines the stand-in map classes' statics
std::vector<Sophus::SE3d> ORB_SLAM3::MultiKeyFrame::mTbc;
std::vector<Sophus::SE3d> ORB_SLAM3::MultiFrame::mTbc;
#endif

namespace {

struct PinholeStandIn : GeometricCamera {   // src/CameraModels/Pinhole.cpp:35-41, 71-81
  double fx, fy, cx, cy;
  explicit PinholeStandIn(const double* k) : fx(k[0]), fy(k[1]), cx(k[2]), cy(k[3]) {}
  float getParameter(const int i) override { const double p[4] = {fx, fy, cx, cy}; return (float)p[i]; }   // float intrinsics (GeometricCamera.h:101)
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

// what LM did in each outer iteration, read after the iteration (sparse_optimizer.cpp:413 postIteration)
struct Recorder : g2o::HyperGraphAction {
  g2o::SparseOptimizer* opt = nullptr;
  g2o::OptimizationAlgorithmLevenberg* alg = nullptr;
  gpba_lm_trace* tr = nullptr;
  g2o::HyperGraphAction* operator()(const g2o::HyperGraph*, Parameters* p) override {
    const int i = static_cast<ParametersIteration*>(p)->iteration;
    if (tr && i >= 0 && i < GPBA_MAX_ITERS) {
      tr->levenberg_iterations[i] = alg->levenbergIteration();
      tr->total_trials += alg->levenbergIteration();
      tr->lambda[i] = alg->currentLambda();
      // the chi2 of the STORED edge errors: the last trial's, accepted or not (the stale-error quirk)
      tr->chi2_after[i] = opt->activeRobustChi2();
      tr->last_trial_chi2 = tr->chi2_after[i];
    }
    return this;
  }
};

}  // namespace

namespace {

// The graph of Optimizer::BundleAdjustment / LocalGPBA (src/Optimizer.cc:66-330, 838-1210) over a flattened problem.
struct BaGraph {
  GaussianProcess gp;
  std::vector<PinholeStandIn> cams;
  std::vector<GeometricCamera*> cam_ptrs;
  g2o::SparseOptimizer optimizer;                      // declared after gp / cams, so destroyed before them: the edges point into both
  g2o::OptimizationAlgorithmLevenberg* solver = nullptr;   // owned by the optimizer
  std::vector<VertexPoseVel*> vkf;
  std::vector<g2o::VertexSBAPointXYZ*> vpt;
  std::vector<g2o::OptimizableGraph::Edge*> eobs;
  std::vector<int> kind;                               // per observation: 0 EdgeMonoGP, 1 EdgeStereoGP, 2 EdgeMono, 3 EdgeStereo

  static GaussianProcess make_gp(const gpba_problem* P) {
    Eigen::Matrix<double, 6, 6> Qc = Eigen::Matrix<double, 6, 6>::Zero();
    for (int i = 0; i < 6; ++i) Qc(i, i) = P->qc[i];
    return GaussianProcess(Qc);
  }
  BaGraph(const gpba_problem* P, int max_trials) : gp(make_gp(P)) {
    for (int c = 0; c < P->n_cam; ++c) cams.emplace_back(P->cam_intr + 4 * c);
    for (auto& c : cams) cam_ptrs.push_back(&c);
    MultiKeyFrame::mTbc.clear();
    for (int c = 0; c < P->n_cam; ++c) MultiKeyFrame::mTbc.push_back(from7(P->cam_Tbc + 7 * c));   // reference camera last
    MultiFrame::mTbc = MultiKeyFrame::mTbc;

    // ---- Optimizer.cc:66-77 / 838-856
    g2o::BlockSolverX::LinearSolverType* linearSolver = new g2o::LinearSolverDense<g2o::BlockSolverX::PoseMatrixType>();
    g2o::BlockSolverX* solver_ptr = new g2o::BlockSolverX(linearSolver);
    solver = new g2o::OptimizationAlgorithmLevenberg(solver_ptr);
    if (P->lambda_init > 0) solver->setUserLambdaInit(P->lambda_init);
    if (max_trials > 0) solver->setMaxTrialsAfterFailure(max_trials);
    optimizer.setAlgorithm(solver);
    optimizer.setVerbose(false);

    // ---- keyframe vertices (:82-95): id = index (ascending id = Hessian order)
    vkf.resize(P->n_kf);
    for (int k = 0; k < P->n_kf; ++k) {
      PoseVelocity pv;
      pv.Twb = from7(P->kf_pose + 7 * k);
      for (int i = 0; i < 6; ++i) pv.Vel(i) = P->kf_vel[6 * k + i];
      pv.time = P->kf_time[k]; pv.bf = P->bf; pv.vpCameras = cam_ptrs;
      VertexPoseVel* v = new VertexPoseVel();
      v->setEstimate(pv);
      v->setId(k);
      v->setFixed(P->kf_fixed[k] != 0);
      optimizer.addVertex(v);
      vkf[k] = v;
    }
    // ---- GP constraints (:98-135): EdgeVelocity with QcInv(2,2), EdgeGaussianPrior with QiInv(dt) (+ Huber 21.026 in global BA)
    for (int i = 0; i < P->n_velp; ++i) {
      EdgeVelocity* e = new EdgeVelocity();
      e->setVertex(0, vkf[P->velp_kf[i]]);
      e->setInformation(gp.mQcInv.block<1, 1>(2, 2));
      optimizer.addEdge(e);
    }
    for (int i = 0; i < P->n_prior; ++i) {
      EdgeGaussianPrior* e = new EdgeGaussianPrior();
      e->setVertex(0, vkf[P->prior_kf1[i]]);
      e->setVertex(1, vkf[P->prior_kf2[i]]);
      if (P->huber_prior > 0) {
        g2o::RobustKernelHuber* rk = new g2o::RobustKernelHuber;
        e->setRobustKernel(rk);
        rk->setDelta(P->huber_prior);
      }
      e->setInformation(gp.QiInv(P->kf_time[P->prior_kf2[i]] - P->kf_time[P->prior_kf1[i]]));
      optimizer.addEdge(e);
    }
    // ---- landmark vertices (:144-153), marginalized
    vpt.resize(P->n_pt);
    for (int p = 0; p < P->n_pt; ++p) {
      g2o::VertexSBAPointXYZ* vP = new g2o::VertexSBAPointXYZ();
      vP->setEstimate(Eigen::Vector3d(P->pt_xyz[3 * p], P->pt_xyz[3 * p + 1], P->pt_xyz[3 * p + 2]));
      vP->setId(P->n_kf + p);
      vP->setMarginalized(true);
      optimizer.addVertex(vP);
      vpt[p] = vP;
    }
    // ---- reprojection edges in insertion order (:168-330)
    eobs.resize((size_t)P->n_obs);
    kind.resize((size_t)P->n_obs);
    for (int64_t i = 0; i < P->n_obs; ++i) {
      const int r = P->obs_rec[i], kf1 = P->rec_kf1[r], kf2 = P->rec_kf2[r], cam = P->rec_cam[r];
      const double ur = P->obs_ur ? P->obs_ur[i] : -1.0, w = P->obs_inv_sigma2[i];
      const unsigned flags = P->obs_flags ? P->obs_flags[i] : 0u;
      const bool stereo = ur >= 0;
      g2o::OptimizableGraph::Edge* edge = nullptr;
      if (kf1 >= 0 && !stereo) {
        EdgeMonoGP* e = new EdgeMonoGP(cam, P->rec_t[r], &gp);
        e->setVertex(0, vkf[kf1]); e->setVertex(1, vkf[kf2]); e->setVertex(2, vpt[P->obs_pt[i]]);
        e->setMeasurement(Eigen::Vector2d(P->obs_u[i], P->obs_v[i]));
        e->setInformation(Eigen::Matrix2d::Identity() * w);
        edge = e; kind[i] = 0;
      } else if (kf1 >= 0) {
        EdgeStereoGP* e = new EdgeStereoGP(cam, P->rec_t[r], &gp);
        e->setVertex(0, vkf[kf1]); e->setVertex(1, vkf[kf2]); e->setVertex(2, vpt[P->obs_pt[i]]);
        e->setMeasurement(Eigen::Vector3d(P->obs_u[i], P->obs_v[i], ur));
        e->setInformation(Eigen::Matrix3d::Identity() * w);
        edge = e; kind[i] = 1;
      } else if (!stereo) {
        EdgeMono* e = new EdgeMono();
        e->setVertex(0, vkf[kf2]); e->setVertex(1, vpt[P->obs_pt[i]]);
        e->setMeasurement(Eigen::Vector2d(P->obs_u[i], P->obs_v[i]));
        e->setInformation(Eigen::Matrix2d::Identity() * w);
        edge = e; kind[i] = 2;
      } else {
        EdgeStereo* e = new EdgeStereo();
        e->setVertex(0, vkf[kf2]); e->setVertex(1, vpt[P->obs_pt[i]]);
        e->setMeasurement(Eigen::Vector3d(P->obs_u[i], P->obs_v[i], ur));
        e->setInformation(Eigen::Matrix3d::Identity() * w);
        edge = e; kind[i] = 3;
      }
      const double delta = stereo ? P->huber_stereo : P->huber_mono;
      if (delta > 0 && !(flags & GPBA_OBS_NO_KERNEL)) {
        g2o::RobustKernelHuber* rk = new g2o::RobustKernelHuber;
        edge->setRobustKernel(rk);
        rk->setDelta(delta);
      }
      if (flags & GPBA_OBS_LEVEL1) edge->setLevel(1);
      optimizer.addEdge(edge);
      eobs[i] = edge;
    }
  }
  bool depth_positive(int64_t i) const {   // the edge's own isDepthPositive(); the stereo edges have none (G2oTypes.h:404-468)
    if (kind[i] == 0) return static_cast<EdgeMonoGP*>(eobs[i])->isDepthPositive();
    if (kind[i] == 2) return static_cast<EdgeMono*>(eobs[i])->isDepthPositive();
    return true;
  }
  // LocalGPBA's inlier check (:1263-1348) on the edges' stored chi2: mono > chi2Mono (x1.5 for close points) or negative
  // depth, stereo > chi2Stereo
  void flags(const gpba_problem* P, const gpba_thresholds& th, uint8_t* fl) const {
    for (int64_t i = 0; i < P->n_obs; ++i) {
      const double c2 = eobs[i]->chi2();
      bool out;
      if (kind[i] == 1 || kind[i] == 3) out = c2 > th.chi2_stereo;
      else {
        const bool close = P->obs_flags && (P->obs_flags[i] & GPBA_OBS_CLOSE);
        out = (c2 > th.chi2_mono && !close) || (c2 > th.chi2_mono_close && close) || !depth_positive(i);
      }
      fl[i] = out ? 1 : 0;
    }
  }
  void read_back(const gpba_problem* P, double* kf_pose_out, double* kf_vel_out, double* pt_out, double* edge_chi2_out) const {
    for (int k = 0; k < P->n_kf; ++k) {
      if (kf_pose_out) to7(vkf[k]->estimate().Twb, kf_pose_out + 7 * k);
      if (kf_vel_out) for (int i = 0; i < 6; ++i) kf_vel_out[6 * k + i] = vkf[k]->estimate().Vel(i);
    }
    if (pt_out) for (int p = 0; p < P->n_pt; ++p) for (int i = 0; i < 3; ++i) pt_out[3 * p + i] = vpt[p]->estimate()(i);
    if (edge_chi2_out) for (int64_t i = 0; i < P->n_obs; ++i) edge_chi2_out[i] = eobs[i]->chi2();
  }
};

// The graph of LocalGPBA with extrinsic vertices (src/Optimizer.cc:983-995, 1100-1135): every asynchronous camera a
// VertexExtrinsic (fixed at first) with an EdgeExtrinsicPrior, every GP observation an EdgeMonoGPExtrinsic.
struct ExtGraph {
  GaussianProcess gp;
  std::vector<PinholeStandIn> cams;
  std::vector<GeometricCamera*> cam_ptrs;
  g2o::SparseOptimizer optimizer;
  g2o::OptimizationAlgorithmLevenberg* solver = nullptr;
  std::vector<VertexPoseVel*> vkf;
  std::vector<VertexExtrinsic*> vext;
  std::vector<g2o::VertexSBAPointXYZ*> vpt;
  ExtGraph(const gpba_problem* P, const double* prior_q, const double* prior_info) : gp(BaGraph::make_gp(P)) {
    for (int c = 0; c < P->n_cam; ++c) cams.emplace_back(P->cam_intr + 4 * c);
    for (auto& c : cams) cam_ptrs.push_back(&c);
    MultiKeyFrame::mTbc.clear();
    for (int c = 0; c < P->n_cam; ++c) MultiKeyFrame::mTbc.push_back(from7(P->cam_Tbc + 7 * c));
    MultiFrame::mTbc = MultiKeyFrame::mTbc;

    g2o::BlockSolverX::LinearSolverType* linearSolver = new g2o::LinearSolverDense<g2o::BlockSolverX::PoseMatrixType>();
    g2o::BlockSolverX* solver_ptr = new g2o::BlockSolverX(linearSolver);
    solver = new g2o::OptimizationAlgorithmLevenberg(solver_ptr);
    if (P->lambda_init > 0) solver->setUserLambdaInit(P->lambda_init);
    optimizer.setAlgorithm(solver);
    optimizer.setVerbose(false);

    vkf.resize(P->n_kf);
    for (int k = 0; k < P->n_kf; ++k) {
      PoseVelocity pv;
      pv.Twb = from7(P->kf_pose + 7 * k);
      for (int i = 0; i < 6; ++i) pv.Vel(i) = P->kf_vel[6 * k + i];
      pv.time = P->kf_time[k]; pv.bf = P->bf; pv.vpCameras = cam_ptrs;
      VertexPoseVel* v = new VertexPoseVel();
      v->setEstimate(pv);
      v->setId(k);
      v->setFixed(P->kf_fixed[k] != 0);
      optimizer.addVertex(v);
      vkf[k] = v;
    }
    for (int i = 0; i < P->n_velp; ++i) {
      EdgeVelocity* e = new EdgeVelocity();
      e->setVertex(0, vkf[P->velp_kf[i]]);
      e->setInformation(gp.mQcInv.block<1, 1>(2, 2));
      optimizer.addEdge(e);
    }
    for (int i = 0; i < P->n_prior; ++i) {
      EdgeGaussianPrior* e = new EdgeGaussianPrior();
      e->setVertex(0, vkf[P->prior_kf1[i]]);
      e->setVertex(1, vkf[P->prior_kf2[i]]);
      if (P->huber_prior > 0) { g2o::RobustKernelHuber* rk = new g2o::RobustKernelHuber; e->setRobustKernel(rk); rk->setDelta(P->huber_prior); }
      e->setInformation(gp.QiInv(P->kf_time[P->prior_kf2[i]] - P->kf_time[P->prior_kf1[i]]));
      optimizer.addEdge(e);
    }
    // ---- extrinsic vertices with their rotation priors (:983-995), ids between the keyframes and the points
    vext.resize(P->n_cam - 1);
    for (int c = 0; c < P->n_cam - 1; ++c) {
      VertexExtrinsic* v = new VertexExtrinsic(MultiKeyFrame::mTbc[c].cast<double>());
      v->setId(P->n_kf + c);
      v->setFixed(true);
      optimizer.addVertex(v);
      vext[c] = v;
      EdgeExtrinsicPrior* e = new EdgeExtrinsicPrior(Sophus::SO3d::fromQuaternion(prior_q[4 * c], prior_q[4 * c + 1], prior_q[4 * c + 2], prior_q[4 * c + 3]));
      e->setVertex(0, v);
      Eigen::Matrix3d W;
      for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) W(i, j) = prior_info[9 * c + 3 * i + j];
      e->setInformation(W);
      optimizer.addEdge(e);
    }
    vpt.resize(P->n_pt);
    for (int p = 0; p < P->n_pt; ++p) {
      g2o::VertexSBAPointXYZ* vP = new g2o::VertexSBAPointXYZ();
      vP->setEstimate(Eigen::Vector3d(P->pt_xyz[3 * p], P->pt_xyz[3 * p + 1], P->pt_xyz[3 * p + 2]));
      vP->setId(P->n_kf + P->n_cam + p);
      vP->setMarginalized(true);
      optimizer.addVertex(vP);
      vpt[p] = vP;
    }
    for (int64_t i = 0; i < P->n_obs; ++i) {
      const int r = P->obs_rec[i], kf1 = P->rec_kf1[r], kf2 = P->rec_kf2[r], cam = P->rec_cam[r];
      const double ur = P->obs_ur ? P->obs_ur[i] : -1.0, w = P->obs_inv_sigma2[i];
      const unsigned flags = P->obs_flags ? P->obs_flags[i] : 0u;
      g2o::OptimizableGraph::Edge* edge = nullptr;
      double delta = P->huber_mono;
      if (kf1 >= 0) {   // :1118-1135
        EdgeMonoGPExtrinsic* e = new EdgeMonoGPExtrinsic(cam, P->rec_t[r], &gp);
        e->setVertex(0, vkf[kf1]); e->setVertex(1, vkf[kf2]); e->setVertex(2, vpt[P->obs_pt[i]]); e->setVertex(3, vext[cam]);
        e->setMeasurement(Eigen::Vector2d(P->obs_u[i], P->obs_v[i]));
        e->setInformation(Eigen::Matrix2d::Identity() * w);
        edge = e;
      } else if (ur < 0) {
        EdgeMono* e = new EdgeMono();
        e->setVertex(0, vkf[kf2]); e->setVertex(1, vpt[P->obs_pt[i]]);
        e->setMeasurement(Eigen::Vector2d(P->obs_u[i], P->obs_v[i]));
        e->setInformation(Eigen::Matrix2d::Identity() * w);
        edge = e;
      } else {
        EdgeStereo* e = new EdgeStereo();
        e->setVertex(0, vkf[kf2]); e->setVertex(1, vpt[P->obs_pt[i]]);
        e->setMeasurement(Eigen::Vector3d(P->obs_u[i], P->obs_v[i], ur));
        e->setInformation(Eigen::Matrix3d::Identity() * w);
        edge = e;
        delta = P->huber_stereo;
      }
      if (delta > 0 && !(flags & GPBA_OBS_NO_KERNEL)) { g2o::RobustKernelHuber* rk = new g2o::RobustKernelHuber; edge->setRobustKernel(rk); rk->setDelta(delta); }
      if (flags & GPBA_OBS_LEVEL1) edge->setLevel(1);
      optimizer.addEdge(edge);
    }
  }
  void read_back(const gpba_problem* P, double* kf_pose_out, double* kf_vel_out, double* pt_out, double* Tbc_out) const {
    for (int k = 0; k < P->n_kf; ++k) {
      if (kf_pose_out) to7(vkf[k]->estimate().Twb, kf_pose_out + 7 * k);
      if (kf_vel_out) for (int i = 0; i < 6; ++i) kf_vel_out[6 * k + i] = vkf[k]->estimate().Vel(i);
    }
    if (pt_out) for (int p = 0; p < P->n_pt; ++p) for (int i = 0; i < 3; ++i) pt_out[3 * p + i] = vpt[p]->estimate()(i);
    if (Tbc_out) {
      for (int c = 0; c < P->n_cam - 1; ++c) to7(vext[c]->estimate(), Tbc_out + 7 * c);
      to7(MultiKeyFrame::mTbc.back(), Tbc_out + 7 * (P->n_cam - 1));
    }
  }
};

}  // namespace
