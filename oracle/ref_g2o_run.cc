// TEST INFRASTRUCTURE, not product code.  THE REFERENCE'S WHOLE OPTIMISATION PATH, run from its own sources: this file builds
// the g2o graph of Optimizer::BundleAdjustment / Optimizer::LocalGPBA (src/Optimizer.cc:61-262, 713-1262) out of a flattened
// gpba_problem -- real VertexPoseVel / VertexSBAPointXYZ vertices, real EdgeMonoGP / EdgeStereoGP / EdgeMono / EdgeStereo /
// EdgeGaussianPrior / EdgeVelocity edges with RobustKernelHuber, information matrices and levels set as the reference sets
// them -- and calls the real g2o::SparseOptimizer::optimize with BlockSolverX + LinearSolverDense +
// OptimizationAlgorithmLevenberg.  Everything that runs is the reference's code, compiled UNMODIFIED from /root/reference:
//   Thirdparty/g2o/g2o/core/*.cpp (hyper graph, optimizable graph, sparse optimizer, block solver, sparse block matrices,
//   base_{unary,binary,multi}_edge.hpp quadratic forms, jacobian workspace, robust kernels, LM), g2o/stuff/*.cpp,
//   g2o/types/types_sba.cpp, solvers/linear_solver_dense.h, src/G2oTypes.cc, src/GaussianProcess.cc, src/Pose3utils.cc
// against the stand-in headers of oracle/ref_shim/ for the absent Eigen3 / Sophus / OpenCV / Boost / map classes (what
// those stand-ins are, and that they are checked against numpy / scipy: oracle/ref_shim/Eigen/Core, tests/test_ref_shim.py).
// Stand-in arithmetic on this path: dense products / inverses, the pivoted LDLT behind LinearSolverDense (third party in the
// reference too: Eigen::LDLT), quaternion and SE(3) exp / log, the pinhole projection.  What this file restates is only the
// graph CONSTRUCTION (which vertices and edges exist is the input); built by `make -C oracle _ref` into
// oracle/_ref/libamc_ref_g2o.so and compared with the oracle in tests/test_ref_g2o.py.
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

extern "C" {

// Returns what SparseOptimizer::optimize returns (the number of iterations run, 0 on Fail, -1 when nothing is free).
// Outputs (any may be NULL): the estimates after the run, the stored chi2 of every reprojection edge (obs order), the trace
// (chi2_before[0] = robust chi2 at the start; chi2_after[i] = chi2 of the stored errors after iteration i, see Recorder),
// sizes[0..3] = #active vertices, #active edges, Hessian dimension of the poses, of the landmarks.
int ref_g2o_optimize(const gpba_problem* P, int iters, int max_trials, double* kf_pose_out, double* kf_vel_out, double* pt_out,
                     double* edge_chi2_out, gpba_lm_trace* tr, int64_t* sizes) {
  Eigen::Matrix<double, 6, 6> Qc = Eigen::Matrix<double, 6, 6>::Zero();
  for (int i = 0; i < 6; ++i) Qc(i, i) = P->qc[i];
  GaussianProcess gp(Qc);
  std::vector<PinholeStandIn> cams;
  for (int c = 0; c < P->n_cam; ++c) cams.emplace_back(P->cam_intr + 4 * c);
  std::vector<GeometricCamera*> cam_ptrs;
  for (auto& c : cams) cam_ptrs.push_back(&c);
  MultiKeyFrame::mTbc.clear();
  for (int c = 0; c < P->n_cam; ++c) MultiKeyFrame::mTbc.push_back(from7(P->cam_Tbc + 7 * c));   // reference camera last
  MultiFrame::mTbc = MultiKeyFrame::mTbc;

  // ---- Optimizer.cc:66-77 / 838-856
  g2o::SparseOptimizer optimizer;
  g2o::BlockSolverX::LinearSolverType* linearSolver = new g2o::LinearSolverDense<g2o::BlockSolverX::PoseMatrixType>();
  g2o::BlockSolverX* solver_ptr = new g2o::BlockSolverX(linearSolver);
  g2o::OptimizationAlgorithmLevenberg* solver = new g2o::OptimizationAlgorithmLevenberg(solver_ptr);
  if (P->lambda_init > 0) solver->setUserLambdaInit(P->lambda_init);
  if (max_trials > 0) solver->setMaxTrialsAfterFailure(max_trials);
  optimizer.setAlgorithm(solver);
  optimizer.setVerbose(false);

  // ---- keyframe vertices (:82-95): id = index (ascending id = Hessian order)
  std::vector<VertexPoseVel*> vkf(P->n_kf);
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
  std::vector<g2o::VertexSBAPointXYZ*> vpt(P->n_pt);
  for (int p = 0; p < P->n_pt; ++p) {
    g2o::VertexSBAPointXYZ* vP = new g2o::VertexSBAPointXYZ();
    vP->setEstimate(Eigen::Vector3d(P->pt_xyz[3 * p], P->pt_xyz[3 * p + 1], P->pt_xyz[3 * p + 2]));
    vP->setId(P->n_kf + p);
    vP->setMarginalized(true);
    optimizer.addVertex(vP);
    vpt[p] = vP;
  }
  // ---- reprojection edges in insertion order (:168-330)
  std::vector<g2o::OptimizableGraph::Edge*> eobs((size_t)P->n_obs);
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
      edge = e;
    } else if (kf1 >= 0) {
      EdgeStereoGP* e = new EdgeStereoGP(cam, P->rec_t[r], &gp);
      e->setVertex(0, vkf[kf1]); e->setVertex(1, vkf[kf2]); e->setVertex(2, vpt[P->obs_pt[i]]);
      e->setMeasurement(Eigen::Vector3d(P->obs_u[i], P->obs_v[i], ur));
      e->setInformation(Eigen::Matrix3d::Identity() * w);
      edge = e;
    } else if (!stereo) {
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

  // ---- optimize (:333-335 / 1253-1256)
  if (tr) { std::memset(tr, 0, sizeof(*tr)); tr->result = GPBA_RESULT_OK; }
  Recorder rec;
  rec.opt = &optimizer; rec.alg = solver; rec.tr = tr;
  optimizer.addPostIterationAction(&rec);
  optimizer.initializeOptimization(0);
  if (tr) { optimizer.computeActiveErrors(); tr->chi2_before[0] = optimizer.activeRobustChi2(); }
  const int n = optimizer.optimize(iters);
  if (tr) tr->n_iters = n;
  if (sizes) {
    sizes[0] = (int64_t)optimizer.activeVertices().size(); sizes[1] = (int64_t)optimizer.activeEdges().size();
    int64_t dp = 0, dl = 0;
    for (auto* v : optimizer.indexMapping()) (v->marginalized() ? dl : dp) += v->dimension();
    sizes[2] = dp; sizes[3] = dl;
  }
  // ---- read back (:338-367 write the estimates into the map)
  for (int k = 0; k < P->n_kf; ++k) {
    if (kf_pose_out) to7(vkf[k]->estimate().Twb, kf_pose_out + 7 * k);
    if (kf_vel_out) for (int i = 0; i < 6; ++i) kf_vel_out[6 * k + i] = vkf[k]->estimate().Vel(i);
  }
  if (pt_out) for (int p = 0; p < P->n_pt; ++p) for (int i = 0; i < 3; ++i) pt_out[3 * p + i] = vpt[p]->estimate()(i);
  if (edge_chi2_out) for (int64_t i = 0; i < P->n_obs; ++i) edge_chi2_out[i] = eobs[i]->chi2();
  optimizer.removePostIterationAction(&rec);
  return n;
}

// The optimisation inside Optimizer::OptimizeEssentialGraph (src/Optimizer.cc:1434-1717): real VertexSim3Expmap / EdgeSim3
// (Thirdparty/g2o/g2o/types/types_seven_dof_expmap.{h,cpp}, sim3.h; EdgeSim3 has no analytic Jacobian, so g2o's numeric
// differentiation of base_binary_edge.hpp runs), BlockSolver_7_3, Levenberg-Marquardt with lambda_0 = 1e-16, identity
// information (:1505), LinearSolverDense in place of LinearSolverEigen.  sim3_out [n_kf][8] = the optimised S_iw.
int ref_g2o_pose_graph(const gpba_pose_graph* G, int iters, double* sim3_out, gpba_lm_trace* tr) {
  g2o::SparseOptimizer optimizer;
  optimizer.setVerbose(false);
  g2o::BlockSolver_7_3::LinearSolverType* linearSolver = new g2o::LinearSolverDense<g2o::BlockSolver_7_3::PoseMatrixType>();
  g2o::BlockSolver_7_3* solver_ptr = new g2o::BlockSolver_7_3(linearSolver);
  g2o::OptimizationAlgorithmLevenberg* solver = new g2o::OptimizationAlgorithmLevenberg(solver_ptr);
  if (G->lambda_init > 0) solver->setUserLambdaInit(G->lambda_init);
  optimizer.setAlgorithm(solver);
  std::vector<g2o::VertexSim3Expmap*> v(G->n_kf);
  for (int k = 0; k < G->n_kf; ++k) {
    const double* p = G->sim3 + 8 * k;
    g2o::VertexSim3Expmap* V = new g2o::VertexSim3Expmap();
    V->setEstimate(g2o::Sim3(Eigen::Quaterniond(p[3], p[0], p[1], p[2]), Eigen::Vector3d(p[4], p[5], p[6]), p[7]));
    if (G->fixed[k]) V->setFixed(true);
    V->setId(k);
    V->setMarginalized(false);
    V->_fix_scale = G->fix_scale != 0;
    optimizer.addVertex(V);
    v[k] = V;
  }
  const Eigen::Matrix<double, 7, 7> matLambda = Eigen::Matrix<double, 7, 7>::Identity();
  for (int64_t e = 0; e < G->n_edge; ++e) {
    const double* p = G->edge_meas + 8 * e;
    g2o::EdgeSim3* E = new g2o::EdgeSim3();
    E->setVertex(1, v[G->edge_j[e]]);
    E->setVertex(0, v[G->edge_i[e]]);
    E->setMeasurement(g2o::Sim3(Eigen::Quaterniond(p[3], p[0], p[1], p[2]), Eigen::Vector3d(p[4], p[5], p[6]), p[7]));
    E->information() = matLambda;
    optimizer.addEdge(E);
  }
  if (tr) { std::memset(tr, 0, sizeof(*tr)); tr->result = GPBA_RESULT_OK; }
  Recorder rec;
  rec.opt = &optimizer; rec.alg = solver; rec.tr = tr;
  optimizer.addPostIterationAction(&rec);
  optimizer.initializeOptimization();
  if (tr) { optimizer.computeActiveErrors(); tr->chi2_before[0] = optimizer.activeRobustChi2(); }
  const int n = optimizer.optimize(iters);
  if (tr) tr->n_iters = n;
  if (sim3_out)
    for (int k = 0; k < G->n_kf; ++k) for (int i = 0; i < 8; ++i) sim3_out[8 * k + i] = v[k]->estimate()[i];
  optimizer.removePostIterationAction(&rec);
  return n;
}

}  // extern "C"
