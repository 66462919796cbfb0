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
// oracle/_ref/libamc_ref_g2o.so and compared with the oracle in tests/test_whole_path_reference.py.
#define REF_G2O_DEFINE_STATICS
#include "ref_g2o_graph.h"

extern "C" {

// Returns what SparseOptimizer::optimize returns (the number of iterations run, 0 on Fail, -1 when nothing is free).
// Outputs (any may be NULL): the estimates after the run, the stored chi2 of every reprojection edge (obs order), the trace
// (chi2_before[0] = robust chi2 at the start; chi2_after[i] = chi2 of the stored errors after iteration i, see Recorder),
// sizes[0..3] = #active vertices, #active edges, Hessian dimension of the poses, of the landmarks; flags_out = LocalGPBA's
// inlier check on the final graph (thresholds th, :1263-1348).
int ref_g2o_optimize(const gpba_problem* P, int iters, int max_trials, double* kf_pose_out, double* kf_vel_out, double* pt_out,
                     double* edge_chi2_out, gpba_lm_trace* tr, int64_t* sizes, const gpba_thresholds* th, uint8_t* flags_out) {
  BaGraph G(P, max_trials);
  g2o::SparseOptimizer& optimizer = G.optimizer;
  // ---- optimize (:333-335 / 1253-1256)
  if (tr) { std::memset(tr, 0, sizeof(*tr)); tr->result = GPBA_RESULT_OK; }
  Recorder rec;
  rec.opt = &optimizer; rec.alg = G.solver; rec.tr = tr;
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
  G.read_back(P, kf_pose_out, kf_vel_out, pt_out, edge_chi2_out);   // (:338-367 write the estimates into the map)
  if (th && flags_out) G.flags(P, *th, flags_out);
  optimizer.removePostIterationAction(&rec);
  return n;
}

// BASELINE config C3's schedule: n_rounds x (initializeOptimization(0) + optimize(iters)), after each round the errors of
// the excluded edges are recomputed ("if (mvbOutlier[idx]) e->computeError()", :591-592), every edge is re-flagged with
// LocalGPBA's inlier check, flagged edges move to level 1 and after the third round the kernels are removed
// ("if (it==2) e->setRobustKernel(0)") -- the round structure of src/Optimizer.cc:548-675 on the BA graph, with the real
// solver and edges.  traces [n_rounds]; flags_out [n_obs] after the last round.
int ref_g2o_rejection_rounds(const gpba_problem* P, int n_rounds, int iters, const gpba_thresholds* th, double* kf_pose_out,
                             double* kf_vel_out, double* pt_out, double* edge_chi2_out, uint8_t* flags_out, gpba_lm_trace* traces) {
  BaGraph G(P, 0);
  g2o::SparseOptimizer& optimizer = G.optimizer;
  std::vector<uint8_t> fl((size_t)P->n_obs, 0);
  for (int64_t i = 0; i < P->n_obs; ++i) fl[i] = (P->obs_flags && (P->obs_flags[i] & GPBA_OBS_LEVEL1)) ? 1 : 0;
  Recorder rec;
  rec.opt = &optimizer; rec.alg = G.solver;
  optimizer.addPostIterationAction(&rec);
  for (int it = 0; it < n_rounds; ++it) {
    gpba_lm_trace* tr = traces ? traces + it : nullptr;
    if (tr) { std::memset(tr, 0, sizeof(*tr)); tr->result = GPBA_RESULT_OK; }
    rec.tr = tr;
    optimizer.initializeOptimization(0);
    if (tr) { optimizer.computeActiveErrors(); tr->chi2_before[0] = optimizer.activeRobustChi2(); }
    const int n = optimizer.optimize(iters);
    if (tr) tr->n_iters = n;
    for (int64_t i = 0; i < P->n_obs; ++i) if (fl[i]) G.eobs[i]->computeError();
    G.flags(P, *th, fl.data());
    for (int64_t i = 0; i < P->n_obs; ++i) {
      G.eobs[i]->setLevel(fl[i] ? 1 : 0);
      if (it == 2) G.eobs[i]->setRobustKernel(0);
    }
  }
  optimizer.removePostIterationAction(&rec);
  G.read_back(P, kf_pose_out, kf_vel_out, pt_out, edge_chi2_out);
  if (flags_out) std::memcpy(flags_out, fl.data(), fl.size());
  return 0;
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

// Optimizer::PoseGPOptimizationFromeLastFrame (src/Optimizer.cc:369-686) for frame f of a batch: the real graph (two
// VertexPoseVel, EdgeMonoGPOnlyPose / EdgeMonoOnlyPose / EdgeStereoOnlyPose with Huber kernels, EdgeGaussianPrior, two
// EdgeVelocity; BlockSolverX + LinearSolverDense + LM with the default lambda, :373-382) and the four rounds of
// optimize(10) + re-flagging (:545-670), the round logic restated from those lines with the reference's float-typed
// thresholds.  Outputs as gpba_pose_optimize: states, mvbOutlier per match of the frame, nInitialCorrespondences - nBad.
int ref_g2o_pose_optimize(const gpba_pose_batch* B, int f, double* cur_pose_out, double* cur_vel_out, double* prev_pose_out,
                          double* prev_vel_out, uint8_t* outlier_out, int32_t* n_inliers_out, gpba_lm_trace* traces) {
  Eigen::Matrix<double, 6, 6> Qc = Eigen::Matrix<double, 6, 6>::Zero();
  for (int i = 0; i < 6; ++i) Qc(i, i) = B->qc[i];
  GaussianProcess gp(Qc);
  std::vector<PinholeStandIn> cams;
  for (int c = 0; c < B->n_cam; ++c) cams.emplace_back(B->cam_intr + 4 * c);
  std::vector<GeometricCamera*> cam_ptrs;
  for (auto& c : cams) cam_ptrs.push_back(&c);
  MultiKeyFrame::mTbc.clear();
  for (int c = 0; c < B->n_cam; ++c) MultiKeyFrame::mTbc.push_back(from7(B->cam_Tbc + 7 * c));
  MultiFrame::mTbc = MultiKeyFrame::mTbc;

  g2o::SparseOptimizer optimizer;
  g2o::BlockSolverX::LinearSolverType* linearSolver = new g2o::LinearSolverDense<g2o::BlockSolverX::PoseMatrixType>();
  g2o::BlockSolverX* solver_ptr = new g2o::BlockSolverX(linearSolver);
  g2o::OptimizationAlgorithmLevenberg* solver = new g2o::OptimizationAlgorithmLevenberg(solver_ptr);
  optimizer.setAlgorithm(solver);
  optimizer.setVerbose(false);

  auto make_pv = [&](const double* T7, const double* vel, double time) {
    PoseVelocity pv;
    pv.Twb = from7(T7);
    for (int i = 0; i < 6; ++i) pv.Vel(i) = vel[i];
    pv.time = time; pv.bf = B->bf; pv.vpCameras = cam_ptrs;
    return pv;
  };
  VertexPoseVel* v1 = new VertexPoseVel();
  v1->setEstimate(make_pv(B->prev_pose + 7 * f, B->prev_vel + 6 * f, B->prev_time[f]));
  v1->setFixed(B->prev_fixed[f] != 0);
  v1->setId(0);
  optimizer.addVertex(v1);
  VertexPoseVel* v2 = new VertexPoseVel();
  v2->setEstimate(make_pv(B->cur_pose + 7 * f, B->cur_vel + 6 * f, B->cur_time[f]));
  v2->setFixed(false);
  v2->setId(1);
  optimizer.addVertex(v2);

  const int64_t ob = B->obs_begin[f], oe = B->obs_begin[f + 1];
  std::vector<uint8_t> outlier((size_t)(oe - ob), 0);
  std::vector<EdgeStereoOnlyPose*> eS; std::vector<EdgeMonoOnlyPose*> eM; std::vector<EdgeMonoGPOnlyPose*> eG;
  std::vector<int64_t> iS, iM, iG;
  const float thHuberMono = (float)B->huber_mono, thHuberStereo = (float)B->huber_stereo;
  for (int64_t i = ob; i < oe; ++i) {
    const int cam_idx = B->obs_cam[i];
    const Eigen::Vector3f Xw((float)B->obs_xw[3 * i], (float)B->obs_xw[3 * i + 1], (float)B->obs_xw[3 * i + 2]);
    const double w = B->obs_inv_sigma2[i];
    const bool out0 = (B->obs_flags && (B->obs_flags[i] & GPBA_OBS_LEVEL1));
    outlier[i - ob] = out0;
    if (cam_idx != B->n_cam - 1) {
      EdgeMonoGPOnlyPose* e = new EdgeMonoGPOnlyPose(Xw, cam_idx, B->cam_time[(size_t)f * B->n_cam + cam_idx], &gp);
      e->setVertex(0, v1); e->setVertex(1, v2);
      e->setMeasurement(Eigen::Vector2d(B->obs_u[i], B->obs_v[i]));
      e->setInformation(Eigen::Matrix2d::Identity() * w);
      g2o::RobustKernelHuber* rk = new g2o::RobustKernelHuber;
      e->setRobustKernel(rk); rk->setDelta(thHuberMono);
      e->setLevel(out0 ? 1 : 0);
      optimizer.addEdge(e);
      eG.push_back(e); iG.push_back(i);
    } else if (!(B->obs_ur && B->obs_ur[i] >= 0)) {
      EdgeMonoOnlyPose* e = new EdgeMonoOnlyPose(Xw);
      e->setVertex(0, v2);
      e->setMeasurement(Eigen::Vector2d(B->obs_u[i], B->obs_v[i]));
      e->setInformation(Eigen::Matrix2d::Identity() * w);
      g2o::RobustKernelHuber* rk = new g2o::RobustKernelHuber;
      e->setRobustKernel(rk); rk->setDelta(thHuberMono);
      e->setLevel(out0 ? 1 : 0);
      optimizer.addEdge(e);
      eM.push_back(e); iM.push_back(i);
    } else {
      EdgeStereoOnlyPose* e = new EdgeStereoOnlyPose(Xw);
      e->setVertex(0, v2);
      e->setMeasurement(Eigen::Vector3d(B->obs_u[i], B->obs_v[i], B->obs_ur[i]));
      e->setInformation(Eigen::Matrix3d::Identity() * w);
      g2o::RobustKernelHuber* rk = new g2o::RobustKernelHuber;
      e->setRobustKernel(rk); rk->setDelta(thHuberStereo);
      e->setLevel(out0 ? 1 : 0);
      optimizer.addEdge(e);
      eS.push_back(e); iS.push_back(i);
    }
  }
  const int nInitialCorrespondences = (int)(oe - ob);
  EdgeGaussianPrior* egp = new EdgeGaussianPrior();
  egp->setVertex(0, v1); egp->setVertex(1, v2);
  egp->setInformation(gp.QiInv(B->cur_time[f] - B->prev_time[f]));
  optimizer.addEdge(egp);
  EdgeVelocity* ev1 = new EdgeVelocity();
  ev1->setVertex(0, v1); ev1->setInformation(gp.mQcInv.block<1, 1>(2, 2));
  optimizer.addEdge(ev1);
  EdgeVelocity* ev2 = new EdgeVelocity();
  ev2->setVertex(0, v2); ev2->setInformation(gp.mQcInv.block<1, 1>(2, 2));
  optimizer.addEdge(ev2);

  const float chi2Mono[4] = {5.991, 5.991, 5.991, 5.991};
  const float chi2Stereo[4] = {15.6f, 9.8f, 7.815f, 7.815f};
  const int its[4] = {10, 10, 10, 10};
  int nBad = 0;
  Recorder rec;
  rec.opt = &optimizer; rec.alg = solver;
  optimizer.addPostIterationAction(&rec);
  for (size_t it = 0; it < 4; ++it) {
    gpba_lm_trace* tr = traces ? traces + it : nullptr;
    if (tr) { std::memset(tr, 0, sizeof(*tr)); tr->result = GPBA_RESULT_OK; }
    rec.tr = tr;
    optimizer.initializeOptimization(0);
    const int n = optimizer.optimize(its[it]);
    if (tr) tr->n_iters = n;
    nBad = 0;
    const float chi2close = 1.5 * chi2Mono[it];
    for (size_t i = 0; i < eG.size(); ++i) {
      EdgeMonoGPOnlyPose* e = eG[i];
      const int64_t idx = iG[i];
      const bool bclose = B->obs_flags && (B->obs_flags[idx] & GPBA_OBS_CLOSE);
      if (outlier[idx - ob]) e->computeError();
      const float chi2 = e->chi2();
      if ((chi2 > chi2Mono[it] && !bclose) || (bclose && chi2 > chi2close) || !e->isDepthPositive()) { outlier[idx - ob] = 1; e->setLevel(1); nBad++; }
      else { outlier[idx - ob] = 0; e->setLevel(0); }
      if (it == 2) e->setRobustKernel(0);
    }
    for (size_t i = 0; i < eS.size(); ++i) {
      EdgeStereoOnlyPose* e = eS[i];
      const int64_t idx = iS[i];
      if (outlier[idx - ob]) e->computeError();
      const float chi2 = e->chi2();
      if (chi2 > chi2Stereo[it]) { outlier[idx - ob] = 1; e->setLevel(1); nBad++; }
      else { outlier[idx - ob] = 0; e->setLevel(0); }
      if (it == 2) e->setRobustKernel(0);
    }
    for (size_t i = 0; i < eM.size(); ++i) {
      EdgeMonoOnlyPose* e = eM[i];
      const int64_t idx = iM[i];
      const bool bclose = B->obs_flags && (B->obs_flags[idx] & GPBA_OBS_CLOSE);
      if (outlier[idx - ob]) e->computeError();
      const float chi2 = e->chi2();
      if ((chi2 > chi2Mono[it] && !bclose) || (bclose && chi2 > chi2close) || !e->isDepthPositive()) { outlier[idx - ob] = 1; e->setLevel(1); nBad++; }
      else { outlier[idx - ob] = 0; e->setLevel(0); }
      if (it == 2) e->setRobustKernel(0);
    }
    if (optimizer.edges().size() < 10) break;
  }
  optimizer.removePostIterationAction(&rec);
  if (cur_pose_out) to7(v2->estimate().Twb, cur_pose_out);
  if (cur_vel_out) for (int i = 0; i < 6; ++i) cur_vel_out[i] = v2->estimate().Vel(i);
  if (prev_pose_out) to7(v1->estimate().Twb, prev_pose_out);
  if (prev_vel_out) for (int i = 0; i < 6; ++i) prev_vel_out[i] = v1->estimate().Vel(i);
  if (outlier_out) std::memcpy(outlier_out, outlier.data(), outlier.size());
  if (n_inliers_out) *n_inliers_out = nInitialCorrespondences - nBad;
  return 0;
}

// Optimizer::OptimizeVel (src/Optimizer.cc:2364-2447) for hypothesis h of a velocity-RANSAC batch: one VertexVel, one
// EdgeVelReproj per match (Huber 5.991, level 0 only for the sampled matches), BlockSolverX + LinearSolverDense + LM with the
// default lambda, optimize(40), then every edge re-evaluated and counted as inlier when |e| <= threshold.  The GN_ITERS
// constant is the batch's `iterations` (40).  Returns the inlier count; vel_out[6], mask_out[n_match].
int ref_g2o_optimize_vel(const gpba_vel_batch* B, int h, double* vel_out, uint8_t* mask_out, gpba_lm_trace* tr) {
  std::vector<PinholeStandIn> cams;
  for (int c = 0; c < B->n_cam; ++c) cams.emplace_back(B->cam_intr + 4 * c);
  MultiKeyFrame::mTbc.clear();
  for (int c = 0; c < B->n_cam; ++c) MultiKeyFrame::mTbc.push_back(from7(B->cam_Tbc + 7 * c));
  MultiFrame::mTbc = MultiKeyFrame::mTbc;
  MultiFrame F1;
  for (auto& c : cams) F1.mvpCamera.push_back(&c);

  g2o::SparseOptimizer optimizer;
  g2o::BlockSolverX::LinearSolverType* linearSolver = new g2o::LinearSolverDense<g2o::BlockSolverX::PoseMatrixType>();
  g2o::BlockSolverX* solver_ptr = new g2o::BlockSolverX(linearSolver);
  g2o::OptimizationAlgorithmLevenberg* solver = new g2o::OptimizationAlgorithmLevenberg(solver_ptr);
  optimizer.setAlgorithm(solver);
  int inliers = 0;
  VertexVel* vVel = new VertexVel();
  V6 v0;
  for (int i = 0; i < 6; ++i) v0(i) = B->vel_init[i];
  vVel->setEstimate(v0);
  vVel->setId(0);
  vVel->setFixed(false);
  optimizer.addVertex(vVel);
  const int32_t* set = B->samples + (size_t)h * B->set_size;
  std::vector<EdgeVelReproj*> vpEdges;
  for (int i = 0; i < B->n_match; ++i) {
    const int cam = B->obs_cam[i];
    EdgeVelReproj* e = new EdgeVelReproj(from7(B->last_pose), B->cam_dt[cam],
                                         Eigen::Vector3d(B->obs_xw[3 * i], B->obs_xw[3 * i + 1], B->obs_xw[3 * i + 2]), cam, &F1);
    e->setVertex(0, dynamic_cast<g2o::OptimizableGraph::Vertex*>(optimizer.vertex(0)));
    e->setMeasurement(Eigen::Vector2d(B->obs_u[i], B->obs_v[i]));
    e->setInformation(Eigen::Matrix2d::Identity() * B->obs_inv_sigma2[i]);
    g2o::RobustKernelHuber* rk = new g2o::RobustKernelHuber;
    e->setRobustKernel(rk);
    rk->setDelta(B->huber_delta);
    bool sampled = false;
    for (int k = 0; k < B->set_size; ++k) sampled = sampled || set[k] == i;
    e->setLevel(sampled ? 0 : 1);
    optimizer.addEdge(e);
    vpEdges.push_back(e);
  }
  if (tr) { std::memset(tr, 0, sizeof(*tr)); tr->result = GPBA_RESULT_OK; }
  Recorder rec;
  rec.opt = &optimizer; rec.alg = solver; rec.tr = tr;
  optimizer.addPostIterationAction(&rec);
  optimizer.initializeOptimization(0);
  const int n = optimizer.optimize(B->iterations);
  if (tr) tr->n_iters = n;
  optimizer.removePostIterationAction(&rec);
  for (size_t i = 0; i < vpEdges.size(); ++i) {
    vpEdges[i]->computeError();
    const bool in = vpEdges[i]->error().norm() <= B->threshold;
    inliers += in;
    if (mask_out) mask_out[i] = in ? 1 : 0;
  }
  if (vel_out) for (int i = 0; i < 6; ++i) vel_out[i] = vVel->estimate()(i);
  return inliers;
}

// LocalGPBA with extrinsic self-calibration (src/Optimizer.cc:983-995, 1100-1135, 1222-1240): every asynchronous camera is a
// VertexExtrinsic with an EdgeExtrinsicPrior (R_ini, information 3x3), every GP observation an EdgeMonoGPExtrinsic on
// (KF_prev, KF_cur, point, extrinsic).  Stage 1: optimize(it1) with the extrinsics fixed; stage 2: the cameras flagged in
// `ext_free` (the caller applies the >= 50 observations rule, :1224-1235) are released, initializeOptimization(),
// computeActiveErrors(), optimize(it2).  Tbc_out [n_cam][7]; tr1 / tr2 the traces of the two stages.
int ref_g2o_local_gpba_ext(const gpba_problem* P, const uint8_t* ext_free, const double* prior_q, const double* prior_info, int it1,
                           int it2, double* kf_pose_out, double* kf_vel_out, double* pt_out, double* Tbc_out, gpba_lm_trace* tr1,
                           gpba_lm_trace* tr2) {
  ExtGraph G(P, prior_q, prior_info);
  // ---- stage 1 (:1219-1224)
  Recorder rec;
  rec.opt = &G.optimizer; rec.alg = G.solver;
  G.optimizer.addPostIterationAction(&rec);
  if (tr1) { std::memset(tr1, 0, sizeof(*tr1)); tr1->result = GPBA_RESULT_OK; }
  rec.tr = tr1;
  G.optimizer.initializeOptimization();
  G.optimizer.computeActiveErrors();
  if (tr1) tr1->chi2_before[0] = G.optimizer.activeRobustChi2();
  int n = G.optimizer.optimize(it1);
  if (tr1) tr1->n_iters = n;
  // ---- stage 2 (:1228-1240)
  for (int c = 0; c < P->n_cam - 1; ++c) if (ext_free[c]) G.vext[c]->setFixed(false);
  if (tr2) { std::memset(tr2, 0, sizeof(*tr2)); tr2->result = GPBA_RESULT_OK; }
  rec.tr = tr2;
  G.optimizer.initializeOptimization();
  G.optimizer.computeActiveErrors();
  if (tr2) tr2->chi2_before[0] = G.optimizer.activeRobustChi2();
  n = G.optimizer.optimize(it2);
  if (tr2) tr2->n_iters = n;
  G.optimizer.removePostIterationAction(&rec);
  G.read_back(P, kf_pose_out, kf_vel_out, pt_out, Tbc_out);
  return n;
}

}  // extern "C"
