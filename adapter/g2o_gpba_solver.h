// adapter/g2o_gpba_solver.h -- the reference-side binding of libgpba.so (include/gpba.h).
//
// Header-only C++11, to be compiled INSIDE the AMC-SLAM tree (it needs the reference's own g2o, Eigen, Sophus and
// G2oTypes.h).  In this repository it is compiled against the reference's real g2o headers with stand-in Eigen / Sophus
// headers and exercised on real g2o graphs by oracle/ref_adapter_check.cc (tests/test_whole_path_reference.py).  It is the code a
// maintainer adds to swap solvers at the two construction sites, without touching Tracking / LocalMapping /
// LoopClosing (see INTEGRATION.md for the exact diff):
//
//   src/Optimizer.cc:68-76    linearSolver = new g2o::LinearSolverEigen<...>; solver_ptr = new g2o::BlockSolverX(linearSolver);
//                             solver = new g2o::OptimizationAlgorithmLevenberg(solver_ptr); solver->setUserLambdaInit(1e-5);
//   src/Optimizer.cc:840-856  same with g2o::LinearSolverDense and lambda 1e0 / 1e-2
//
// Two classes, one C ABI:
//   gpba::GpBaLevenberg   : g2o::OptimizationAlgorithm   (seam A', Thirdparty/g2o/g2o/core/optimization_algorithm.h:46-110)
//       solve(iteration 0) runs the WHOLE optimize() on the device (gpba_optimize) and writes the estimates back
//       with setEstimate(); later iterations return Terminate.  This is the fast path: residuals, Jacobians, Schur,
//       reduced solve, back-substitution and the LM accept/reject all stay on the GPU.
//   gpba::GpBaBlockSolver : g2o::BlockSolverBase         (seam B, block_solver.h:83-91, solver.h:43-148)
//       drop-in for g2o::BlockSolverX under the stock OptimizationAlgorithmLevenberg: buildStructure / buildSystem /
//       setLambda / solve / restoreDiagonal map 1:1 onto the L1 entry points; x() and b() are filled so that
//       computeScale() and SparseOptimizer::update() work unchanged (optimization_algorithm_levenberg.cpp:115,187-194).
//       g2o still evaluates residuals and applies oplus on its CPU objects, so this path only moves the linear
//       algebra; it exists so that the g2o::Solver contract named in the north star is honoured literally.
//
// Flattening rules (SURVEY.md §8b "Flattening"): vertices in ascending id (= Hessian order,
// sparse_optimizer.cpp:166-190), edges in activeEdges() order (= internalId order, sparse_optimizer.cpp:482-487).
#pragma once

#include <map>
#include <vector>
#include <limits>
#include <cstring>
#include <iostream>

#include "Thirdparty/g2o/g2o/core/block_solver.h"
#include "Thirdparty/g2o/g2o/core/optimization_algorithm.h"
#include "Thirdparty/g2o/g2o/core/sparse_optimizer.h"
#include "Thirdparty/g2o/g2o/core/robust_kernel_impl.h"
#include "Thirdparty/g2o/g2o/types/types_sba.h"
#include "G2oTypes.h"       // VertexPoseVel, VertexExtrinsic, EdgeMonoGP, EdgeMonoGPExtrinsic, EdgeStereoGP, EdgeMono, EdgeStereo, EdgeGaussianPrior, EdgeVelocity
#include "KeyFrame.h"       // MultiKeyFrame::mTbc
#include "gpba.h"

namespace gpba {

// Flattened copy of the active graph in the layout of gpba_problem (include/gpba.h).
struct FlatGraph {
  std::vector<double> cam_intr, cam_Tbc, kf_pose, kf_vel, kf_time, pt_xyz, rec_t, obs_u, obs_v, obs_ur, obs_w;
  std::vector<uint8_t> kf_fixed, obs_flags;
  std::vector<int32_t> rec_kf1, rec_kf2, rec_cam, obs_rec, obs_pt, prior_kf1, prior_kf2, velp_kf;
  std::vector<ORB_SLAM3::VertexPoseVel*> kf_vertex;
  std::vector<g2o::VertexSBAPointXYZ*> pt_vertex;
  std::vector<g2o::OptimizableGraph::Edge*> obs_edge;   // reprojection edge of every observation (write-back of chi2)
  double qc[6], bf = 0, huber_mono = 0, huber_stereo = 0, huber_prior = 0;
  // extrinsic self-calibration: VertexExtrinsic of camera c (from the EdgeMonoGPExtrinsic that reference it), which of them
  // are un-fixed, and the EdgeExtrinsicPrior edges (R_ini, information)
  std::map<int, ORB_SLAM3::VertexExtrinsic*> ext_vertex, ext_vertex_by_id;
  std::map<int, uint8_t> ext_free;
  std::vector<ORB_SLAM3::EdgeExtrinsicPrior*> ext_priors;
  std::vector<double> ext_prior_R, ext_prior_info;
  bool have_ext_prior = false;
  bool ok = true;
  std::string why;

  static void put_se3(const Sophus::SE3d& T, std::vector<double>& out) {
    const Eigen::Quaterniond& q = T.unit_quaternion();
    out.push_back(q.x()); out.push_back(q.y()); out.push_back(q.z()); out.push_back(q.w());
    out.push_back(T.translation()(0)); out.push_back(T.translation()(1)); out.push_back(T.translation()(2));
  }

  // kernel delta of an edge: RobustKernelHuber::delta() (robust_kernel.h:72) or 0 when setRobustKernel(0)
  static double delta_of(const g2o::OptimizableGraph::Edge* e) { return e->robustKernel() ? e->robustKernel()->delta() : 0.0; }

  int record(std::map<std::pair<std::pair<int, int>, std::pair<int, long long> >, int>& recs, int k1, int k2, int cam, double t) {
    long long tb; std::memcpy(&tb, &t, sizeof(t));
    const auto key = std::make_pair(std::make_pair(k1, k2), std::make_pair(cam, tb));
    auto it = recs.find(key);
    if (it != recs.end()) return it->second;
    const int r = (int)rec_kf1.size();
    rec_kf1.push_back(k1); rec_kf2.push_back(k2); rec_cam.push_back(cam); rec_t.push_back(t);
    recs[key] = r;
    return r;
  }

  // Walks optimizer->vertices() / activeEdges() exactly once.  Extrinsic vertices are fixed during LocalGPBA's first
  // optimize() (src/Optimizer.cc:983-988); in the bExtrinsic second stage (:1228-1240) the un-fixed ones, with their
  // EdgeExtrinsicPrior edges, travel through gpba_set_extrinsics (see apply_extrinsics below).
  void build(g2o::SparseOptimizer* opt) {
    using namespace ORB_SLAM3;
    std::map<int, int> kf_index, pt_index;   // vertex id -> flat index, ascending id == Hessian order
    std::vector<std::pair<int, g2o::OptimizableGraph::Vertex*> > vs;
    for (auto& kv : opt->vertices()) vs.push_back(std::make_pair(kv.first, static_cast<g2o::OptimizableGraph::Vertex*>(kv.second)));
    std::sort(vs.begin(), vs.end(), [](const std::pair<int, g2o::OptimizableGraph::Vertex*>& a, const std::pair<int, g2o::OptimizableGraph::Vertex*>& b) { return a.first < b.first; });
    std::vector<Sophus::SE3d> Tbc(MultiKeyFrame::mTbc.size());
    for (size_t c = 0; c < Tbc.size(); ++c) Tbc[c] = MultiKeyFrame::mTbc[c].cast<double>();   // G2oTypes.cc:330
    for (auto& iv : vs) {
      if (VertexPoseVel* v = dynamic_cast<VertexPoseVel*>(iv.second)) {
        kf_index[iv.first] = (int)kf_vertex.size();
        kf_vertex.push_back(v);
        const PoseVelocity& pv = v->estimate();
        put_se3(pv.Twb, kf_pose);
        for (int i = 0; i < 6; ++i) kf_vel.push_back(pv.Vel(i));
        kf_time.push_back(pv.time);
        kf_fixed.push_back(v->fixed() ? 1 : 0);
        bf = pv.bf;
        if (cam_intr.empty())
          for (GeometricCamera* cam : pv.vpCameras) for (int i = 0; i < 4; ++i) cam_intr.push_back((double)cam->getParameter(i));   // float intrinsics, GeometricCamera.h:81
      } else if (g2o::VertexSBAPointXYZ* p = dynamic_cast<g2o::VertexSBAPointXYZ*>(iv.second)) {
        pt_index[iv.first] = (int)pt_vertex.size();
        pt_vertex.push_back(p);
        for (int i = 0; i < 3; ++i) pt_xyz.push_back(p->estimate()(i));
      } else if (VertexExtrinsic* x = dynamic_cast<VertexExtrinsic*>(iv.second)) {
        ext_vertex_by_id[iv.first] = x;   // camera index comes from the edges that reference it
      }
    }
    std::map<std::pair<std::pair<int, int>, std::pair<int, long long> >, int> recs;
    bool have_qc = false;
    auto take_gp = [&](const GaussianProcess* gp) {
      if (have_qc || !gp) return;
      for (int i = 0; i < 6; ++i) qc[i] = gp->mQc(i, i);
      have_qc = true;
    };
    auto add_obs = [&](g2o::OptimizableGraph::Edge* e, int rec, int pt, double u, double v, double ur, double w) {
      obs_edge.push_back(e); obs_rec.push_back(rec); obs_pt.push_back(pt);
      obs_u.push_back(u); obs_v.push_back(v); obs_ur.push_back(ur); obs_w.push_back(w);
      uint8_t f = 0;
      if (e->level() != 0) f |= GPBA_OBS_LEVEL1;
      if (!e->robustKernel()) f |= GPBA_OBS_NO_KERNEL;
      obs_flags.push_back(f);   // GPBA_OBS_CLOSE is filled by the caller from MapPoint::mvTrackDepth (Optimizer.cc:1273)
    };
    const int n_cam = (int)Tbc.size();
    // every edge of the graph, in insertion order, including level-1 edges (they widen the Hschur pattern,
    // block_solver.hpp:262-288)
    std::vector<g2o::OptimizableGraph::Edge*> edges;
    for (auto* he : opt->edges()) edges.push_back(static_cast<g2o::OptimizableGraph::Edge*>(he));
    std::sort(edges.begin(), edges.end(), [](const g2o::OptimizableGraph::Edge* a, const g2o::OptimizableGraph::Edge* b) { return a->internalId() < b->internalId(); });
    for (g2o::OptimizableGraph::Edge* e : edges) {
      if (EdgeMonoGP* g = dynamic_cast<EdgeMonoGP*>(e)) {
        take_gp(g->gp);
        const int r = record(recs, kf_index[g->vertices()[0]->id()], kf_index[g->vertices()[1]->id()], g->cam_idx, g->t);
        add_obs(e, r, pt_index[g->vertices()[2]->id()], g->measurement()(0), g->measurement()(1), -1.0, g->information()(0, 0));
        huber_mono = delta_of(e) > 0 ? delta_of(e) : huber_mono;
      } else if (EdgeMonoGPExtrinsic* g = dynamic_cast<EdgeMonoGPExtrinsic*>(e)) {
        take_gp(g->gp);
        const int r = record(recs, kf_index[g->vertices()[0]->id()], kf_index[g->vertices()[1]->id()], g->cam_idx, g->t);
        add_obs(e, r, pt_index[g->vertices()[2]->id()], g->measurement()(0), g->measurement()(1), -1.0, g->information()(0, 0));
        VertexExtrinsic* xv = static_cast<VertexExtrinsic*>(g->vertices()[3]);
        Tbc[g->cam_idx] = xv->estimate();   // the extrinsic vertex wins over mTbc
        ext_vertex[g->cam_idx] = xv;
        if (!xv->fixed()) ext_free[g->cam_idx] = 1;
        huber_mono = delta_of(e) > 0 ? delta_of(e) : huber_mono;
      } else if (EdgeStereoGP* g = dynamic_cast<EdgeStereoGP*>(e)) {
        take_gp(g->gp);
        const int r = record(recs, kf_index[g->vertices()[0]->id()], kf_index[g->vertices()[1]->id()], g->cam_idx, g->t);
        add_obs(e, r, pt_index[g->vertices()[2]->id()], g->measurement()(0), g->measurement()(1), g->measurement()(2), g->information()(0, 0));
        huber_stereo = delta_of(e) > 0 ? delta_of(e) : huber_stereo;
      } else if (EdgeMono* m = dynamic_cast<EdgeMono*>(e)) {
        const int k = kf_index[m->vertices()[0]->id()];
        const int r = record(recs, -1, k, n_cam - 1, kf_time[k]);   // synchronous reference camera = mTbc.back() (G2oTypes.cc:49)
        add_obs(e, r, pt_index[m->vertices()[1]->id()], m->measurement()(0), m->measurement()(1), -1.0, m->information()(0, 0));
        huber_mono = delta_of(e) > 0 ? delta_of(e) : huber_mono;
      } else if (EdgeStereo* s = dynamic_cast<EdgeStereo*>(e)) {
        const int k = kf_index[s->vertices()[0]->id()];
        const int r = record(recs, -1, k, n_cam - 1, kf_time[k]);
        add_obs(e, r, pt_index[s->vertices()[1]->id()], s->measurement()(0), s->measurement()(1), s->measurement()(2), s->information()(0, 0));
        huber_stereo = delta_of(e) > 0 ? delta_of(e) : huber_stereo;
      } else if (EdgeGaussianPrior* p = dynamic_cast<EdgeGaussianPrior*>(e)) {
        prior_kf1.push_back(kf_index[p->vertices()[0]->id()]); prior_kf2.push_back(kf_index[p->vertices()[1]->id()]);
        huber_prior = delta_of(e);   // 21.026 in BundleAdjustment (Optimizer.cc:128-130), none in LocalGPBA (:903-910)
      } else if (dynamic_cast<EdgeVelocity*>(e)) {
        velp_kf.push_back(kf_index[e->vertices()[0]->id()]);
      } else if (EdgeExtrinsicPrior* xp = dynamic_cast<EdgeExtrinsicPrior*>(e)) {
        // inactive while the extrinsic vertex is fixed (allVerticesFixed, sparse_optimizer.cpp:236-247); kept for the second stage
        ext_priors.push_back(xp);
      } else {
        ok = false; why = "edge type outside the GP-BA path";
      }
    }
    for (int c = 0; c < n_cam; ++c) put_se3(Tbc[c], cam_Tbc);
    if (!have_qc) for (int i = 0; i < 6; ++i) qc[i] = 1.0;
    // EdgeExtrinsicPrior -> (camera, R_ini, information): R_ is stored inverted (G2oTypes.h:474)
    ext_prior_R.assign(4 * (size_t)n_cam, 0.0); ext_prior_info.assign(9 * (size_t)n_cam, 0.0);
    for (int c = 0; c < n_cam; ++c) ext_prior_R[4 * c + 3] = 1.0;
    for (EdgeExtrinsicPrior* xp : ext_priors)
      for (auto& kv : ext_vertex)
        if (kv.second == xp->vertices()[0]) {
          const Eigen::Quaterniond q = xp->R_.inverse().unit_quaternion();
          const int c = kv.first;
          ext_prior_R[4 * c] = q.x(); ext_prior_R[4 * c + 1] = q.y(); ext_prior_R[4 * c + 2] = q.z(); ext_prior_R[4 * c + 3] = q.w();
          for (int r = 0; r < 3; ++r) for (int k = 0; k < 3; ++k) ext_prior_info[9 * c + 3 * r + k] = xp->information()(r, k);
          have_ext_prior = true;
        }
  }

  // un-fixed VertexExtrinsic + EdgeExtrinsicPrior -> the handle (no-op when every extrinsic is fixed)
  int apply_extrinsics(gpba_handle* h) const {
    bool any = false;
    for (auto& kv : ext_free) any = any || kv.second;
    if (!any) return GPBA_OK;
    const int n_cam = (int)(cam_Tbc.size() / 7);
    std::vector<uint8_t> fr((size_t)n_cam, 0);
    for (auto& kv : ext_free) fr[kv.first] = kv.second;
    gpba_extrinsics e;
    e.free_mask = fr.data();
    e.prior_R = have_ext_prior ? ext_prior_R.data() : nullptr;
    e.prior_info = have_ext_prior ? ext_prior_info.data() : nullptr;
    return gpba_set_extrinsics(h, &e);
  }

  gpba_problem view(double lambda_init, int linear_solver) const {
    gpba_problem P;
    std::memset(&P, 0, sizeof(P));
    P.n_cam = (int32_t)(cam_Tbc.size() / 7); P.cam_intr = cam_intr.data(); P.cam_Tbc = cam_Tbc.data(); P.bf = bf;
    P.n_kf = (int32_t)kf_time.size(); P.kf_pose = kf_pose.data(); P.kf_vel = kf_vel.data(); P.kf_time = kf_time.data(); P.kf_fixed = kf_fixed.data();
    P.n_pt = (int32_t)pt_vertex.size(); P.pt_xyz = pt_xyz.data();
    P.n_rec = (int32_t)rec_kf1.size(); P.rec_kf1 = rec_kf1.data(); P.rec_kf2 = rec_kf2.data(); P.rec_cam = rec_cam.data(); P.rec_t = rec_t.data();
    P.n_obs = (int64_t)obs_u.size(); P.obs_u = obs_u.data(); P.obs_v = obs_v.data(); P.obs_ur = obs_ur.data(); P.obs_inv_sigma2 = obs_w.data();
    P.obs_rec = obs_rec.data(); P.obs_pt = obs_pt.data(); P.obs_flags = obs_flags.data();
    P.n_prior = (int32_t)prior_kf1.size(); P.prior_kf1 = prior_kf1.data(); P.prior_kf2 = prior_kf2.data();
    P.n_velp = (int32_t)velp_kf.size(); P.velp_kf = velp_kf.data();
    for (int i = 0; i < 6; ++i) P.qc[i] = qc[i];
    P.huber_mono = huber_mono; P.huber_stereo = huber_stereo; P.huber_prior = huber_prior;
    P.lambda_init = lambda_init; P.linear_solver = linear_solver;
    return P;
  }

  // The reference leaves every active edge with the _error of the LAST EVALUATED state: optimize() never re-evaluates after
  // its last trial, so after a rejected trial e->chi2() and activeRobustChi2() (Optimizer.cc:1254, 1273-1330) see the
  // rejected state's errors while the vertices hold the accepted estimate.  Reproduced here: the reprojection errors come
  // from gpba_edge_errors; the few prior edges are re-evaluated by g2o itself at the last evaluated keyframe states, which
  // are set and then replaced by the estimate again.
  void write_back_errors(gpba_handle* h, g2o::SparseOptimizer* opt) {
    std::vector<double> err(3 * obs_edge.size());
    if (gpba_edge_errors(h, err.data()) == GPBA_OK)
      for (size_t i = 0; i < obs_edge.size(); ++i) {
        if (err[3 * i] != err[3 * i]) continue;                     // NaN: inactive edge, keeps its own error
        double* ed = obs_edge[i]->errorData();
        for (int d = 0; d < obs_edge[i]->dimension(); ++d) ed[d] = err[3 * i + d];
      }
    std::vector<double> kp(kf_pose.size()), kv(kf_vel.size()), tb(cam_Tbc.size());
    if (gpba_download_evaluated_state(h, kp.data(), kv.data(), tb.data()) != GPBA_OK) return;
    std::vector<ORB_SLAM3::PoseVelocity> keep(kf_vertex.size());
    for (size_t k = 0; k < kf_vertex.size(); ++k) {
      keep[k] = kf_vertex[k]->estimate();
      if (kf_vertex[k]->fixed()) continue;
      ORB_SLAM3::PoseVelocity pv = keep[k];
      const double* q = &kp[7 * k];
      pv.Twb = Sophus::SE3d(Eigen::Quaterniond(q[3], q[0], q[1], q[2]), Eigen::Vector3d(q[4], q[5], q[6]));
      for (int i = 0; i < 6; ++i) pv.Vel(i) = kv[6 * k + i];
      kf_vertex[k]->setEstimate(pv);
    }
    std::map<int, Sophus::SE3d> keep_ext;
    for (auto& kvp : ext_vertex) {
      keep_ext[kvp.first] = kvp.second->estimate();
      if (kvp.second->fixed()) continue;
      const double* q = &tb[7 * kvp.first];
      kvp.second->setEstimate(Sophus::SE3d(Eigen::Quaterniond(q[3], q[0], q[1], q[2]), Eigen::Vector3d(q[4], q[5], q[6])));
    }
    for (g2o::OptimizableGraph::Edge* e : opt->activeEdges())
      if (dynamic_cast<ORB_SLAM3::EdgeGaussianPrior*>(e) || dynamic_cast<ORB_SLAM3::EdgeVelocity*>(e) || dynamic_cast<ORB_SLAM3::EdgeExtrinsicPrior*>(e))
        e->computeError();
    for (size_t k = 0; k < kf_vertex.size(); ++k) kf_vertex[k]->setEstimate(keep[k]);
    for (auto& kvp : ext_vertex) kvp.second->setEstimate(keep_ext[kvp.first]);
  }

  // vertex->setEstimate() for every free vertex (what Optimizer.cc:324-366 / 1360-1430 read afterwards)
  void write_back(gpba_handle* h) {
    std::vector<double> kp(kf_pose.size()), kv(kf_vel.size()), pt(pt_xyz.size());
    if (gpba_download_state(h, kp.data(), kv.data(), pt.data()) != GPBA_OK) return;
    for (size_t k = 0; k < kf_vertex.size(); ++k) {
      if (kf_vertex[k]->fixed()) continue;
      ORB_SLAM3::PoseVelocity pv = kf_vertex[k]->estimate();
      const double* q = &kp[7 * k];
      pv.Twb = Sophus::SE3d(Eigen::Quaterniond(q[3], q[0], q[1], q[2]), Eigen::Vector3d(q[4], q[5], q[6]));
      for (int i = 0; i < 6; ++i) pv.Vel(i) = kv[6 * k + i];
      kf_vertex[k]->setEstimate(pv);
    }
    for (size_t p = 0; p < pt_vertex.size(); ++p) pt_vertex[p]->setEstimate(Eigen::Vector3d(pt[3 * p], pt[3 * p + 1], pt[3 * p + 2]));
    // free extrinsics: VertexExtrinsic::setEstimate (Optimizer.cc:1419-1428 reads them back into MultiKeyFrame::mTbc)
    std::vector<double> tbc(cam_Tbc.size());
    if (!ext_vertex.empty() && gpba_get_extrinsics(h, tbc.data()) == GPBA_OK)
      for (auto& kv : ext_vertex) {
        if (kv.second->fixed()) continue;
        const double* q = &tbc[7 * kv.first];
        kv.second->setEstimate(Sophus::SE3d(Eigen::Quaterniond(q[3], q[0], q[1], q[2]), Eigen::Vector3d(q[4], q[5], q[6])));
      }
  }
};

// ------------------------------------------------------------------------------------------------ seam A'
// Replaces `new g2o::OptimizationAlgorithmLevenberg(new g2o::BlockSolverX(linearSolver))`.
class GpBaLevenberg : public g2o::OptimizationAlgorithm {
 public:
  // linear_solver: GPBA_SOLVER_DENSE_CHOL where the reference uses LinearSolverDense (LocalGPBA) or wants the direct
  // factorization, GPBA_SOLVER_SPARSE_CHOL where it uses LinearSolverEigen (BundleAdjustment; the device runs the same tile
  // Cholesky for both).  GPBA_SOLVER_PCG exists but is not recommended.
  explicit GpBaLevenberg(int linear_solver = GPBA_SOLVER_DENSE_CHOL, int device = -1)
      : _h(nullptr), _linear(linear_solver), _device(device), _lambda_init(-1.0), _iters(10), _done(false) { std::memset(&_trace, 0, sizeof(_trace)); }
  virtual ~GpBaLevenberg() { gpba_destroy(_h); }
  void setUserLambdaInit(double l) { _lambda_init = l; }          // OptimizationAlgorithmLevenberg::setUserLambdaInit
  void setMaxIterations(int n) { _iters = n; }                      // = the argument of optimizer.optimize(n)
  const gpba_lm_trace& trace() const { return _trace; }             // per-iteration chi2 / lambda / trials (G2OBatchStatistics analogue)

  virtual bool init(bool /*online*/ = false) { _done = false; return true; }

  virtual SolverResult solve(int iteration, bool /*online*/ = false) {
    if (iteration > 0 || _done) return Terminate;                  // the whole LM loop ran in iteration 0
    FlatGraph G;
    G.build(_optimizer);
    if (!G.ok) { std::cerr << "GpBaLevenberg: " << G.why << std::endl; return Fail; }
    const gpba_problem P = G.view(_lambda_init, _linear);
    gpba_destroy(_h); _h = nullptr;
    if (gpba_create(&P, _device, &_h) != GPBA_OK) { std::cerr << "gpba_create: " << gpba_last_error() << std::endl; return Fail; }
    if (G.apply_extrinsics(_h) != GPBA_OK) { std::cerr << "gpba_set_extrinsics: " << gpba_last_error() << std::endl; return Fail; }
    // setForceStopFlag(bool*) (sparse_optimizer.h:188): bool and unsigned char share size and representation here
    const volatile unsigned char* stop = reinterpret_cast<const volatile unsigned char*>(_optimizer->forceStopFlag());
    if (gpba_optimize(_h, _iters, stop, nullptr, &_trace) != GPBA_OK) { std::cerr << "gpba_optimize: " << gpba_last_error() << std::endl; return Fail; }
    G.write_back(_h);
    G.write_back_errors(_h, _optimizer);                            // edge->chi2() / activeRobustChi2() for the caller (Optimizer.cc:1254, 1263-1348)
    _done = true;
    return _trace.result == GPBA_FAIL ? Fail : Terminate;
  }

  virtual bool computeMarginals(g2o::SparseBlockMatrix<g2o::MatrixXd>&, const std::vector<std::pair<int, int> >&) { return false; }   // unused by AMC-SLAM
  virtual bool updateStructure(const std::vector<g2o::HyperGraph::Vertex*>&, const g2o::HyperGraph::EdgeSet&) { return false; }
  virtual void printVerbose(std::ostream& os) const {
    os << "\t gpba iterations= " << _trace.n_iters << "\t trials= " << _trace.total_trials;
    if (_trace.n_iters > 0) os << "\t chi2= " << _trace.chi2_after[_trace.n_iters - 1] << "\t lambda= " << _trace.lambda[_trace.n_iters - 1];
  }

 private:
  gpba_handle* _h;
  int _linear, _device;
  double _lambda_init;
  int _iters;
  bool _done;
  gpba_lm_trace _trace;
};

// ------------------------------------------------------------------------------------------------ seam B
// Replaces `new g2o::BlockSolverX(linearSolver)` under the stock g2o::OptimizationAlgorithmLevenberg.
class GpBaBlockSolver : public g2o::BlockSolverBase {
 public:
  explicit GpBaBlockSolver(int linear_solver = GPBA_SOLVER_DENSE_CHOL, int device = -1)
      : _h(nullptr), _linear(linear_solver), _device(device), _schur(true), _writeDebug(false) {}
  virtual ~GpBaBlockSolver() { gpba_destroy(_h); }

  virtual bool init(g2o::SparseOptimizer* optimizer, bool /*online*/ = false) { _optimizer = optimizer; return true; }   // block_solver.hpp:489-499

  virtual bool buildStructure(bool /*zeroBlocks*/ = false) {      // block_solver.hpp:142-295
    _G = FlatGraph();
    _G.build(_optimizer);
    if (!_G.ok) return false;
    // Seam B exchanges Solver::_x / _b in g2o's layout, where a VertexExtrinsic has 6 entries; the library keeps a 12-slot
    // per extrinsic (include/gpba.h).  The second stage of LocalGPBA (free extrinsics) goes through seam A' (GpBaLevenberg).
    for (auto& kv : _G.ext_free) if (kv.second) return false;
    const gpba_problem P = _G.view(0.0, _linear);
    gpba_destroy(_h); _h = nullptr;
    if (gpba_create(&P, _device, &_h) != GPBA_OK) return false;
    gpba_structure_info info;
    if (gpba_build_structure(_h, &info) != GPBA_OK) return false;
    int64_t n = 0;
    gpba_vector_size(_h, &n);
    resizeVector((size_t)n);                                        // Solver::_x / _b (solver.cpp:46-72)
    return true;
  }
  virtual bool updateStructure(const std::vector<g2o::HyperGraph::Vertex*>&, const g2o::HyperGraph::EdgeSet&) { return false; }   // online mode unused (block_solver.hpp:313-316 aborts too)

  // g2o has just changed the vertex estimates on the CPU (update / pop): mirror them, then linearise on the device.
  virtual bool buildSystem() {                                      // block_solver.hpp:502-560
    push_estimates();
    double chi2;
    if (gpba_compute_errors(_h, &chi2) != GPBA_OK || gpba_build_system(_h) != GPBA_OK) return false;
    gpba_get_b(_h, _b);
    return false;   // the reference returns 0 on success and the caller ignores it (block_solver.hpp:559)
  }
  virtual bool setLambda(double lambda, bool backup = false) { return gpba_set_lambda(_h, lambda, backup) == GPBA_OK; }   // :563-589
  virtual void restoreDiagonal() { gpba_restore_diagonal(_h); }                                                        // :591-604
  virtual bool solve() {                                            // :353-486
    int ok = 0;
    if (gpba_solve(_h, &ok) != GPBA_OK) return false;
    gpba_get_x(_h, _x);
    return ok != 0;
  }
  virtual bool computeMarginals(g2o::SparseBlockMatrix<g2o::MatrixXd>&, const std::vector<std::pair<int, int> >&) { return false; }
  virtual bool supportsSchur() { return true; }
  virtual bool schur() { return _schur; }
  virtual void setSchur(bool s) { _schur = s; }
  virtual void setWriteDebug(bool b) { _writeDebug = b; }
  virtual bool writeDebug() const { return _writeDebug; }
  virtual bool saveHessian(const std::string&) const { return false; }
  virtual void multiplyHessian(double*, const double*) const {}     // only used by g2o's PCG / dogleg, not on this path

 private:
  void push_estimates() {
    FlatGraph cur;
    cur.build(_optimizer);   // cheap relative to the CPU-side residual evaluation this seam keeps on g2o
    gpba_reset_state(_h, cur.kf_pose.data(), cur.kf_vel.data(), cur.pt_xyz.data());
  }
  gpba_handle* _h;
  int _linear, _device;
  bool _schur, _writeDebug;
  FlatGraph _G;
};

}  // namespace gpba
