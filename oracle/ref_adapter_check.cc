// TEST INFRASTRUCTURE, not product code.  Compiles the reference-side binding adapter/g2o_gpba_solver.h -- the code a
// maintainer adds to the AMC-SLAM tree -- against the reference's REAL g2o headers and G2oTypes.h (stand-in Eigen / Sophus,
// oracle/ref_shim/), which the repository could not do before, and runs its graph flattening on a real g2o graph:
//   gpba_problem P -> the reference's graph (BaGraph, ref_g2o_graph.h) -> initializeOptimization(0)
//                  -> gpba::FlatGraph::build (the adapter) -> gpba_problem Q -> the ORACLE's optimize on Q.
// tests/test_whole_path_reference.py requires the oracle's run on Q to be the oracle's run on P: then the adapter hands the C ABI what the
// graph holds.  (The C ABI itself is exercised on the device by the GPU suite; here libgpba.so is only linked so that the
// header's calls resolve.)  Built by `make -C oracle _ref` into oracle/_ref/libadapter_check.so.
#define REF_G2O_DEFINE_STATICS
#include "ref_g2o_graph.h"
#include "../adapter/g2o_gpba_solver.h"

extern "C" {   // oracle/gpba_oracle.cc
void* oracle_create(const gpba_problem* p);
void oracle_destroy(void* h);
int oracle_optimize(void* h, int iters, const volatile unsigned char* stop, const gpba_lm_params* P, gpba_lm_trace* tr);
int oracle_download_state(void* h, double* kf_pose, double* kf_vel, double* pt_xyz);
int oracle_edge_chi2(void* h, double* chi2);

// counts[0..5] = n_kf, n_pt, n_rec, n_obs, n_prior, n_velp of the flattened problem; the state and per-observation chi2
// come back in the ADAPTER's order (ascending vertex id, activeEdges() order), which for a BaGraph is P's own order.
// Returns 0, or 1 when FlatGraph::build refused the graph.
int ref_adapter_roundtrip(const gpba_problem* P, int iters, int64_t* counts, double* kf_pose_out, double* kf_vel_out, double* pt_out,
                          double* edge_chi2_out, gpba_lm_trace* tr) {
  BaGraph G(P, 0);
  G.optimizer.initializeOptimization(0);
  gpba::FlatGraph F;
  F.build(&G.optimizer);
  if (!F.ok) { std::cerr << "FlatGraph: " << F.why << std::endl; return 1; }
  const gpba_problem Q = F.view(P->lambda_init, P->linear_solver);
  counts[0] = Q.n_kf; counts[1] = Q.n_pt; counts[2] = Q.n_rec; counts[3] = Q.n_obs; counts[4] = Q.n_prior; counts[5] = Q.n_velp;
  void* o = oracle_create(&Q);
  oracle_optimize(o, iters, nullptr, nullptr, tr);
  oracle_download_state(o, kf_pose_out, kf_vel_out, pt_out);
  if (edge_chi2_out) oracle_edge_chi2(o, edge_chi2_out);
  oracle_destroy(o);
  return 0;
}

// The two adapter classes inside the real SparseOptimizer without a device: GpBaLevenberg::solve must fail cleanly
// (gpba_create reports GPBA_ERR_NO_DEVICE) and leave the graph untouched.  Returns optimize()'s return value.
int ref_adapter_no_device(const gpba_problem* P, double* kf_pose_out) {
  BaGraph G(P, 0);
  G.optimizer.setAlgorithm(new gpba::GpBaLevenberg(P->linear_solver, 0));   // replaces (and deletes) the stock algorithm
  G.solver = nullptr;
  G.optimizer.initializeOptimization(0);
  const int n = G.optimizer.optimize(5);
  G.read_back(P, kf_pose_out, nullptr, nullptr, nullptr);
  return n;
}

// The drop-in itself: gpba::GpBaLevenberg as the algorithm of the reference's real SparseOptimizer, on a CUDA device.
// optimize(iters) flattens the graph, runs gpba_optimize on the device and writes estimates and stale errors back into the
// g2o objects; the outputs are read from those objects the way Optimizer.cc reads them (:338-367, 1263-1348).
// Returns optimize()'s return value (0 = the adapter reported Fail).
int ref_adapter_optimize(const gpba_problem* P, int iters, int device, double* kf_pose_out, double* kf_vel_out, double* pt_out,
                         double* edge_chi2_out, gpba_lm_trace* tr) {
  BaGraph G(P, 0);
  gpba::GpBaLevenberg* alg = new gpba::GpBaLevenberg(P->linear_solver, device);
  if (P->lambda_init > 0) alg->setUserLambdaInit(P->lambda_init);
  alg->setMaxIterations(iters);
  G.optimizer.setAlgorithm(alg);   // replaces (and deletes) the stock algorithm
  G.solver = nullptr;
  G.optimizer.initializeOptimization(0);
  const int n = G.optimizer.optimize(iters);
  G.read_back(P, kf_pose_out, kf_vel_out, pt_out, edge_chi2_out);
  if (tr) *tr = alg->trace();
  return n;
}

// Seam B: gpba::GpBaBlockSolver as the g2o::Solver of the reference's STOCK OptimizationAlgorithmLevenberg -- g2o evaluates the
// residuals and applies oplus on its own objects, the level-1 entry points do the linear algebra.
int ref_adapter_block_solver(const gpba_problem* P, int iters, int device, double* kf_pose_out, double* kf_vel_out, double* pt_out,
                             double* edge_chi2_out, gpba_lm_trace* tr) {
  BaGraph G(P, 0);
  gpba::GpBaBlockSolver* bs = new gpba::GpBaBlockSolver(P->linear_solver, device);
  g2o::OptimizationAlgorithmLevenberg* alg = new g2o::OptimizationAlgorithmLevenberg(bs);
  if (P->lambda_init > 0) alg->setUserLambdaInit(P->lambda_init);
  G.optimizer.setAlgorithm(alg);
  G.solver = alg;
  if (tr) { std::memset(tr, 0, sizeof(*tr)); tr->result = GPBA_RESULT_OK; }
  Recorder rec;
  rec.opt = &G.optimizer; rec.alg = alg; rec.tr = tr;
  G.optimizer.addPostIterationAction(&rec);
  G.optimizer.initializeOptimization(0);
  if (tr) { G.optimizer.computeActiveErrors(); tr->chi2_before[0] = G.optimizer.activeRobustChi2(); }
  const int n = G.optimizer.optimize(iters);
  if (tr) tr->n_iters = n;
  G.optimizer.removePostIterationAction(&rec);
  G.read_back(P, kf_pose_out, kf_vel_out, pt_out, edge_chi2_out);
  return n;
}

// LocalGPBA's two stages through the binding (src/Optimizer.cc:1219-1240): the graph holds VertexExtrinsic / EdgeMonoGPExtrinsic /
// EdgeExtrinsicPrior; stage 1 with the extrinsics fixed, stage 2 with those of `ext_free` released -- each stage one
// optimize() of gpba::GpBaLevenberg, which passes the released extrinsics and their priors through gpba_set_extrinsics and
// writes the calibrated extrinsics back into the VertexExtrinsic objects.
int ref_adapter_ext(const gpba_problem* P, const uint8_t* ext_free, const double* prior_q, const double* prior_info, int it1, int it2,
                    int device, double* kf_pose_out, double* kf_vel_out, double* pt_out, double* Tbc_out, gpba_lm_trace* tr1,
                    gpba_lm_trace* tr2) {
  ExtGraph G(P, prior_q, prior_info);
  gpba::GpBaLevenberg* alg = new gpba::GpBaLevenberg(P->linear_solver, device);
  if (P->lambda_init > 0) alg->setUserLambdaInit(P->lambda_init);
  G.optimizer.setAlgorithm(alg);
  G.solver = nullptr;
  alg->setMaxIterations(it1);
  G.optimizer.initializeOptimization();
  G.optimizer.computeActiveErrors();
  int n = G.optimizer.optimize(it1);
  if (tr1) *tr1 = alg->trace();
  for (int c = 0; c < P->n_cam - 1; ++c) if (ext_free[c]) G.vext[c]->setFixed(false);
  alg->setMaxIterations(it2);
  G.optimizer.initializeOptimization();
  G.optimizer.computeActiveErrors();
  n = G.optimizer.optimize(it2);
  if (tr2) *tr2 = alg->trace();
  G.read_back(P, kf_pose_out, kf_vel_out, pt_out, Tbc_out);
  return n;
}

}  // extern "C"
