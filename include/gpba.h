/*
 * gpba.h -- C ABI of libgpba.so, the B200-native GP-interpolated bundle-adjustment
 * solver that replaces what runs inside g2o::SparseOptimizer::optimize for AMC-SLAM's
 * Optimizer::LocalGPBA / Optimizer::BundleAdjustment.
 *
 * Every entry point names the reference interface it replaces (paths relative to the
 * AMC-SLAM tree).  All arrays are caller-owned HOST pointers, SoA, f64 / int32 / uint8.
 * All functions return 0 on success and a negative gpba_status on error; no exception
 * ever crosses this boundary.  One handle is used by one host thread (same rule as one
 * g2o::SparseOptimizer instance, Thirdparty/g2o/g2o/core/sparse_optimizer.h).
 *
 * Two levels share one handle:
 *   L1  Solver-shaped   mirrors g2o::Solver / g2o::BlockSolver<Traits>
 *                       (Thirdparty/g2o/g2o/core/solver.h:43-148, block_solver.h:96-176)
 *   L2  whole optimize  mirrors g2o::SparseOptimizer::optimize + OptimizationAlgorithmLevenberg::solve
 *                       (sparse_optimizer.cpp:354-419, optimization_algorithm_levenberg.cpp:61-169)
 */
#ifndef GPBA_H
#define GPBA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GPBA_MAX_ITERS 64

typedef enum gpba_status {
  GPBA_OK = 0,
  GPBA_ERR_INVALID = -1,   /* bad argument / inconsistent problem            */
  GPBA_ERR_CUDA = -2,      /* CUDA runtime error (message via gpba_last_error) */
  GPBA_ERR_NO_DEVICE = -3, /* no CUDA device: there is NO CPU fallback        */
  GPBA_ERR_STATE = -4,     /* call order violated (e.g. solve before build)   */
  GPBA_ERR_NCCL = -5
} gpba_status;

/* g2o::OptimizationAlgorithm::SolverResult (optimization_algorithm.h) */
typedef enum gpba_solver_result { GPBA_TERMINATE = 2, GPBA_RESULT_OK = 1, GPBA_FAIL = -1 } gpba_solver_result;

/* Reduced-camera-system solver = which g2o::LinearSolver the caller would have constructed.
 *   DENSE_CHOL   LinearSolverDense (Eigen::LDLT, g2o/solvers/linear_solver_dense.h:65-113; LocalGPBA, Optimizer.cc:841)
 *   SPARSE_CHOL  LinearSolverEigen (SimplicialLDLT + AMD, linear_solver_eigen.h:94-124; BundleAdjustment, Optimizer.cc:70)
 *   PCG          block-Jacobi preconditioned CG on the Hschur blocks (an inexact alternative, no reference counterpart)
 * On the device DENSE_CHOL and SPARSE_CHOL run the same tile-sparse DMMA Cholesky (a dense system is the case where
 * every tile is present); the distinction matters to the CPU oracle, which restates the two Eigen solvers separately. */
typedef enum gpba_linear_solver { GPBA_SOLVER_DENSE_CHOL = 0, GPBA_SOLVER_PCG = 1, GPBA_SOLVER_SPARSE_CHOL = 2 } gpba_linear_solver;

/* obs_flags bits */
#define GPBA_OBS_CLOSE 0x1u   /* MapPoint::mvTrackDepth[cam] < 10 m  (Optimizer.cc:1273)     */
#define GPBA_OBS_LEVEL1 0x2u  /* edge->level()==1: inactive          (optimizable_graph.h:467) */
#define GPBA_OBS_NO_KERNEL 0x4u /* setRobustKernel(0)                 (Optimizer.cc:611)       */

/*
 * Flattened problem = what the adapter extracts from the g2o graph that
 * Optimizer::BundleAdjustment (src/Optimizer.cc:61-367) / Optimizer::LocalGPBA (:713-1432)
 * build.  Keyframes are listed in ascending vertex id, so the Hessian index of a free
 * keyframe is its rank among free keyframes (SparseOptimizer::buildIndexMapping,
 * sparse_optimizer.cpp:166-190); points follow (marginalized).
 *
 * A "record" is one (KF_prev, KF_cur, camera, capture time) tuple: everything a GP edge
 * computes that does not depend on the landmark.  rec_kf1 == -1 marks a
 * synchronous record (EdgeMono / EdgeStereo on the reference camera, G2oTypes.h:423-468).
 * An observation with obs_ur >= 0 is a stereo edge (EdgeStereo / EdgeStereoGP).
 */
typedef struct gpba_problem {
  int32_t n_cam;                 /* async cameras + the reference camera (last)                    */
  const double* cam_intr;        /* [n_cam][4] fx fy cx cy, float-rounded (GeometricCamera.h:101)  */
  const double* cam_Tbc;         /* [n_cam][7] qx qy qz qw tx ty tz of MultiKeyFrame::mTbc / VertexExtrinsic */
  double bf;                     /* PoseVelocity::bf (G2oTypes.h:78)                               */

  int32_t n_kf;
  const double* kf_pose;         /* [n_kf][7] Twb: qx qy qz qw tx ty tz   (PoseVelocity::Twb)      */
  const double* kf_vel;          /* [n_kf][6] [lin; ang]                  (PoseVelocity::Vel)      */
  const double* kf_time;         /* [n_kf]                                (PoseVelocity::time)     */
  const uint8_t* kf_fixed;       /* [n_kf] vertex->fixed()                                         */

  int32_t n_pt;
  const double* pt_xyz;          /* [n_pt][3] VertexSBAPointXYZ estimate                           */

  int32_t n_rec;
  const int32_t* rec_kf1;        /* [n_rec] index into kf arrays, -1 = synchronous record          */
  const int32_t* rec_kf2;        /* [n_rec]                                                        */
  const int32_t* rec_cam;        /* [n_rec] cam_idx (EdgeMonoGP::cam_idx, G2oTypes.h:398)          */
  const double* rec_t;           /* [n_rec] capture time t (EdgeMonoGP::t, G2oTypes.h:399)         */

  int64_t n_obs;                 /* reprojection edges, in g2o insertion (internalId) order        */
  const double* obs_u;           /* [n_obs] measurement()[0]                                       */
  const double* obs_v;           /* [n_obs] measurement()[1]                                       */
  const double* obs_ur;          /* [n_obs] measurement()[2], <0 => mono; NULL => all mono         */
  const double* obs_inv_sigma2;  /* [n_obs] information()(0,0) (float-rounded, Optimizer.cc:182)   */
  const int32_t* obs_rec;        /* [n_obs]                                                        */
  const int32_t* obs_pt;         /* [n_obs]                                                        */
  const uint8_t* obs_flags;      /* [n_obs] GPBA_OBS_* bits; NULL => 0                             */

  int32_t n_prior;               /* EdgeGaussianPrior (G2oTypes.h:147-184), info = QiInv(dt)       */
  const int32_t* prior_kf1;
  const int32_t* prior_kf2;
  int32_t n_velp;                /* EdgeVelocity (G2oTypes.h:496-519), info = QcInv(2,2)           */
  const int32_t* velp_kf;

  double qc[6];                  /* diagonal of GaussianProcess::mQc (GaussianProcess.h:59)        */
  double huber_mono;             /* RobustKernelHuber delta, (float)sqrt(5.991); 0 = no kernel     */
  double huber_stereo;           /* (float)sqrt(7.815)                                             */
  double huber_prior;            /* 21.026 in BundleAdjustment (Optimizer.cc:128-130), 0 in LocalGPBA */
  double lambda_init;            /* setUserLambdaInit: 1e-5 global (:75), 1.0 local (:854)         */
  int32_t linear_solver;         /* gpba_linear_solver                                             */
} gpba_problem;

/* Per-outer-iteration record of what OptimizationAlgorithmLevenberg::solve did
 * (optimization_algorithm_levenberg.cpp:61-169); field names follow G2OBatchStatistics
 * (g2o/core/batch_stats.h:39-78) where one exists. */
typedef struct gpba_lm_trace {
  int32_t n_iters;                       /* cjIterations returned by optimize()              */
  int32_t result;                        /* gpba_solver_result of the last solve()           */
  int32_t levenberg_iterations[GPBA_MAX_ITERS]; /* trials (qmax) of each outer iteration     */
  double chi2_before[GPBA_MAX_ITERS];    /* iniChi: robust chi2 at linearisation             */
  double chi2_after[GPBA_MAX_ITERS];     /* currentChi when solve() returned                 */
  double lambda[GPBA_MAX_ITERS];         /* _currentLambda when solve() returned             */
  int32_t total_trials;
  int32_t cg_iterations;                 /* PCG only: total CG iterations                    */
  double last_trial_chi2;                /* tempChi of the last evaluated trial (stale-error quirk) */
} gpba_lm_trace;

typedef struct gpba_lm_params {
  int32_t max_trials_after_failure;      /* 10  (optimization_algorithm_levenberg.cpp:50)    */
  double tau;                            /* 1e-5 (:46) used when lambda_init <= 0            */
  double good_step_lower;                /* 1/3 (:48) */
  double good_step_upper;                /* 2/3 (:47) */
  double pcg_tolerance;                  /* relative residual for PCG (parity mode 1e-12)    */
  int32_t pcg_max_iterations;
} gpba_lm_params;

/* chi2 thresholds of LocalGPBA's inlier check (Optimizer.cc:975-978,1263-1348), float-rounded. */
typedef struct gpba_thresholds {
  double chi2_mono;        /* (float)5.991  */
  double chi2_mono_close;  /* 1.5f*(float)5.991 */
  double chi2_stereo;      /* (float)7.815  */
} gpba_thresholds;

/* Sizes fixed by gpba_build_structure (BlockSolver::buildStructure, block_solver.hpp:142-295). */
typedef struct gpba_structure_info {
  int32_t n_free_kf;       /* _numPoses                                 */
  int32_t n_active_pt;     /* _numLandmarks                             */
  int64_t n_active_obs;    /* active reprojection edges                 */
  int64_t n_hpl;           /* #(pose, landmark) blocks of Hpl (counted only when `info` is asked for) */
  int32_t n_hpp;           /* #upper blocks of Hpp (incl. diagonal)     */
  int32_t n_hschur;        /* #upper blocks of Hschur (incl. diagonal)  */
} gpba_structure_info;

typedef struct gpba_handle gpba_handle;

/* ---- lifetime ------------------------------------------------------------------ */
/* replaces: new g2o::BlockSolverX(linearSolver) + new OptimizationAlgorithmLevenberg
 * (Optimizer.cc:68-76, 840-856).  device < 0 => current device.  Uploads the problem. */
int gpba_create(const gpba_problem* prob, int device, gpba_handle** out);
/* replaces: ~SparseOptimizer -> ~OptimizationAlgorithmWithHessian -> ~BlockSolver
 * (sparse_optimizer.cpp:56-59, optimization_algorithm_with_hessian.cpp:45-48). */
int gpba_destroy(gpba_handle* h);
const char* gpba_last_error(void);
void gpba_default_lm_params(gpba_lm_params* p);
/* Multi-GPU (global BA): this rank owns the points with (pt % nranks == rank)-style shard
 * given by gpba_shard_points; partial Hschur/bschur/chi2 are summed with ncclAllReduce.
 * nccl_unique_id = 128 bytes from ncclGetUniqueId on rank 0 (gpba_nccl_unique_id).  The communicator is created at
 * the first gpba_create_dist with a given id and shared by every later handle of the process that passes the same id
 * (one id per set of GPUs, many BA calls). */
int gpba_nccl_unique_id(unsigned char id_out[128]);
int gpba_create_dist(const gpba_problem* prob, int device, int rank, int nranks,
                     const unsigned char nccl_unique_id[128], gpba_handle** out);

/* General form of the two calls above.  GPBA_CREATE_ASYNC_UPLOAD: the per-observation measurement arrays (obs_u, obs_v,
 * obs_ur, obs_inv_sigma2 -- two thirds of the upload) are copied on a second stream while gpba_build_structure already
 * sorts and pairs the index arrays; with this flag those four arrays must stay valid and unchanged until the first
 * gpba_build_structure / gpba_optimize / gpba_destroy call on the handle returns (without it every array may be freed
 * as soon as gpba_create returns).  Pinned host memory is needed for the copies to overlap. */
#define GPBA_CREATE_ASYNC_UPLOAD 0x1u
typedef struct gpba_create_options {
  int32_t device;                /* < 0: current device                         */
  int32_t rank, nranks;          /* nranks <= 1: single GPU                     */
  const unsigned char* nccl_id;  /* [128], required when nranks > 1             */
  uint32_t flags;                /* GPBA_CREATE_*                               */
} gpba_create_options;
int gpba_create_ex(const gpba_problem* prob, const gpba_create_options* opt, gpba_handle** out);

/* ---- L1: g2o::Solver / BlockSolver-shaped --------------------------------------- */
/* BlockSolver::buildStructure (block_solver.hpp:142-295) + SparseOptimizer::initializeOptimization
 * (sparse_optimizer.cpp:199-267): active set, index mapping, Hpp/Hpl/Hschur block pattern. */
int gpba_build_structure(gpba_handle* h, gpba_structure_info* info);
/* Upper-triangular block (row, col) lists, sorted by (col, row) like SparseBlockMatrix columns
 * (sparse_block_matrix.h); used for the bit-exact pattern comparison. */
int gpba_get_hpp_pattern(gpba_handle* h, int32_t* rows, int32_t* cols);
int gpba_get_hschur_pattern(gpba_handle* h, int32_t* rows, int32_t* cols);
/* SparseOptimizer::computeActiveErrors + activeRobustChi2 (sparse_optimizer.cpp:61-114). */
int gpba_compute_errors(gpba_handle* h, double* robust_chi2);
/* BlockSolver::buildSystem (block_solver.hpp:502-560). */
int gpba_build_system(gpba_handle* h);
/* BlockSolver::setLambda / restoreDiagonal (block_solver.hpp:563-604). */
int gpba_set_lambda(gpba_handle* h, double lambda, int backup);
int gpba_restore_diagonal(gpba_handle* h);
/* BlockSolver::solve (block_solver.hpp:353-486): Schur, reduced solve, back-substitution.
 * *ok = 0 mirrors solve()==false (non-positive reduced system). */
int gpba_solve(gpba_handle* h, int* ok);
/* Solver::x() / Solver::b() / vectorSize() (solver.h:95-102): poses (12 each) then landmarks (3 each). */
int gpba_vector_size(gpba_handle* h, int64_t* n);
int gpba_get_x(gpba_handle* h, double* x);
int gpba_get_b(gpba_handle* h, double* b);
/* Block values for parity tests; layouts: Hpp/Hschur blocks 12x12 row-major in pattern order,
 * Hll 3x3 row-major per active landmark, Hpl 12x3 row-major in (landmark, pose) order. */
int gpba_get_hpp(gpba_handle* h, double* blocks);
int gpba_get_hschur(gpba_handle* h, double* blocks, double* bschur);
int gpba_get_hll(gpba_handle* h, double* blocks);
int gpba_get_hpl(gpba_handle* h, int64_t* lm_begin /* [n_active_pt+1] */, int32_t* pose /* [n_hpl] */, double* blocks /* [n_hpl][12][3] */);
/* SparseOptimizer::update (oplus on every free vertex, sparse_optimizer.cpp:422-435),
 * push / pop / discardTop (sparse_optimizer.cpp:600-613). */
int gpba_oplus(gpba_handle* h, const double* x /* NULL = the solver's own x */);
int gpba_push(gpba_handle* h);
int gpba_pop(gpba_handle* h);
int gpba_discard_top(gpba_handle* h);

/* ---- L2: whole optimize ---------------------------------------------------------- */
/* SparseOptimizer::optimize(iters) with OptimizationAlgorithmLevenberg (one call = the whole
 * LM loop on device).  stop_flag mirrors setForceStopFlag(bool*) (sparse_optimizer.h:188);
 * it is polled once per LM trial and between outer iterations; may be NULL. */
int gpba_optimize(gpba_handle* h, int iters, const volatile unsigned char* stop_flag,
                  const gpba_lm_params* params, gpba_lm_trace* trace);
/* vertex->estimate() read-back (Optimizer.cc:324-366, 1360-1430). Any pointer may be NULL. */
int gpba_download_state(gpba_handle* h, double* kf_pose, double* kf_vel, double* pt_xyz);
/* edge->chi2() of every reprojection edge from its stored _error (base_edge.h:58-61),
 * original observation order; inactive edges keep their last computed error. */
int gpba_edge_chi2(gpba_handle* h, double* chi2);
/* BaseEdge::_error of every ACTIVE reprojection edge as the last evaluation left it ([n_obs][3], third slot 0 for a
 * monocular edge; NaN for inactive edges = leave the edge's error alone).  After a rejected last trial these are the
 * rejected state's errors (the stale-error quirk of sparse_optimizer.cpp:354-419 + optimization_algorithm_levenberg.cpp:
 * 102-166): the adapter writes them into OptimizableGraph::Edge::errorData() instead of recomputing at the estimate. */
int gpba_edge_errors(gpba_handle* h, double* err3);
/* Keyframe states / extrinsics of the last EVALUATED state (= gpba_download_state's unless the last trial was rejected). */
int gpba_download_evaluated_state(gpba_handle* h, double* kf_pose, double* kf_vel, double* cam_Tbc);
/* activeRobustChi2() over the stored errors (what LocalGPBA reads as err / err_end, Optimizer.cc:1223,1254). */
int gpba_active_robust_chi2(gpba_handle* h, double* chi2);
/* LocalGPBA's inlier check (Optimizer.cc:1263-1348): flag = chi2 > threshold || !isDepthPositive. */
int gpba_outlier_flags(gpba_handle* h, const gpba_thresholds* th, uint8_t* flags);
/* e->setLevel / e->setRobustKernel(0) / e->computeError() on inactive edges, as the chi2
 * rejection rounds do (Optimizer.cc:559-675). */
int gpba_set_levels(gpba_handle* h, const uint8_t* level /* [n_obs] 0/1 */);
int gpba_set_robust_kernel(gpba_handle* h, int enabled);
int gpba_compute_errors_inactive(gpba_handle* h);
/* 4 x (initializeOptimization(0) + optimize(iters)) with re-flagging after every round,
 * kernel removed after round index 2 (structure of Optimizer.cc:548-675 with LocalGPBA thresholds). */
int gpba_rejection_rounds(gpba_handle* h, int n_rounds, int iters, const gpba_thresholds* th,
                          const gpba_lm_params* params, uint8_t* flags_out, gpba_lm_trace* traces /* [n_rounds] */);

/* ---- extrinsic self-calibration ---------------------------------- */
/* LocalGPBA's second stage (src/Optimizer.cc:983-995, 1228-1240, 1419-1428): every asynchronous camera carries a
 * VertexExtrinsic (6-dim, Tbc <- Tbc exp(delta), include/G2oTypes.h:83-102), fixed during the first optimize(); cameras
 * with at least extrin_thresh = 50 EdgeMonoGPExtrinsic are then un-fixed and the graph is optimised again, with an
 * EdgeExtrinsicPrior (e = Log(R_ini^-1 R_bc), information MultiFrame::mRbc_ini_cov[c], G2oTypes.h:470-494) on every free
 * extrinsic.  A free extrinsic is a non-marginalized vertex whose id follows every keyframe id (:986): in every vector and
 * block pattern of this library it is a pose block AFTER the keyframes' -- a 12-slot whose first six entries are the
 * extrinsic's tangent [translation; rotation] and whose last six are padding (zero). */
typedef struct gpba_extrinsics {
  const uint8_t* free_mask;  /* [n_cam] 1 = VertexExtrinsic::setFixed(false) (:1236)                                   */
  const double* prior_R;     /* [n_cam][4] qx qy qz qw of MultiFrame::mRbc_ini[c]; NULL = no EdgeExtrinsicPrior edges  */
  const double* prior_info;  /* [n_cam][9] MultiFrame::mRbc_ini_cov[c], row-major; NULL iff prior_R is NULL            */
} gpba_extrinsics;
/* Takes effect at the next gpba_build_structure / gpba_optimize (= initializeOptimization, :1238).  Single GPU only. */
int gpba_set_extrinsics(gpba_handle* h, const gpba_extrinsics* ext);
/* VertexExtrinsic::estimate() read-back (:1419-1428): [n_cam][7] qx qy qz qw tx ty tz of the current estimate. */
int gpba_get_extrinsics(gpba_handle* h, double* cam_Tbc);
/* cam_obs[c]: EdgeMonoGPExtrinsic of camera c in the graph, whatever their level (`cam_obs[c]++`, :1129). */
int gpba_count_camera_observations(gpba_handle* h, int64_t* cam_obs /* [n_cam] */);
/* The whole second stage (:1226-1240): un-fix the candidates with cam_obs >= min_obs (50 in the reference), then
 * initializeOptimization + optimize(iters).  freed_out [n_cam] (optional) reports which extrinsics were un-fixed. */
int gpba_calibrate_extrinsics(gpba_handle* h, const gpba_extrinsics* candidates, int min_obs, int iters,
                              const gpba_lm_params* params, gpba_lm_trace* trace, uint8_t* freed_out);

/* ---- pose-only GP optimisation ---------------------------------- */
/* Optimizer::PoseGPOptimizationFromeLastFrame (src/Optimizer.cc:369-686) for a batch of independent frames: per frame
 * two VertexPoseVel (previous frame, fixed iff prev_fixed; current frame, free), one EdgeMonoGPOnlyPose per match of an
 * asynchronous camera, EdgeMonoOnlyPose / EdgeStereoOnlyPose per match of the reference camera (the world point is a
 * constant of the edge, src/G2oTypes.cc:120-223), one EdgeGaussianPrior without kernel (:526-533) and an EdgeVelocity on
 * both vertices (:535-543).  Four rounds of initializeOptimization(0) + optimize(10) with default lambda
 * (tau * max diag), re-flagging after every round with float-typed chi2 and thresholds chi2Mono = 5.991 (x1.5 for close
 * points), chi2Stereo = {15.6, 9.8, 7.815, 7.815}, kernels removed after the third round (:545-670). */
#define GPBA_POSE_ROUNDS 4
typedef struct gpba_pose_batch {
  int32_t n_cam;                 /* async cameras + the reference camera (last)                              */
  const double* cam_intr;        /* [n_cam][4] fx fy cx cy (float-rounded)                                   */
  const double* cam_Tbc;         /* [n_cam][7] qx qy qz qw tx ty tz of MultiKeyFrame::mTbc                   */
  double bf;
  double qc[6];                  /* diagonal of GaussianProcess::mQc                                         */
  int32_t n_frames;
  const double* prev_pose;       /* [n_frames][7] Twb of pFrame->mpPrevFrame                                 */
  const double* prev_vel;        /* [n_frames][6]                                                            */
  const double* prev_time;       /* [n_frames]                                                               */
  const uint8_t* prev_fixed;     /* [n_frames] the `fix` argument                                            */
  const double* cur_pose;        /* [n_frames][7] Twb of pFrame (initial estimate)                           */
  const double* cur_vel;         /* [n_frames][6]                                                            */
  const double* cur_time;        /* [n_frames] pFrame->mTimeStamp                                            */
  const double* cam_time;        /* [n_frames][n_cam] pFrame->mvTimeStamps                                   */
  const int64_t* obs_begin;      /* [n_frames+1] matches of frame f = [obs_begin[f], obs_begin[f+1])         */
  const double* obs_u;           /* kpUn.pt.x                                                                */
  const double* obs_v;
  const double* obs_ur;          /* mvuRight (reference camera only), < 0 => monocular; NULL => all mono     */
  const double* obs_inv_sigma2;  /* mvInvLevelSigma2[octave] / uncertainty2 (float-rounded)                  */
  const double* obs_xw;          /* [n_obs][3] pMP->GetWorldPos(), constant                                  */
  const int32_t* obs_cam;        /* mmpKeyToCam                                                              */
  const uint8_t* obs_flags;      /* GPBA_OBS_CLOSE (mvTrackDepth < 10 m), GPBA_OBS_LEVEL1 (mvbOutlier on entry) */
  double huber_mono;             /* (float)sqrt(5.991)                                                       */
  double huber_stereo;           /* (float)sqrt(7.815)                                                       */
} gpba_pose_batch;
/* One CTA per frame runs the whole 4 x 10 LM schedule on the device.  Outputs (any may be NULL): optimised states,
 * mvbOutlier per match, nInitialCorrespondences - nBad per frame (the return value of the reference function) and the
 * LM trace of every round ([n_frames][GPBA_POSE_ROUNDS]). */
int gpba_pose_optimize(const gpba_pose_batch* batch, int device, double* cur_pose_out, double* cur_vel_out,
                       double* prev_pose_out, double* prev_vel_out, uint8_t* outlier_out, int32_t* n_inliers_out,
                       gpba_lm_trace* traces);

/* ---- velocity RANSAC -------------------------------------------- */
/* Tracking::MCRansac (src/Tracking.cc:1939-2002) = maxIt x Optimizer::OptimizeVel (src/Optimizer.cc:2364-2447) over one
 * frame pair, as ONE call: every hypothesis is a g2o graph with a single VertexVel (6-dim body twist, additive oplus,
 * include/G2oTypes.h:128-145) and one EdgeVelReproj per matched feature (G2oTypes.h:521-547, src/G2oTypes.cc:497-510:
 * e = obs - project((T_last exp(v dt_cam) T_bc)^-1 X_w), Huber delta = 5.991, information = I * invSigma2), of which only
 * the `set_size` sampled edges are active (level 0); optimize(40) with the default lambda; afterwards every edge is
 * re-evaluated and counted as inlier when |e| <= threshold.  The caller draws the sample sets (std::mt19937 in the
 * reference) and keeps the best hypothesis: best_out = first hypothesis with the largest inlier count (`inliers >
 * bestInliers`, Tracking.cc:1973). */
typedef struct gpba_vel_batch {
  int32_t n_cam;
  const double* cam_intr;        /* [n_cam][4] fx fy cx cy                                              */
  const double* cam_Tbc;         /* [n_cam][7] MultiFrame::mTbc                                         */
  const double* cam_dt;          /* [n_cam] pF1->mvTimeStamps[cam] - pF2->mTimeStamp                    */
  double last_pose[7];           /* pF2->GetPoseW(): Twb of the last frame                              */
  double vel_init[6];            /* pF1->GetVelocity(): initial estimate of every hypothesis            */
  int32_t n_match;               /* vMatchedFeatures.size()                                             */
  const double* obs_u;           /* [n_match] kpUn.pt.x                                                 */
  const double* obs_v;
  const double* obs_inv_sigma2;  /* [n_match] mvInvLevelSigma2[octave]                                  */
  const double* obs_xw;          /* [n_match][3] pMP->GetWorldPos()                                     */
  const int32_t* obs_cam;        /* [n_match] mmpKeyToCam                                               */
  int32_t n_hyp;                 /* maxIt (23 at Tracking.cc:2029)                                      */
  int32_t set_size;              /* min_set = 3 (Tracking.cc:1947); <= 8                                */
  const int32_t* samples;        /* [n_hyp][set_size] indices into the match arrays, distinct per row   */
  double huber_delta;            /* 5.991 (Optimizer.cc:2410)                                           */
  double threshold;              /* 2.0 px (include/Optimizer.h:86)                                     */
  int32_t iterations;            /* 40 (Optimizer.cc:2423)                                              */
} gpba_vel_batch;
/* Outputs (any may be NULL): vel_out [n_hyp][6], inliers_out [n_hyp], inlier_mask_out [n_hyp][n_match] (vbInliers),
 * best_out (index of the winning hypothesis, -1 if no hypothesis has an inlier), traces [n_hyp]. */
int gpba_vel_ransac(const gpba_vel_batch* batch, int device, double* vel_out, int32_t* inliers_out,
                    uint8_t* inlier_mask_out, int32_t* best_out, gpba_lm_trace* traces);

/* ---- essential-graph optimisation ------------------------------- */
/* The optimisation inside Optimizer::OptimizeEssentialGraph (src/Optimizer.cc:1434-1717), the step right before the global
 * BA: a pose graph of VertexSim3Expmap (7-dim, S <- Sim3(update) S, update = [omega; upsilon; sigma], sigma forced to 0 when
 * fix_scale; Thirdparty/g2o/g2o/types/types_seven_dof_expmap.h:48-96, sim3.h) joined by EdgeSim3 (e = Log(S_ji S_i S_j^-1),
 * information = I_7, :99-126, Optimizer.cc:1505, 1535-1541), solved with BlockSolver_7_3 + LinearSolverEigen and
 * Levenberg-Marquardt, lambda_0 = 1e-16 (:1441-1448), optimize(20) (:1667).  EdgeSim3 has no analytic Jacobian: g2o
 * differentiates numerically (central differences, delta = 1e-9, core/base_binary_edge.hpp:131-200) and so does the kernel.
 * WHICH edges exist (loop connections, spanning tree, loop edges, covisibility >= 100, :1507-1640) is the caller's map
 * logic; the call takes the flattened graph.  Keyframes in ascending id (= Hessian order). */
typedef struct gpba_pose_graph {
  int32_t n_kf;
  const double* sim3;        /* [n_kf][8] qx qy qz qw tx ty tz s: VertexSim3Expmap estimate S_iw (:1468-1482)        */
  const uint8_t* fixed;      /* [n_kf] setFixed(true) for the map's initial keyframe (:1484-1485)                    */
  int32_t fix_scale;         /* VertexSim3Expmap::_fix_scale = bFixScale (:1489)                                     */
  int64_t n_edge;
  const int32_t* edge_i;     /* [n_edge] vertex 0 of the edge (setVertex(0, ...nIDi), :1533)                         */
  const int32_t* edge_j;     /* [n_edge] vertex 1 (:1532)                                                            */
  const double* edge_meas;   /* [n_edge][8] measurement S_ji (:1534)                                                 */
  double lambda_init;        /* setUserLambdaInit(1e-16) (:1447); <= 0: tau * max diagonal                           */
} gpba_pose_graph;
/* sim3_out [n_kf][8]: the optimised S_iw (CorrectedSiw, :1677); trace may be NULL. */
int gpba_pose_graph_optimize(const gpba_pose_graph* graph, int device, int iters, const gpba_lm_params* params,
                             double* sim3_out, gpba_lm_trace* trace);
/* Map point correction after the pose graph (:1687-1712): P <- S_wr' (S_rw P), r = ref_kf[p] the point's reference
 * keyframe, S_rw from sim3_before, S_wr' the inverse of sim3_after[r].  xyz_out may alias xyz. */
int gpba_correct_points(int device, int64_t n_pt, const double* xyz, const int32_t* ref_kf, int32_t n_kf,
                        const double* sim3_before, const double* sim3_after, double* xyz_out);

/* ---- measurement ------------------------------------------------------------------- */
/* Total device time (ms, CUDA event pairs recorded on the library stream, read back only here) and
 * launch count per stage since the last reset: 0 records (K0), 1 residuals (K1), 2 linearize landmarks
 * (K2a), 3 linearize poses + priors (K2b/K2c/K3), 4 schur prepare (K4a), 5 schur record-pair products (K4b),
 * 6 reduced-system factorization + forward substitution / PCG (K5), 7 backward substitution,
 * 8 back-substitution + update (K6), 9 collective, 10 schur expansion into Hschur / bschur (K4c).
 * Grouping follows G2OBatchStatistics (g2o/core/batch_stats.h:39-78): timeResiduals = 0+1,
 * timeQuadraticForm = 2+3, timeSchurComplement = 4+5+10, timeLinearSolver = 6+7, timeUpdate = 8. */
#define GPBA_N_STAGES 11
int gpba_stage_stats(gpba_handle* h, double ms_total[GPBA_N_STAGES], int64_t launches[GPBA_N_STAGES], int reset);
int gpba_set_profiling(gpba_handle* h, int enabled);
/* Sizes of the record-space Schur structures (for the algorithmic-byte accounting of bench.py / DESIGN.md):
 * out[0] observation pairs, out[1] record pairs (6x6 accumulators), out[2] work items of K4b,
 * out[3] (record pair -> Hschur block) contributions of K4c. */
int gpba_schur_stats(gpba_handle* h, int64_t out[4]);
/* Shape of the reduced-system factorization (the analyzePattern of linear_solver_eigen.h:147-201): out[0] tile columns,
 * out[1] levels of the schedule (independent tile columns share a level), out[2] parts (leaves + separators) of the
 * nested-dissection order (1 = plain banded order), out[3] non-zero 48x48 tiles of the factor, out[4] 48x48x48 tile
 * products of the left-looking update (x 2 * 48^3 = its flops), out[5] CTAs of the update launches. */
int gpba_solver_stats(gpba_handle* h, int64_t out[6]);
/* The symbolic phase alone, on the host (no device needed): nested-dissection order of the pose blocks by BFS level
 * structures, tile-level symbolic factorization and level schedule, for an upper block pattern (row <= col) such as the one
 * gpba_get_hschur_pattern returns.  nd_depth < 0: the default (8 for >= 256 pose blocks, else the plain banded order);
 * 0 switches the dissection off.  perm_out[b] (optional) = position of pose block b, in pose-block slots: every part of
 * the order starts on a tile boundary, so positions may skip slots.  out[0] tile columns, out[1] levels, out[2] parts,
 * out[3] non-zero tiles of the factor, out[4] tile pairs of the trailing updates (x 2 * 48^3 = their flops). */
int gpba_symbolic_analyze(int32_t n_pose, int32_t n_hs, const int32_t* hs_row, const int32_t* hs_col, int32_t nd_depth,
                          int32_t* perm_out, int64_t out[5]);
/* Host only (no device needed): builds the task list of the persistent factorization kernel for an upper block pattern and
 * checks what its in-kernel waits rely on -- every task depends only on tasks in front of it in the list (source columns of
 * a chunk complete before it, all chunks of a column before its panel tasks, solve-only tasks behind the task that publishes
 * their diagonal factor) and the completion counts match.  out[0] tasks, out[1] update chunks, out[2] tile products,
 * out[3] panel tasks, out[4] violations (0 for a valid schedule). */
int gpba_factor_schedule_check(int32_t n_pose, int32_t n_hs, const int32_t* hs_row, const int32_t* hs_col, int64_t out[5]);
/* cudaStream_t every kernel of this handle is launched on (for CUDA-event timing by the caller). */
void* gpba_get_stream(gpba_handle* h);
/* Re-upload estimates only (same structure): lets a benchmark repeat optimize() from the same start. */
int gpba_reset_state(gpba_handle* h, const double* kf_pose, const double* kf_vel, const double* pt_xyz);

#ifdef __cplusplus
}
#endif
#endif /* GPBA_H */
