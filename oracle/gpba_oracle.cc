// oracle/gpba_oracle.cc -- TEST INFRASTRUCTURE ONLY.
// CPU restatement of everything AMC-SLAM executes inside g2o::SparseOptimizer::optimize for
// Optimizer::BundleAdjustment / Optimizer::LocalGPBA.  Only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs may load this library; libgpba.so never does.
//
// PARITY (the reference holds no golden vectors and its own build -- cmake, Eigen3, OpenCV, Boost -- cannot run
// here): PINNED against the reference's own code.  oracle/_ref/libamc_ref_g2o.so is g2o's core, BlockSolverX,
// LinearSolverDense, Levenberg-Marquardt and AMC-SLAM's edge sources compiled UNMODIFIED against stand-in Eigen / Sophus headers
// (oracle/ref_shim/, oracle/ref_g2o_run.cc); tests/test_whole_path_reference.py holds optimize() of this file to its runs on 8 problems incl.
// BASELINE C1: identical iteration and trial counts, cost 1e-9, poses 1e-8 m.  Beside it the LM controller alone (g2o's
// optimization_algorithm_levenberg.cpp on this file's level-1 steps: bit for bit, oracle/ref_lm_pin.cc) and every edge probe by
// probe (tests/test_ref_pin.py).  Not reachable that way: the sparse linear solver (Eigen's SimplicialLDLT; unique solution,
// checked against numpy normal equations) and Optimizer.cc's own logic (flags, rejection rounds, extrinsic stage: restated).
//
// Follows (paths relative to the AMC-SLAM tree, g2o = Thirdparty/g2o/g2o):
//   g2o/core/sparse_optimizer.cpp:199-267  initializeOptimization (active set)      -> build_structure()
//   g2o/core/sparse_optimizer.cpp:166-190  buildIndexMapping                        -> build_structure()
//   g2o/core/block_solver.hpp:142-295      buildStructure (Hpp/Hpl/Hll/Hschur)      -> build_structure()
//   g2o/core/sparse_optimizer.cpp:61-114   computeActiveErrors / activeRobustChi2   -> compute_errors()/robust_chi2()
//   g2o/core/block_solver.hpp:502-560      buildSystem                              -> build_system()
//   g2o/core/base_{unary,binary,multi}_edge.hpp constructQuadraticForm              -> accumulate_*()
//   g2o/core/base_edge.h:58-61,96-102      chi2, robustInformation (rho' * Omega only)
//   g2o/core/block_solver.hpp:563-604      setLambda / restoreDiagonal
//   g2o/core/block_solver.hpp:353-486      solve (Schur, linear solve, back-substitution)
//   g2o/solvers/linear_solver_dense.h:65-113  dense LDLT (Eigen::LDLT, pivoted; third party) -> ldlt_dense()
//   g2o/solvers/linear_solver_eigen.h:94-124  SimplicialLDLT + AMD (third party)             -> BlockSparseChol
//   g2o/core/optimization_algorithm_levenberg.cpp:61-194  LM controller             -> lm_solve()
//   g2o/core/sparse_optimizer.cpp:354-435,600-613  optimize / update / push / pop   -> optimize()
//   src/Optimizer.cc:1263-1348             LocalGPBA inlier check                   -> outlier_flags()
//   src/Optimizer.cc:548-675               chi2 rejection rounds structure          -> rejection_rounds()
#include "../include/gpba.h"
#include "gp_edges.h"
#include "pose_graph.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <limits>
#include <map>
#include <set>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

using namespace ora;

namespace {

typedef Mat<12, 3> M12x3;

// ----------------------------------------------------------------------------- dense LDLT
// Eigen::LDLT<MatrixXd, Lower> (robust Cholesky with diagonal pivoting), restated from its
// published algorithm (Eigen/src/Cholesky/LDLT.h, ldlt_inplace<Lower>::unblocked); Eigen itself
// is un-vendored and unpinned (CMakeLists.txt:48).  Returns isPositive().
static bool ldlt_dense(int n, std::vector<double>& A /* n*n row-major, full symmetric */, const double* b, double* x) {
  std::vector<int> transp(n);
  int sign = 0;  // ZeroSign; 1 = PositiveSemiDef, -1 = NegativeSemiDef, 2 = Indefinite
  std::vector<double> temp(n);
  auto at = [&](int r, int c) -> double& { return A[(size_t)r * n + c]; };
  for (int k = 0; k < n; ++k) {
    int p = k;
    double big = std::fabs(at(k, k));
    for (int i = k + 1; i < n; ++i)
      if (std::fabs(at(i, i)) > big) { big = std::fabs(at(i, i)); p = i; }
    transp[k] = p;
    if (p != k) {  // symmetric swap in the lower triangle
      int s = n - p - 1;
      for (int c = 0; c < k; ++c) std::swap(at(k, c), at(p, c));
      for (int r = p + 1; r < p + 1 + s; ++r) std::swap(at(r, k), at(r, p));
      std::swap(at(k, k), at(p, p));
      for (int i = k + 1; i < p; ++i) std::swap(at(i, k), at(p, i));
    }
    int rs = n - k - 1;
    if (k > 0) {
      for (int j = 0; j < k; ++j) temp[j] = at(j, j) * at(k, j);
      double s = 0;
      for (int j = 0; j < k; ++j) s += at(k, j) * temp[j];
      at(k, k) -= s;
      if (rs > 0)
        for (int i = k + 1; i < n; ++i) {
          double acc = 0;
          const double* row = &A[(size_t)i * n];
          for (int j = 0; j < k; ++j) acc += row[j] * temp[j];
          at(i, k) -= acc;
        }
    }
    double realAkk = at(k, k);
    if (rs > 0 && std::fabs(realAkk) > 0.0)
      for (int i = k + 1; i < n; ++i) at(i, k) /= realAkk;
    if (sign == 1) { if (realAkk < 0) sign = 2; }
    else if (sign == -1) { if (realAkk > 0) sign = 2; }
    else if (sign == 0) { if (realAkk > 0) sign = 1; else if (realAkk < 0) sign = -1; }
  }
  bool positive = (sign == 1 || sign == 0);
  if (!positive) return false;
  // solve: x = P^T L^-T D^+ L^-1 P b
  std::vector<double> y(b, b + n);
  for (int k = 0; k < n; ++k) std::swap(y[k], y[transp[k]]);
  for (int i = 0; i < n; ++i) {
    double s = y[i];
    const double* row = &A[(size_t)i * n];
    for (int j = 0; j < i; ++j) s -= row[j] * y[j];
    y[i] = s;
  }
  const double tol = 1.0 / std::numeric_limits<double>::max();
  for (int i = 0; i < n; ++i) y[i] = std::fabs(at(i, i)) > tol ? y[i] / at(i, i) : 0.0;
  for (int i = n - 1; i >= 0; --i) {
    double s = y[i];
    for (int j = i + 1; j < n; ++j) s -= at(j, i) * y[j];
    y[i] = s;
  }
  for (int k = n - 1; k >= 0; --k) std::swap(y[k], y[transp[k]]);
  for (int i = 0; i < n; ++i) x[i] = y[i];
  return true;
}

// ----------------------------------------------------------------------------- block-sparse Cholesky
// Stand-in for Eigen::SimplicialLDLT<Upper> with AMD ordering (linear_solver_eigen.h:94-124,147-201):
// a fill-reducing (greedy minimum-degree) block ordering computed once per structure
// ("symbolic once", :151) and a numeric block factorization per call.  The solution of the SPD
// system is unique, so any backward-stable factorization agrees to ~cond * 2^-53.
struct BlockSparseChol {
  int nb = 0;
  std::vector<int> perm, iperm;             // perm[new] = old
  std::vector<std::vector<int>> col_rows;   // per new column k: sorted rows > k of L (after fill)
  std::vector<std::vector<M12>> col_blk;    // matching blocks
  std::vector<M12> diag;
  bool analyzed = false;

  void analyze(int nb_, const std::vector<std::pair<int, int>>& upper_rc) {
    nb = nb_;
    std::vector<std::set<int>> adj(nb);
    for (auto& rc : upper_rc)
      if (rc.first != rc.second) { adj[rc.first].insert(rc.second); adj[rc.second].insert(rc.first); }
    perm.clear(); iperm.assign(nb, -1);
    std::vector<char> done(nb, 0);
    std::vector<std::set<int>> work = adj;
    std::multimap<int, int> dummy;
    std::set<std::pair<int, int>> heap;  // (degree, node)
    for (int i = 0; i < nb; ++i) heap.insert({(int)work[i].size(), i});
    std::vector<std::vector<int>> elim_nb(nb);
    while (!heap.empty()) {
      int v = heap.begin()->second;
      heap.erase(heap.begin());
      done[v] = 1;
      iperm[v] = (int)perm.size();
      perm.push_back(v);
      std::vector<int> nbrs(work[v].begin(), work[v].end());
      elim_nb[v] = nbrs;
      for (int a : nbrs) {
        heap.erase({(int)work[a].size(), a});
        work[a].erase(v);
      }
      for (size_t i = 0; i < nbrs.size(); ++i)
        for (size_t j = i + 1; j < nbrs.size(); ++j) { work[nbrs[i]].insert(nbrs[j]); work[nbrs[j]].insert(nbrs[i]); }
      for (int a : nbrs) heap.insert({(int)work[a].size(), a});
      work[v].clear();
    }
    col_rows.assign(nb, {});
    for (int k = 0; k < nb; ++k) {
      int v = perm[k];
      for (int a : elim_nb[v]) col_rows[k].push_back(iperm[a]);
      std::sort(col_rows[k].begin(), col_rows[k].end());
    }
    col_blk.assign(nb, {});
    for (int k = 0; k < nb; ++k) col_blk[k].resize(col_rows[k].size());
    diag.resize(nb);
    analyzed = true;
  }

  int find(int k, int r) const {
    const std::vector<int>& v = col_rows[k];
    return (int)(std::lower_bound(v.begin(), v.end(), r) - v.begin());
  }

  // blocks: upper (row<=col) 12x12 row-major.  Returns false on a non-positive pivot.
  bool factor_solve(const std::vector<std::pair<int, int>>& upper_rc, const std::vector<M12>& blocks, const double* b,
                    double* x) {
    for (int k = 0; k < nb; ++k) {
      diag[k] = M12::Zero();
      for (auto& m : col_blk[k]) m = M12::Zero();
    }
    for (size_t i = 0; i < upper_rc.size(); ++i) {
      int r = iperm[upper_rc[i].first], c = iperm[upper_rc[i].second];
      if (r == c) { diag[r] = blocks[i]; continue; }
      // lower storage: L(row > col); block(i=first,j=second) is A_ij (upper). A_ji = A_ij^T.
      if (r > c) col_blk[c][find(c, r)] = blocks[i];                // A_(r,c) = original (first,second) with first->r: A_rc = block
      else col_blk[r][find(r, c)] = transpose(blocks[i]);           // A_(c,r) = block^T
    }
    for (int k = 0; k < nb; ++k) {
      // dense Cholesky of the 12x12 diagonal block (lower)
      M12& D = diag[k];
      for (int j = 0; j < 12; ++j) {
        double s = D(j, j);
        for (int p = 0; p < j; ++p) s -= D(j, p) * D(j, p);
        if (!(s > 0.0)) return false;
        double l = std::sqrt(s);
        D(j, j) = l;
        for (int i = j + 1; i < 12; ++i) {
          double t = D(i, j);
          for (int p = 0; p < j; ++p) t -= D(i, p) * D(j, p);
          D(i, j) = t / l;
        }
        for (int i = 0; i < j; ++i) D(i, j) = 0.0;
      }
      // L_ik = A_ik * Lkk^-T
      for (auto& B : col_blk[k])
        for (int r = 0; r < 12; ++r)
          for (int j = 0; j < 12; ++j) {
            double t = B(r, j);
            for (int p = 0; p < j; ++p) t -= B(r, p) * D(j, p);
            B(r, j) = t / D(j, j);
          }
      // trailing update
      const std::vector<int>& rows = col_rows[k];
      for (size_t a = 0; a < rows.size(); ++a) {
        const M12& La = col_blk[k][a];
        int ca = rows[a];
        M12 u = La * transpose(La);
        diag[ca] = diag[ca] - u;
        for (size_t c = a + 1; c < rows.size(); ++c) {
          const M12& Lc = col_blk[k][c];
          M12 v = Lc * transpose(La);
          int pos = find(ca, rows[c]);
          col_blk[ca][pos] = col_blk[ca][pos] - v;
        }
      }
    }
    std::vector<double> y((size_t)nb * 12);
    for (int k = 0; k < nb; ++k)
      for (int i = 0; i < 12; ++i) y[(size_t)k * 12 + i] = b[(size_t)perm[k] * 12 + i];
    for (int k = 0; k < nb; ++k) {  // forward
      double* yk = &y[(size_t)k * 12];
      const M12& D = diag[k];
      for (int i = 0; i < 12; ++i) {
        double s = yk[i];
        for (int p = 0; p < i; ++p) s -= D(i, p) * yk[p];
        yk[i] = s / D(i, i);
      }
      for (size_t a = 0; a < col_rows[k].size(); ++a) {
        double* yr = &y[(size_t)col_rows[k][a] * 12];
        const M12& L = col_blk[k][a];
        for (int i = 0; i < 12; ++i) {
          double s = 0;
          for (int p = 0; p < 12; ++p) s += L(i, p) * yk[p];
          yr[i] -= s;
        }
      }
    }
    for (int k = nb - 1; k >= 0; --k) {  // backward
      double* yk = &y[(size_t)k * 12];
      for (size_t a = 0; a < col_rows[k].size(); ++a) {
        const double* yr = &y[(size_t)col_rows[k][a] * 12];
        const M12& L = col_blk[k][a];
        for (int p = 0; p < 12; ++p) {
          double s = 0;
          for (int i = 0; i < 12; ++i) s += L(i, p) * yr[i];
          yk[p] -= s;
        }
      }
      const M12& D = diag[k];
      for (int i = 11; i >= 0; --i) {
        double s = yk[i];
        for (int p = i + 1; p < 12; ++p) s -= D(p, i) * yk[p];
        yk[i] = s / D(i, i);
      }
    }
    for (int k = 0; k < nb; ++k)
      for (int i = 0; i < 12; ++i) x[(size_t)perm[k] * 12 + i] = y[(size_t)k * 12 + i];
    return true;
  }
};

// ----------------------------------------------------------------------------- the optimizer
}  // namespace
#include "pose_only.h"
#include "vel_ransac.h"
namespace {

struct Oracle {
  // ---- problem (deep copy)
  int n_cam = 0, n_kf = 0, n_pt = 0, n_rec = 0, n_prior = 0, n_velp = 0;
  int64_t n_obs = 0;
  std::vector<Pinhole> cams;
  std::vector<SE3> Tbc;
  double bf = 0;
  std::vector<KfState> kf;
  std::vector<uint8_t> kf_fixed;
  std::vector<V3> pt;
  std::vector<int> rec_kf1, rec_kf2, rec_cam;
  std::vector<double> rec_t;
  std::vector<double> obs_u, obs_v, obs_ur, obs_w;
  std::vector<int> obs_rec, obs_pt;
  std::vector<uint8_t> obs_flags;
  std::vector<int> prior_kf1, prior_kf2, velp_kf;
  GaussianProcess G;
  Huber hub_mono, hub_stereo, hub_prior;
  bool has_mono_k = false, has_stereo_k = false, has_prior_k = false;
  double lambda_init = 0;
  int linear_solver = 0;
  int threads = 1;

  // ---- stored edge errors (BaseEdge::_error)
  std::vector<double> obs_err;    // 3 per obs
  std::vector<double> prior_err;  // 12 per prior
  std::vector<double> velp_err;   // 1 per vel edge

  // ---- structure
  std::vector<int> kf_h, pt_h;  // hessianIndex or -1 (pt_h: index among landmarks)
  std::vector<uint8_t> obs_active, prior_active, velp_active;
  std::vector<int64_t> active_obs;  // indices, insertion order
  int numPoses = 0, numLandmarks = 0;
  std::vector<int> lm_pt;  // landmark -> point index
  std::vector<std::pair<int, int>> hpp_rc, hs_rc;
  std::map<std::pair<int, int>, int> hpp_idx, hs_idx;
  std::vector<M12> hpp, hs;
  std::vector<int64_t> lm_begin;  // CSR landmark -> Hpl blocks
  std::vector<int> hpl_pose;
  std::vector<M12x3> hpl;
  std::vector<M3> hll, dinv;
  std::vector<int> obs_slot1, obs_slot2;  // Hpl slot of (kf1, landmark), (kf2, landmark) or -1
  std::vector<std::vector<int>> hs_row_cols;  // per row: sorted cols in Hschur (for the iterator walk)
  std::vector<std::vector<int>> hs_row_idx;
  std::vector<double> b, x, bschur, coeff;
  std::vector<double> diag_backup_pose, diag_backup_lm;
  BlockSparseChol sparse;
#ifdef _OPENMP
  std::vector<omp_lock_t> row_locks;
#endif
  bool structure_ok = false, system_ok = false;
  bool structure_fresh = false;   // build_structure() ran and neither levels nor kernels changed since

  // ---- backup stack (BaseVertex::push/pop)
  struct Snapshot { std::vector<KfState> kf; std::vector<V3> pt; std::vector<SE3> Tbc; };
  std::vector<KfState> eval_kf;
  std::vector<SE3> eval_Tbc;
  // ---- extrinsic self-calibration: VertexExtrinsic + EdgeExtrinsicPrior of LocalGPBA (src/Optimizer.cc:983-995, 1228-1240).
  // An extrinsic is a non-marginalized 6-dim vertex whose id follows every keyframe id (:986), so its Hessian block follows
  // the keyframes' (sparse_optimizer.cpp:166-190).  Here it occupies a 12-slot whose last six dimensions are padding (zero
  // Jacobian, unit diagonal in the reduced system, zero update): the solution is the one of the 6-dim block.
  std::vector<uint8_t> ext_free, ext_prior_on, ext_prior_active;
  std::vector<Quat> ext_prior_qinv;     // R_ini^-1 (EdgeExtrinsicPrior::R_, G2oTypes.h:474)
  std::vector<M3> ext_prior_info;       // MultiFrame::mRbc_ini_cov[c]
  std::vector<V3> ext_prior_err;
  std::vector<int> ext_h;               // Hessian index of camera c's extrinsic or -1
  std::vector<int> obs_slot3;           // Hpl slot of (extrinsic, landmark) or -1
  int numPosesKf = 0;
  std::vector<Snapshot> stack;

  // ---- LM state
  double currentLambda = -1, ni = 2;
  int nBad = 0, levenbergIterations = 0;
  double last_trial_chi2 = 0;
  // ---- per-stage wall time, field names of G2OBatchStatistics (g2o/core/batch_stats.h:39-78)
  double timeResiduals = 0, timeQuadraticForm = 0, timeSchurComplement = 0, timeLinearSolver = 0, timeUpdate = 0;
  static double now() {
#ifdef _OPENMP
    return omp_get_wtime();
#else
    return (double)clock() / CLOCKS_PER_SEC;
#endif
  }

  bool obs_is_gp(int64_t i) const { return rec_kf1[obs_rec[i]] >= 0; }
  int obs_dim(int64_t i) const { return obs_ur[i] >= 0 ? 3 : 2; }
  bool obs_has_kernel(int64_t i) const {
    if (obs_flags[i] & GPBA_OBS_NO_KERNEL) return false;
    return obs_dim(i) == 3 ? has_stereo_k : has_mono_k;
  }
  const Huber& obs_kernel(int64_t i) const { return obs_dim(i) == 3 ? hub_stereo : hub_mono; }

  void load(const gpba_problem* p) {
    n_cam = p->n_cam; n_kf = p->n_kf; n_pt = p->n_pt; n_rec = p->n_rec; n_obs = p->n_obs;
    n_prior = p->n_prior; n_velp = p->n_velp; bf = p->bf;
    cams.resize(n_cam); Tbc.resize(n_cam);
    for (int c = 0; c < n_cam; ++c) {
      cams[c] = {p->cam_intr[4 * c], p->cam_intr[4 * c + 1], p->cam_intr[4 * c + 2], p->cam_intr[4 * c + 3]};
      const double* q = p->cam_Tbc + 7 * c;
      Tbc[c].q = {q[0], q[1], q[2], q[3]};
      Tbc[c].t[0] = q[4]; Tbc[c].t[1] = q[5]; Tbc[c].t[2] = q[6];
    }
    kf.resize(n_kf); kf_fixed.assign(p->kf_fixed, p->kf_fixed + n_kf);
    for (int k = 0; k < n_kf; ++k) {
      const double* q = p->kf_pose + 7 * k;
      kf[k].Twb.q = {q[0], q[1], q[2], q[3]};
      kf[k].Twb.t[0] = q[4]; kf[k].Twb.t[1] = q[5]; kf[k].Twb.t[2] = q[6];
      for (int i = 0; i < 6; ++i) kf[k].vel[i] = p->kf_vel[6 * k + i];
      kf[k].time = p->kf_time[k];
    }
    pt.resize(n_pt);
    for (int i = 0; i < n_pt; ++i)
      for (int c = 0; c < 3; ++c) pt[i][c] = p->pt_xyz[3 * (size_t)i + c];
    rec_kf1.assign(p->rec_kf1, p->rec_kf1 + n_rec); rec_kf2.assign(p->rec_kf2, p->rec_kf2 + n_rec);
    rec_cam.assign(p->rec_cam, p->rec_cam + n_rec); rec_t.assign(p->rec_t, p->rec_t + n_rec);
    obs_u.assign(p->obs_u, p->obs_u + n_obs); obs_v.assign(p->obs_v, p->obs_v + n_obs);
    if (p->obs_ur) obs_ur.assign(p->obs_ur, p->obs_ur + n_obs); else obs_ur.assign(n_obs, -1.0);
    obs_w.assign(p->obs_inv_sigma2, p->obs_inv_sigma2 + n_obs);
    obs_rec.assign(p->obs_rec, p->obs_rec + n_obs); obs_pt.assign(p->obs_pt, p->obs_pt + n_obs);
    if (p->obs_flags) obs_flags.assign(p->obs_flags, p->obs_flags + n_obs); else obs_flags.assign(n_obs, 0);
    prior_kf1.assign(p->prior_kf1, p->prior_kf1 + n_prior); prior_kf2.assign(p->prior_kf2, p->prior_kf2 + n_prior);
    velp_kf.assign(p->velp_kf, p->velp_kf + n_velp);
    G.set_diag(p->qc);
    has_mono_k = p->huber_mono > 0; has_stereo_k = p->huber_stereo > 0; has_prior_k = p->huber_prior > 0;
    if (has_mono_k) hub_mono.setDelta(p->huber_mono);
    if (has_stereo_k) hub_stereo.setDelta(p->huber_stereo);
    if (has_prior_k) hub_prior.setDelta(p->huber_prior);
    lambda_init = p->lambda_init; linear_solver = p->linear_solver;
    obs_err.assign((size_t)n_obs * 3, 0.0); prior_err.assign((size_t)n_prior * 12, 0.0); velp_err.assign(n_velp, 0.0);
    ext_free.assign(n_cam, 0); ext_prior_on.assign(n_cam, 0); ext_prior_active.assign(n_cam, 0); ext_h.assign(n_cam, -1);
    ext_prior_qinv.assign(n_cam, Quat{0, 0, 0, 1}); ext_prior_info.assign(n_cam, M3::Zero()); ext_prior_err.assign(n_cam, V3());
  }

  // ------------------------------------------------------------------ structure
  void build_structure() {
    // initializeOptimization(level 0): an edge is active iff level matches and !allVerticesFixed
    // (points are never fixed, so every level-0 reprojection edge is active); a vertex is active iff
    // it has at least one such edge (sparse_optimizer.cpp:218-240).
    obs_active.assign(n_obs, 0); prior_active.assign(n_prior, 0); velp_active.assign(n_velp, 0);
    std::vector<char> kf_act(n_kf, 0), pt_act(n_pt, 0);
    active_obs.clear();
    for (int64_t i = 0; i < n_obs; ++i) {
      if (obs_flags[i] & GPBA_OBS_LEVEL1) continue;
      obs_active[i] = 1; active_obs.push_back(i);
      int r = obs_rec[i];
      pt_act[obs_pt[i]] = 1;
      if (rec_kf1[r] >= 0) kf_act[rec_kf1[r]] = 1;
      kf_act[rec_kf2[r]] = 1;
    }
    for (int i = 0; i < n_prior; ++i)
      if (!(kf_fixed[prior_kf1[i]] && kf_fixed[prior_kf2[i]])) { prior_active[i] = 1; kf_act[prior_kf1[i]] = kf_act[prior_kf2[i]] = 1; }
    for (int i = 0; i < n_velp; ++i)
      if (!kf_fixed[velp_kf[i]]) { velp_active[i] = 1; kf_act[velp_kf[i]] = 1; }
    // buildIndexMapping: free non-marginalized (KFs, ascending id) first, then marginalized (points)
    kf_h.assign(n_kf, -1); pt_h.assign(n_pt, -1);
    numPoses = 0;
    for (int k = 0; k < n_kf; ++k)
      if (kf_act[k] && !kf_fixed[k]) kf_h[k] = numPoses++;
    numPosesKf = numPoses;
    // free extrinsic vertices follow the keyframes (their ids are larger, Optimizer.cc:986); active iff they carry an active
    // edge: the prior (active iff the vertex is not fixed) or an active EdgeMonoGPExtrinsic of their camera
    {
      std::vector<char> cam_act(n_cam, 0);
      for (int64_t i : active_obs) if (rec_kf1[obs_rec[i]] >= 0) cam_act[rec_cam[obs_rec[i]]] = 1;
      for (int c = 0; c < n_cam; ++c) {
        ext_prior_active[c] = ext_free[c] && ext_prior_on[c];
        ext_h[c] = (ext_free[c] && (ext_prior_on[c] || cam_act[c])) ? numPoses++ : -1;
      }
    }
    auto obs_ext = [&](int64_t i) { const int r = obs_rec[i]; return rec_kf1[r] >= 0 ? ext_h[rec_cam[r]] : -1; };
    numLandmarks = 0; lm_pt.clear();
    for (int i = 0; i < n_pt; ++i)
      if (pt_act[i]) { pt_h[i] = numLandmarks++; lm_pt.push_back(i); }

    // Hpp pattern: diagonals + pose-pose pairs of active edges (block_solver.hpp:181-254)
    std::set<std::pair<int, int>> pp, sch;  // stored (col,row) so iteration is column-major like SparseBlockMatrix
    for (int i = 0; i < numPoses; ++i) pp.insert({i, i});
    auto add_pair = [&](std::set<std::pair<int, int>>& s, int a, int b2) {
      if (a < 0 || b2 < 0) return;
      if (a > b2) std::swap(a, b2);
      s.insert({b2, a});
    };
    for (int i = 0; i < n_prior; ++i)
      if (prior_active[i]) add_pair(pp, kf_h[prior_kf1[i]], kf_h[prior_kf2[i]]);
    for (int64_t i : active_obs) {
      int r = obs_rec[i];
      if (rec_kf1[r] >= 0) add_pair(pp, kf_h[rec_kf1[r]], kf_h[rec_kf2[r]]);
      const int he = obs_ext(i);   // EdgeMonoGPExtrinsic: (kf1, ext), (kf2, ext) (base_multi_edge.hpp:60-90 maps every vertex pair)
      if (he >= 0) { add_pair(pp, kf_h[rec_kf1[r]], he); add_pair(pp, kf_h[rec_kf2[r]], he); }
    }
    // Hschur pattern = Hpp pattern U pairs of free poses attached to ANY edge (active or not) of an
    // active landmark (block_solver.hpp:262-288 walks v->edges()).
    sch = pp;
    {
      std::vector<std::vector<int>> lm_poses(numLandmarks);
      for (int64_t i = 0; i < n_obs; ++i) {
        int l = pt_h[obs_pt[i]];
        if (l < 0) continue;
        int r = obs_rec[i];
        if (rec_kf1[r] >= 0 && kf_h[rec_kf1[r]] >= 0) lm_poses[l].push_back(kf_h[rec_kf1[r]]);
        if (kf_h[rec_kf2[r]] >= 0) lm_poses[l].push_back(kf_h[rec_kf2[r]]);
        if (obs_ext(i) >= 0) lm_poses[l].push_back(obs_ext(i));
      }
      for (int l = 0; l < numLandmarks; ++l) {
        std::vector<int>& v = lm_poses[l];
        std::sort(v.begin(), v.end());
        v.erase(std::unique(v.begin(), v.end()), v.end());
        for (size_t a = 0; a < v.size(); ++a)
          for (size_t c = a; c < v.size(); ++c) sch.insert({v[c], v[a]});
      }
    }
    auto finalize = [](const std::set<std::pair<int, int>>& s, std::vector<std::pair<int, int>>& rc,
                       std::map<std::pair<int, int>, int>& idx, std::vector<M12>& val) {
      rc.clear(); idx.clear();
      for (auto& cr : s) { idx[{cr.second, cr.first}] = (int)rc.size(); rc.push_back({cr.second, cr.first}); }
      val.assign(rc.size(), M12::Zero());
    };
    finalize(pp, hpp_rc, hpp_idx, hpp);
    finalize(sch, hs_rc, hs_idx, hs);
    hs_row_cols.assign(numPoses, {}); hs_row_idx.assign(numPoses, {});
    {
      std::vector<std::pair<std::pair<int, int>, int>> tmp;
      for (size_t i = 0; i < hs_rc.size(); ++i) tmp.push_back({hs_rc[i], (int)i});
      std::sort(tmp.begin(), tmp.end());
      for (auto& t : tmp) { hs_row_cols[t.first.first].push_back(t.first.second); hs_row_idx[t.first.first].push_back(t.second); }
    }
    // Hpl blocks: active edges only (block_solver.hpp:206-254), CSR by landmark sorted by pose
    {
      std::vector<std::vector<int>> lm_poses(numLandmarks);
      for (int64_t i : active_obs) {
        int l = pt_h[obs_pt[i]];
        int r = obs_rec[i];
        if (rec_kf1[r] >= 0 && kf_h[rec_kf1[r]] >= 0) lm_poses[l].push_back(kf_h[rec_kf1[r]]);
        if (kf_h[rec_kf2[r]] >= 0) lm_poses[l].push_back(kf_h[rec_kf2[r]]);
        if (obs_ext(i) >= 0) lm_poses[l].push_back(obs_ext(i));
      }
      lm_begin.assign(numLandmarks + 1, 0); hpl_pose.clear();
      for (int l = 0; l < numLandmarks; ++l) {
        std::vector<int>& v = lm_poses[l];
        std::sort(v.begin(), v.end());
        v.erase(std::unique(v.begin(), v.end()), v.end());
        lm_begin[l] = (int64_t)hpl_pose.size();
        hpl_pose.insert(hpl_pose.end(), v.begin(), v.end());
      }
      lm_begin[numLandmarks] = (int64_t)hpl_pose.size();
      hpl.assign(hpl_pose.size(), M12x3::Zero());
      obs_slot1.assign(n_obs, -1); obs_slot2.assign(n_obs, -1); obs_slot3.assign(n_obs, -1);
      for (int64_t i : active_obs) {
        int l = pt_h[obs_pt[i]];
        int r = obs_rec[i];
        auto slot_h = [&](int h) -> int {
          if (h < 0) return -1;
          const int* bgn = &hpl_pose[lm_begin[l]];
          const int* end = &hpl_pose[0] + lm_begin[l + 1];
          return (int)(lm_begin[l] + (std::lower_bound(bgn, end, h) - bgn));
        };
        auto slot = [&](int kfi) -> int { return kfi < 0 ? -1 : slot_h(kf_h[kfi]); };
        obs_slot1[i] = slot(rec_kf1[r]);
        obs_slot2[i] = slot(rec_kf2[r]);
        obs_slot3[i] = slot_h(obs_ext(i));
      }
    }
    hll.assign(numLandmarks, M3::Zero()); dinv.assign(numLandmarks, M3::Zero());
    size_t nvec = (size_t)numPoses * 12 + (size_t)numLandmarks * 3;
    b.assign(nvec, 0.0); x.assign(nvec, 0.0); bschur.assign((size_t)numPoses * 12, 0.0); coeff.assign(nvec, 0.0);
    sparse.analyzed = false;
    structure_ok = true; system_ok = false;
  }

  // ------------------------------------------------------------------ errors
  void compute_obs_error(int64_t i) {
    int r = obs_rec[i];
    bool gp = rec_kf1[r] >= 0;
    double obs[3] = {obs_u[i], obs_v[i], obs_ur[i]};
    reproj_error(G, gp, obs_dim(i), gp ? &kf[rec_kf1[r]] : nullptr, kf[rec_kf2[r]], rec_t[r], Tbc[rec_cam[r]],
                 cams[rec_cam[r]], bf, pt[obs_pt[i]], obs, &obs_err[(size_t)i * 3]);
  }
  double obs_chi2(int64_t i) const {  // BaseEdge::chi2 with Omega = I * invSigma2
    const double* e = &obs_err[(size_t)i * 3];
    double w = obs_w[i];
    double s = e[0] * (w * e[0]) + e[1] * (w * e[1]);
    if (obs_ur[i] >= 0) s += e[2] * (w * e[2]);
    return s;
  }
  M12 prior_info(int i) const { return G.QiInv(kf[prior_kf2[i]].time - kf[prior_kf1[i]].time); }
  double prior_chi2(int i) const {
    M12 O = prior_info(i);
    const double* e = &prior_err[(size_t)i * 12];
    double s = 0;
    for (int r = 0; r < 12; ++r) {
      double t = 0;
      for (int c = 0; c < 12; ++c) t += O(r, c) * e[c];
      s += e[r] * t;
    }
    return s;
  }
  double velp_chi2(int i) const { return velp_err[i] * (G.QcInv(2, 2) * velp_err[i]); }
  double ext_prior_chi2(int c) const {
    const V3& e = ext_prior_err[c];
    const M3& O = ext_prior_info[c];
    double s = 0;
    for (int r = 0; r < 3; ++r) { double t = 0; for (int k = 0; k < 3; ++k) t += O(r, k) * e[k]; s += e[r] * t; }
    return s;
  }

  void compute_errors() {  // SparseOptimizer::computeActiveErrors
    const double t_res = now();
    eval_kf = kf; eval_Tbc = Tbc;   // the state the stored errors belong to (stale-error quirk: may differ from the estimate)
    for (int i = 0; i < n_velp; ++i)
      if (velp_active[i]) velp_err[i] = kf[velp_kf[i]].vel[2];
    for (int i = 0; i < n_prior; ++i)
      if (prior_active[i]) prior_error(kf[prior_kf1[i]], kf[prior_kf2[i]], &prior_err[(size_t)i * 12]);
    for (int c = 0; c < n_cam; ++c)
      if (ext_prior_active[c]) ext_prior_err[c] = ext_prior_error(ext_prior_qinv[c], Tbc[c]);   // EdgeExtrinsicPrior::computeError
    const int64_t na = (int64_t)active_obs.size();
#pragma omp parallel for schedule(static) num_threads(threads) if (threads > 1)
    for (int64_t k = 0; k < na; ++k) compute_obs_error(active_obs[k]);
    timeResiduals += now() - t_res;
  }
  double robust_chi2() const {  // SparseOptimizer::activeRobustChi2
    double chi = 0.0, rho[3];
    for (int i = 0; i < n_velp; ++i)
      if (velp_active[i]) chi += velp_chi2(i);
    for (int i = 0; i < n_prior; ++i)
      if (prior_active[i]) {
        if (has_prior_k) { hub_prior.robustify(prior_chi2(i), rho); chi += rho[0]; }
        else chi += prior_chi2(i);
      }
    for (int c = 0; c < n_cam; ++c)
      if (ext_prior_active[c]) chi += ext_prior_chi2(c);   // no robust kernel (Optimizer.cc:990-994)
    for (int64_t i : active_obs) {
      if (obs_has_kernel(i)) { obs_kernel(i).robustify(obs_chi2(i), rho); chi += rho[0]; }
      else chi += obs_chi2(i);
    }
    return chi;
  }

  // ------------------------------------------------------------------ buildSystem
  int hpp_at(int r, int c) const { return hpp_idx.find({r, c})->second; }
  M12& hpp_block(int r, int c) { return hpp[hpp_at(r, c)]; }
  const M12& hpp_block(int r, int c) const { return hpp[hpp_at(r, c)]; }

  template <int D>
  static void add_AtOB(M12& H, const double* A, const double* B, const double* O /* D x D */, bool transposed) {
    // H += A^T O B   (A, B: D x 12 row-major); transposed: H += (A^T O B)^T  (base_multi_edge.hpp:204-208)
    double AtO[12 * D];
    for (int a = 0; a < 12; ++a)
      for (int d = 0; d < D; ++d) {
        double s = 0;
        for (int e = 0; e < D; ++e) s += A[e * 12 + a] * O[e * D + d];
        AtO[a * D + d] = s;
      }
    for (int a = 0; a < 12; ++a)
      for (int c = 0; c < 12; ++c) {
        double s = 0;
        for (int d = 0; d < D; ++d) s += AtO[a * D + d] * B[d * 12 + c];
        if (transposed) H(c, a) += s; else H(a, c) += s;
      }
  }

  // pose-side accumulators of one thread (Hpp values + pose part of b)
  struct PoseAcc { std::vector<M12> hpp; std::vector<double> bp; };

  void accumulate_obs(int64_t i, std::vector<M12>& Hpp, double* bp) {
    int r = obs_rec[i];
    const bool gp = rec_kf1[r] >= 0;
    const int dim = obs_dim(i);
    double J1[36], J2[36], Jp[9], Je[18], J3[36];
    const int h3 = gp ? ext_h[rec_cam[r]] : -1;
    reproj_jacobian(G, gp, dim, gp ? &kf[rec_kf1[r]] : nullptr, kf[rec_kf2[r]], rec_t[r], Tbc[rec_cam[r]],
                    cams[rec_cam[r]], bf, pt[obs_pt[i]], J1, J2, Jp, h3 >= 0 ? Je : nullptr);
    if (h3 >= 0)   // the extrinsic's 12-slot: [J_ext | 0]
      for (int d = 0; d < dim; ++d)
        for (int c = 0; c < 12; ++c) J3[d * 12 + c] = c < 6 ? Je[d * 6 + c] : 0.0;
    // constructQuadraticForm (base_multi_edge.hpp:36-48 / base_binary_edge.hpp:55-120)
    const double* e = &obs_err[(size_t)i * 3];
    double w = obs_w[i];
    double rho1 = 1.0;
    if (obs_has_kernel(i)) { double rho[3]; obs_kernel(i).robustify(obs_chi2(i), rho); rho1 = rho[1]; }
    double omega_r[3];  // -Omega * e * rho'
    for (int d = 0; d < dim; ++d) omega_r[d] = -(w * e[d]) * rho1;
    const double wr = rho1 * w;  // robustInformation = rho' * Omega
    const int h1 = gp ? kf_h[rec_kf1[r]] : -1, h2 = kf_h[rec_kf2[r]];
    const int l = pt_h[obs_pt[i]];
    double O[9] = {wr, 0, 0, 0, wr, 0, 0, 0, wr};
    double O2[4] = {wr, 0, 0, wr};
    const double* Om = dim == 3 ? O : O2;
    auto AtOB = [&](M12& H, const double* A, const double* B2, bool tr) {
      if (dim == 3) add_AtOB<3>(H, A, B2, Om, tr); else add_AtOB<2>(H, A, B2, Om, tr);
    };
    auto add_b = [&](double* bb, const double* A, int n, int ld) {
      for (int a = 0; a < n; ++a) {
        double s = 0;
        for (int d = 0; d < dim; ++d) s += A[d * ld + a] * omega_r[d];
        bb[a] += s;
      }
    };
    auto add_hpl = [&](int slot, const double* A) {
      M12x3& H = hpl[slot];
      for (int a = 0; a < 12; ++a)
        for (int c = 0; c < 3; ++c) {
          double s = 0;
          for (int d = 0; d < dim; ++d) s += (A[d * 12 + a] * wr) * Jp[d * 3 + c];
          H(a, c) += s;
        }
    };
    if (h1 >= 0) {
      AtOB(Hpp[hpp_at(h1, h1)], J1, J1, false);
      add_b(bp + (size_t)h1 * 12, J1, 12, 12);
      if (h2 >= 0) {
        if (h1 <= h2) AtOB(Hpp[hpp_at(h1, h2)], J1, J2, false);
        else AtOB(Hpp[hpp_at(h2, h1)], J1, J2, true);
      }
      add_hpl(obs_slot1[i], J1);
    }
    if (h2 >= 0) {
      AtOB(Hpp[hpp_at(h2, h2)], J2, J2, false);
      add_b(bp + (size_t)h2 * 12, J2, 12, 12);
      add_hpl(obs_slot2[i], J2);
    }
    if (h3 >= 0) {   // the extrinsic follows every keyframe: (kf, ext) blocks are stored as they are
      AtOB(Hpp[hpp_at(h3, h3)], J3, J3, false);
      add_b(bp + (size_t)h3 * 12, J3, 12, 12);
      if (h1 >= 0) AtOB(Hpp[hpp_at(h1, h3)], J1, J3, false);
      if (h2 >= 0) AtOB(Hpp[hpp_at(h2, h3)], J2, J3, false);
      add_hpl(obs_slot3[i], J3);
    }
    M3& Hl = hll[l];
    for (int a = 0; a < 3; ++a)
      for (int c = 0; c < 3; ++c) {
        double s = 0;
        for (int d = 0; d < dim; ++d) s += (Jp[d * 3 + a] * wr) * Jp[d * 3 + c];
        Hl(a, c) += s;
      }
    add_b(&b[(size_t)numPoses * 12 + (size_t)l * 3], Jp, 3, 3);
  }

  void accumulate_priors() {
    for (int i = 0; i < n_velp; ++i) {
      if (!velp_active[i]) continue;
      int h = kf_h[velp_kf[i]];
      double O = G.QcInv(2, 2);
      hpp_block(h, h)(8, 8) += O;           // J = [0_6 | 0 0 1 0 0 0]  (G2oTypes.h:509-513)
      b[(size_t)h * 12 + 8] += -(O * velp_err[i]);
    }
    for (int c = 0; c < n_cam; ++c) {   // EdgeExtrinsicPrior (G2oTypes.h:470-494): J = [0 | Jr(e)^-1], no kernel
      if (!ext_prior_active[c]) continue;
      const int h = ext_h[c];
      const M3 Ji3 = inverse3_cofactor(RightJacobianSO3_orb(ext_prior_err[c]));
      const M3& O = ext_prior_info[c];
      const M3 JtO = transpose(Ji3) * O;
      const M3 H3 = JtO * Ji3;
      M12& B = hpp_block(h, h);
      for (int a = 0; a < 3; ++a)
        for (int k = 0; k < 3; ++k) B(3 + a, 3 + k) += H3(a, k);
      for (int a = 0; a < 3; ++a) {
        double t = 0;
        for (int k = 0; k < 3; ++k) t += JtO(a, k) * ext_prior_err[c][k];
        b[(size_t)h * 12 + 3 + a] += -t;
      }
    }
    for (int i = 0; i < n_prior; ++i) {
      if (!prior_active[i]) continue;
      M12 Ji, Jj;
      prior_jacobian(kf[prior_kf1[i]], kf[prior_kf2[i]], &Ji, &Jj);
      M12 O = prior_info(i);
      const double* e = &prior_err[(size_t)i * 12];
      V12 ev;
      for (int k = 0; k < 12; ++k) ev[k] = e[k];
      double rho1 = 1.0;
      if (has_prior_k) { double rho[3]; hub_prior.robustify(prior_chi2(i), rho); rho1 = rho[1]; }
      V12 omega_r = (-(O * ev)) * rho1;
      M12 Or = rho1 * O;
      int hi = kf_h[prior_kf1[i]], hj = kf_h[prior_kf2[i]];
      M12 AtO = transpose(Ji) * Or;   // base_binary_edge.hpp:80-105
      if (hi >= 0) {
        hpp_block(hi, hi) = hpp_block(hi, hi) + AtO * Ji;
        V12 bi = transpose(Ji) * omega_r;
        for (int k = 0; k < 12; ++k) b[(size_t)hi * 12 + k] += bi[k];
      }
      if (hj >= 0) {
        M12 BtO = transpose(Jj) * Or;
        hpp_block(hj, hj) = hpp_block(hj, hj) + BtO * Jj;
        V12 bj = transpose(Jj) * omega_r;
        for (int k = 0; k < 12; ++k) b[(size_t)hj * 12 + k] += bj[k];
        if (hi >= 0) {
          M12 off = AtO * Jj;
          if (hi <= hj) hpp_block(hi, hj) = hpp_block(hi, hj) + off;
          else hpp_block(hj, hi) = hpp_block(hj, hi) + transpose(off);
        }
      }
    }
  }

  void build_system() {
    const double t_qf = now();
    std::fill(b.begin(), b.end(), 0.0);
    for (auto& m : hpp) m = M12::Zero();
    for (auto& m : hll) m = M3::Zero();
    for (auto& m : hpl) m = M12x3::Zero();
    accumulate_priors();
    if (threads <= 1) {
      for (int64_t i : active_obs) accumulate_obs(i, hpp, b.data());
    } else {
      // OpenMP variant: threads own disjoint landmark ranges (so Hll/Hpl/b_l need no locks, the role
      // of g2o's per-vertex lockQuadraticForm), pose-side sums go to per-thread copies reduced at the end.
      std::vector<std::vector<int64_t>> by_lm(numLandmarks);
      for (int64_t i : active_obs) by_lm[pt_h[obs_pt[i]]].push_back(i);
      std::vector<PoseAcc> acc(threads);
#pragma omp parallel num_threads(threads)
      {
#ifdef _OPENMP
        int tid = omp_get_thread_num();
#else
        int tid = 0;
#endif
        PoseAcc& A = acc[tid];
        A.hpp.assign(hpp.size(), M12::Zero());
        A.bp.assign((size_t)numPoses * 12, 0.0);
#pragma omp for schedule(static)
        for (int l = 0; l < numLandmarks; ++l)
          for (int64_t i : by_lm[l]) accumulate_obs(i, A.hpp, A.bp.data());
      }
      for (int t = 0; t < threads; ++t) {
        if (acc[t].hpp.empty()) continue;
        for (size_t k = 0; k < hpp.size(); ++k) hpp[k] = hpp[k] + acc[t].hpp[k];
        for (size_t k = 0; k < acc[t].bp.size(); ++k) b[k] += acc[t].bp[k];
      }
    }
    system_ok = true;
    timeQuadraticForm += now() - t_qf;
  }

  void set_lambda(double lambda, bool backup) {
    if (backup) { diag_backup_pose.resize((size_t)numPoses * 12); diag_backup_lm.resize((size_t)numLandmarks * 3); }
    for (int i = 0; i < numPoses; ++i) {
      M12& B = hpp_block(i, i);
      for (int k = 0; k < 12; ++k) { if (backup) diag_backup_pose[(size_t)i * 12 + k] = B(k, k); B(k, k) += lambda; }
    }
    for (int i = 0; i < numLandmarks; ++i)
      for (int k = 0; k < 3; ++k) { if (backup) diag_backup_lm[(size_t)i * 3 + k] = hll[i](k, k); hll[i](k, k) += lambda; }
  }
  void restore_diagonal() {
    for (int i = 0; i < numPoses; ++i) {
      M12& B = hpp_block(i, i);
      for (int k = 0; k < 12; ++k) B(k, k) = diag_backup_pose[(size_t)i * 12 + k];
    }
    for (int i = 0; i < numLandmarks; ++i)
      for (int k = 0; k < 3; ++k) hll[i](k, k) = diag_backup_lm[(size_t)i * 3 + k];
  }

  // ------------------------------------------------------------------ solve (block_solver.hpp:353-486)
  bool solve() {
    const size_t sizePoses = (size_t)numPoses * 12;
    for (auto& m : hs) m = M12::Zero();
    for (size_t i = 0; i < hpp_rc.size(); ++i) { M12& t = hs[hs_idx.find(hpp_rc[i])->second]; t = t + hpp[i]; }  // _Hpp->add(_Hschur)
    for (int h = numPosesKf; h < numPoses; ++h) {   // padding of the extrinsics' 12-slots: identity rows, zero right-hand side
      M12& t = hs[hs_idx.find({h, h})->second];
      for (int k = 6; k < 12; ++k) t(k, k) += 1.0;
    }
    std::fill(coeff.begin(), coeff.begin() + sizePoses, 0.0);
    const double t_schur = now();
    // G2O_OPENMP build of the reference: "#pragma omp parallel for schedule(dynamic, 10)" over the landmarks with one
    // mutex per pose row (block_solver.hpp:378-380, 397-399).  threads == 1 is the reference's actual configuration.
#ifdef _OPENMP
    if (threads > 1 && (int)row_locks.size() != numPoses) {
      for (auto& lk : row_locks) omp_destroy_lock(&lk);
      row_locks.resize(numPoses);
      for (auto& lk : row_locks) omp_init_lock(&lk);
    }
#pragma omp parallel for schedule(dynamic, 10) num_threads(threads) if (threads > 1)
#endif
    for (int l = 0; l < numLandmarks; ++l) {
      M3 Dinv = inverse<3>(hll[l]);
      dinv[l] = Dinv;
      V3 db;
      for (int j = 0; j < 3; ++j) db[j] = b[sizePoses + (size_t)l * 3 + j];
      db = Dinv * db;
      for (int64_t s = lm_begin[l]; s < lm_begin[l + 1]; ++s) {
        const int i1 = hpl_pose[s];
        const M12x3& Bi = hpl[s];
        M12x3 BDinv = Bi * Dinv;
        V12 Bb = Bi * db;
#ifdef _OPENMP
        if (threads > 1) omp_set_lock(&row_locks[i1]);
#endif
        for (int k = 0; k < 12; ++k) coeff[(size_t)i1 * 12 + k] += Bb[k];
        size_t it = 0;
        const std::vector<int>& cols = hs_row_cols[i1];
        for (int64_t s2 = s; s2 < lm_begin[l + 1]; ++s2) {
          const int i2 = hpl_pose[s2];
          while (cols[it] < i2) ++it;
          M12& H = hs[hs_row_idx[i1][it]];
          const M12x3& Bj = hpl[s2];
          for (int r = 0; r < 12; ++r)
            for (int c = 0; c < 12; ++c) {
              double t = 0;
              for (int k = 0; k < 3; ++k) t += BDinv(r, k) * Bj(c, k);
              H(r, c) -= t;
            }
        }
#ifdef _OPENMP
        if (threads > 1) omp_unset_lock(&row_locks[i1]);
#endif
      }
    }
    for (size_t i = 0; i < sizePoses; ++i) bschur[i] = b[i] - coeff[i];
    timeSchurComplement += now() - t_schur;
    const double t_lin = now();
    bool ok;
    if (linear_solver == GPBA_SOLVER_DENSE_CHOL) {
      int n = (int)sizePoses;
      std::vector<double> H((size_t)n * n, 0.0);
      for (size_t k = 0; k < hs_rc.size(); ++k) {
        int r0 = hs_rc[k].first * 12, c0 = hs_rc[k].second * 12;
        for (int r = 0; r < 12; ++r)
          for (int c = 0; c < 12; ++c) {
            H[(size_t)(r0 + r) * n + c0 + c] = hs[k](r, c);
            if (r0 != c0) H[(size_t)(c0 + c) * n + r0 + r] = hs[k](r, c);
          }
      }
      ok = ldlt_dense(n, H, bschur.data(), x.data());
    } else {
      if (!sparse.analyzed) sparse.analyze(numPoses, hs_rc);
      ok = sparse.factor_solve(hs_rc, hs, bschur.data(), x.data());
    }
    timeLinearSolver += now() - t_lin;
    if (!ok) return false;
    // landmarks: xl = Dinv * (bl - Hpl^T xp)   (block_solver.hpp:459-483; under G2O_OPENMP the two products run in parallel, sparse_block_matrix_ccs.h:117, sparse_block_matrix_diagonal.h:90)
#pragma omp parallel for schedule(static) num_threads(threads) if (threads > 1)
    for (int l = 0; l < numLandmarks; ++l) {
      V3 cl;
      for (int j = 0; j < 3; ++j) cl[j] = b[sizePoses + (size_t)l * 3 + j];
      for (int64_t s = lm_begin[l]; s < lm_begin[l + 1]; ++s) {
        const double* xp = &x[(size_t)hpl_pose[s] * 12];
        const M12x3& B = hpl[s];
        for (int c = 0; c < 3; ++c) {
          double t = 0;
          for (int r = 0; r < 12; ++r) t += B(r, c) * xp[r];
          cl[c] -= t;
        }
      }
      V3 xl = dinv[l] * cl;
      for (int j = 0; j < 3; ++j) x[sizePoses + (size_t)l * 3 + j] = xl[j];
    }
    return true;
  }

  // ------------------------------------------------------------------ state
  void oplus(const double* upd) {  // SparseOptimizer::update
    const double t_up = now();
    for (int k = 0; k < n_kf; ++k) {
      if (kf_h[k] < 0) continue;
      const double* u = upd + (size_t)kf_h[k] * 12;
      V6 du;
      for (int i = 0; i < 6; ++i) du[i] = u[i];
      kf[k].Twb = se3_mul(kf[k].Twb, se3_exp(du));  // PoseVelocity::Update, G2oTypes.cc:41-46
      for (int i = 0; i < 6; ++i) kf[k].vel[i] += u[6 + i];
    }
    for (int c = 0; c < n_cam; ++c) {   // VertexExtrinsic::oplusImpl (G2oTypes.h:98-100)
      if (ext_h[c] < 0) continue;
      const double* u = upd + (size_t)ext_h[c] * 12;
      V6 du;
      for (int i = 0; i < 6; ++i) du[i] = u[i];
      Tbc[c] = se3_mul(Tbc[c], se3_exp(du));
    }
    const double* ul = upd + (size_t)numPoses * 12;
    for (int l = 0; l < numLandmarks; ++l)
      for (int c = 0; c < 3; ++c) pt[lm_pt[l]][c] += ul[(size_t)l * 3 + c];  // VertexSBAPointXYZ::oplusImpl
    timeUpdate += now() - t_up;
  }
  void push() { stack.push_back({kf, pt, Tbc}); }
  void pop() { kf = stack.back().kf; pt = stack.back().pt; Tbc = stack.back().Tbc; stack.pop_back(); }
  void discard_top() { stack.pop_back(); }

  // ------------------------------------------------------------------ LM (optimization_algorithm_levenberg.cpp)
  double compute_lambda_init(const gpba_lm_params& P) {
    if (lambda_init > 0) return lambda_init;
    double maxDiagonal = 0;
    for (int i = 0; i < numPoses; ++i) {
      const M12& B = hpp_block(i, i);
      for (int k = 0; k < 12; ++k) maxDiagonal = std::max(std::fabs(B(k, k)), maxDiagonal);
    }
    for (int l = 0; l < numLandmarks; ++l)
      for (int k = 0; k < 3; ++k) maxDiagonal = std::max(std::fabs(hll[l](k, k)), maxDiagonal);
    return P.tau * maxDiagonal;
  }
  double compute_scale() const {
    double scale = 0;
    for (size_t j = 0; j < x.size(); ++j) scale += x[j] * (currentLambda * x[j] + b[j]);
    return scale;
  }
  int lm_solve(int iteration, const gpba_lm_params& P, const volatile unsigned char* stop, gpba_lm_trace* tr) {
    // buildStructure belongs to the first solve() (optimization_algorithm_levenberg.cpp:65-70); a structure built by the
    // caller right before (same levels, nothing changed since) is the same structure and is used as it is
    if (iteration == 0 && !structure_fresh) build_structure();
    structure_fresh = false;
    compute_errors();
    double currentChi = robust_chi2();
    double tempChi = currentChi;
    double iniChi = currentChi;
    build_system();
    if (iteration == 0) { currentLambda = compute_lambda_init(P); ni = 2; nBad = 0; }
    double rho = 0;
    int& qmax = levenbergIterations;
    qmax = 0;
    do {
      push();
      set_lambda(currentLambda, true);
      bool ok2 = solve();
      oplus(x.data());
      restore_diagonal();
      compute_errors();
      tempChi = robust_chi2();
      if (!ok2) tempChi = std::numeric_limits<double>::max();
      last_trial_chi2 = tempChi;
      rho = (currentChi - tempChi);
      double scale = compute_scale();
      scale += 1e-3;
      rho /= scale;
      if (rho > 0 && std::isfinite(tempChi)) {
        double alpha = 1. - std::pow((2 * rho - 1), 3);
        alpha = (std::min)(alpha, P.good_step_upper);
        double scaleFactor = (std::max)(P.good_step_lower, alpha);
        currentLambda *= scaleFactor;
        ni = 2;
        currentChi = tempChi;
        discard_top();
      } else {
        currentLambda *= ni;
        ni *= 2;
        pop();
      }
      qmax++;
    } while (rho < 0 && qmax < P.max_trials_after_failure && !(stop && *stop));
    if (tr && iteration < GPBA_MAX_ITERS) {
      tr->levenberg_iterations[iteration] = qmax;
      tr->chi2_before[iteration] = iniChi;
      tr->chi2_after[iteration] = currentChi;
      tr->lambda[iteration] = currentLambda;
      tr->total_trials += qmax;
      tr->last_trial_chi2 = last_trial_chi2;
    }
    if (qmax == P.max_trials_after_failure || rho == 0) return GPBA_TERMINATE;
    if ((iniChi - currentChi) * 1e3 < iniChi) nBad++; else nBad = 0;
    if (nBad >= 3) return GPBA_TERMINATE;
    return GPBA_RESULT_OK;
  }
  int optimize(int iterations, const volatile unsigned char* stop, const gpba_lm_params& P, gpba_lm_trace* tr) {
    if (tr) { std::memset(tr, 0, sizeof(*tr)); tr->result = GPBA_RESULT_OK; }
    int cj = 0;
    bool ok = true;
    int result = GPBA_RESULT_OK;
    for (int i = 0; i < iterations && !(stop && *stop) && ok; i++) {
      result = lm_solve(i, P, stop, tr);
      ok = (result == GPBA_RESULT_OK);
      ++cj;
    }
    if (tr) { tr->n_iters = cj; tr->result = result; }
    return cj;
  }

  // ------------------------------------------------------------------ LocalGPBA inlier check (Optimizer.cc:1263-1348)
  bool obs_depth_positive(int64_t i) const {
    int r = obs_rec[i];
    const SE3& T = Tbc[rec_cam[r]];
    bool ok = depth_positive(kf[rec_kf2[r]].Twb, T, pt[obs_pt[i]]);
    if (rec_kf1[r] >= 0) ok = depth_positive(kf[rec_kf1[r]].Twb, T, pt[obs_pt[i]]) && ok;
    return ok;
  }
  void outlier_flags(const gpba_thresholds& th, uint8_t* flags) const {
    for (int64_t i = 0; i < n_obs; ++i) {
      double c2 = obs_chi2(i);
      bool out;
      if (obs_ur[i] >= 0) out = c2 > th.chi2_stereo;  // EdgeStereo / EdgeStereoGP: chi2 only (:1283-1296, 1317-1330)
      else {
        bool close = obs_flags[i] & GPBA_OBS_CLOSE;
        out = (c2 > th.chi2_mono && !close) || (c2 > th.chi2_mono_close && close) || !obs_depth_positive(i);
      }
      flags[i] = out ? 1 : 0;
    }
  }
  void rejection_rounds(int n_rounds, int iters, const gpba_thresholds& th, const gpba_lm_params& P, uint8_t* flags,
                        gpba_lm_trace* traces) {
    std::vector<uint8_t> fl(n_obs, 0);
    for (int64_t i = 0; i < n_obs; ++i) fl[i] = (obs_flags[i] & GPBA_OBS_LEVEL1) ? 1 : 0;
    for (int it = 0; it < n_rounds; ++it) {
      optimize(iters, nullptr, P, traces ? &traces[it] : nullptr);  // initializeOptimization(0) + optimize
      for (int64_t i = 0; i < n_obs; ++i)
        if (fl[i]) compute_obs_error(i);  // "if (mvbOutlier[idx]) e->computeError()"  (Optimizer.cc:591-592)
      outlier_flags(th, fl.data());
      for (int64_t i = 0; i < n_obs; ++i) {
        if (fl[i]) obs_flags[i] |= GPBA_OBS_LEVEL1; else obs_flags[i] &= ~GPBA_OBS_LEVEL1;
        if (it == 2) obs_flags[i] |= GPBA_OBS_NO_KERNEL;  // "if (it==2) e->setRobustKernel(0)"
      }
    }
    std::memcpy(flags, fl.data(), (size_t)n_obs);
  }
};

}  // namespace

// ================================================================================= C API
#define ORA(h) reinterpret_cast<Oracle*>(h)
extern "C" {

void* oracle_create(const gpba_problem* p) {
  Oracle* o = new Oracle();
  o->load(p);
  return o;
}
void oracle_destroy(void* h) { delete ORA(h); }
// Re-load the estimates only (same graph): lets the benchmark repeat optimize() from the same start on one instance.
void oracle_reset_state(void* h, const double* kf_pose, const double* kf_vel, const double* pt_xyz) {
  Oracle* o = ORA(h);
  for (int k = 0; k < o->n_kf; ++k) {
    if (kf_pose) {
      const double* q = kf_pose + 7 * k;
      o->kf[k].Twb.q = {q[0], q[1], q[2], q[3]};
      o->kf[k].Twb.t[0] = q[4]; o->kf[k].Twb.t[1] = q[5]; o->kf[k].Twb.t[2] = q[6];
    }
    if (kf_vel) for (int i = 0; i < 6; ++i) o->kf[k].vel[i] = kf_vel[6 * k + i];
  }
  if (pt_xyz) for (int i = 0; i < o->n_pt; ++i) for (int c = 0; c < 3; ++c) o->pt[i][c] = pt_xyz[3 * (size_t)i + c];
  o->stack.clear();
}
// LocalGPBA's extrinsic vertices (src/Optimizer.cc:983-995): which are free (setFixed(false), :1236) and their
// EdgeExtrinsicPrior (R_ini as quaternion xyzw, information 3 x 3 row-major); prior_q == NULL: no prior edges.
int oracle_set_extrinsics(void* h, const uint8_t* free_, const double* prior_q, const double* prior_info) {
  Oracle* o = ORA(h);
  o->structure_fresh = false;
  for (int c = 0; c < o->n_cam; ++c) {
    o->ext_free[c] = free_ ? free_[c] : 0;
    o->ext_prior_on[c] = prior_q != nullptr;
    if (prior_q) {
      o->ext_prior_qinv[c] = quat_inv(Quat{prior_q[4 * c], prior_q[4 * c + 1], prior_q[4 * c + 2], prior_q[4 * c + 3]});
      for (int r = 0; r < 3; ++r) for (int k = 0; k < 3; ++k) o->ext_prior_info[c](r, k) = prior_info[9 * c + 3 * r + k];
    }
  }
  return 0;
}
int oracle_get_extrinsics(void* h, double* Tbc7) {
  Oracle* o = ORA(h);
  for (int c = 0; c < o->n_cam; ++c) {
    double* q = Tbc7 + 7 * c;
    q[0] = o->Tbc[c].q.x; q[1] = o->Tbc[c].q.y; q[2] = o->Tbc[c].q.z; q[3] = o->Tbc[c].q.w;
    q[4] = o->Tbc[c].t[0]; q[5] = o->Tbc[c].t[1]; q[6] = o->Tbc[c].t[2];
  }
  return 0;
}
// cam_obs[c]++ for every EdgeMonoGPExtrinsic of camera c, whatever its level (src/Optimizer.cc:1129)
int oracle_count_camera_observations(void* h, int64_t* cam_obs) {
  Oracle* o = ORA(h);
  for (int c = 0; c < o->n_cam; ++c) cam_obs[c] = 0;
  for (int64_t i = 0; i < o->n_obs; ++i) if (o->rec_kf1[o->obs_rec[i]] >= 0) cam_obs[o->rec_cam[o->obs_rec[i]]]++;
  return 0;
}
// ---- essential-graph optimisation (oracle/pose_graph.h; src/Optimizer.cc:1434-1717)
int oracle_pose_graph_optimize(const gpba_pose_graph* g, int iters, const gpba_lm_params* params, double* sim3_out, gpba_lm_trace* trace) {
  PoseGraphOracle o;
  o.load(g);
  gpba_lm_params P;
  P.max_trials_after_failure = 10; P.tau = 1e-5; P.good_step_lower = 1. / 3.; P.good_step_upper = 2. / 3.; P.pcg_tolerance = 0; P.pcg_max_iterations = 0;
  if (params) P = *params;
  o.optimize(iters, P, trace, ldlt_dense);
  if (sim3_out) for (int i = 0; i < o.n; ++i) sim3_to8(o.S[i], sim3_out + 8 * i);
  return 0;
}
// src/Optimizer.cc:1687-1712: P <- correctedSwr.map(Srw.map(P)) with r the point's reference keyframe
int oracle_correct_points(int64_t n_pt, const double* xyz, const int32_t* ref_kf, const double* sim3_before, const double* sim3_after, double* xyz_out) {
  for (int64_t i = 0; i < n_pt; ++i) {
    const Sim3 Srw = sim3_from8(sim3_before + 8 * (size_t)ref_kf[i]);
    const Sim3 Swr = sim3_inv(sim3_from8(sim3_after + 8 * (size_t)ref_kf[i]));
    V3 p; p[0] = xyz[3 * i]; p[1] = xyz[3 * i + 1]; p[2] = xyz[3 * i + 2];
    const V3 q = sim3_map(Swr, sim3_map(Srw, p));
    xyz_out[3 * i] = q[0]; xyz_out[3 * i + 1] = q[1]; xyz_out[3 * i + 2] = q[2];
  }
  return 0;
}
void oracle_sim3_exp(const double* u7, double* S8) { sim3_to8(sim3_exp(u7), S8); }
void oracle_sim3_log(const double* S8, double* u7) { sim3_log(sim3_from8(S8), u7); }
void oracle_sim3_mul(const double* a, const double* b, double* o8) { sim3_to8(sim3_mul(sim3_from8(a), sim3_from8(b)), o8); }
void oracle_sim3_inv(const double* a, double* o8) { sim3_to8(sim3_inv(sim3_from8(a)), o8); }
void oracle_set_threads(void* h, int n) { ORA(h)->threads = n < 1 ? 1 : n; }
// seconds per stage since the last reset: timeResiduals, timeQuadraticForm (linearize + quadratic form), timeSchurComplement,
// timeLinearSolver, timeUpdate -- the G2OBatchStatistics fields (g2o/core/batch_stats.h:39-78)
void oracle_batch_stats(void* h, double out[5], int reset) {
  auto* o = ORA(h);
  out[0] = o->timeResiduals; out[1] = o->timeQuadraticForm; out[2] = o->timeSchurComplement; out[3] = o->timeLinearSolver; out[4] = o->timeUpdate;
  if (reset) o->timeResiduals = o->timeQuadraticForm = o->timeSchurComplement = o->timeLinearSolver = o->timeUpdate = 0;
}
void oracle_default_lm_params(gpba_lm_params* p) {
  p->max_trials_after_failure = 10; p->tau = 1e-5; p->good_step_lower = 1. / 3.; p->good_step_upper = 2. / 3.;
  p->pcg_tolerance = 1e-12; p->pcg_max_iterations = 2000;
}
int oracle_build_structure(void* h, gpba_structure_info* info) {
  Oracle* o = ORA(h);
  o->build_structure();
  o->structure_fresh = true;
  if (info) {
    info->n_free_kf = o->numPoses; info->n_active_pt = o->numLandmarks; info->n_active_obs = (int64_t)o->active_obs.size();
    info->n_hpl = (int64_t)o->hpl_pose.size(); info->n_hpp = (int)o->hpp_rc.size(); info->n_hschur = (int)o->hs_rc.size();
  }
  return 0;
}
static void copy_pattern(const std::vector<std::pair<int, int>>& rc, int32_t* rows, int32_t* cols) {
  for (size_t i = 0; i < rc.size(); ++i) { rows[i] = rc[i].first; cols[i] = rc[i].second; }
}
int oracle_get_hpp_pattern(void* h, int32_t* rows, int32_t* cols) { copy_pattern(ORA(h)->hpp_rc, rows, cols); return 0; }
int oracle_get_hschur_pattern(void* h, int32_t* rows, int32_t* cols) { copy_pattern(ORA(h)->hs_rc, rows, cols); return 0; }
int oracle_compute_errors(void* h, double* chi2) { ORA(h)->compute_errors(); if (chi2) *chi2 = ORA(h)->robust_chi2(); return 0; }
int oracle_build_system(void* h) { ORA(h)->build_system(); return 0; }
int oracle_set_lambda(void* h, double l, int backup) { ORA(h)->set_lambda(l, backup != 0); return 0; }
int oracle_restore_diagonal(void* h) { ORA(h)->restore_diagonal(); return 0; }
int oracle_solve(void* h, int* ok) { bool r = ORA(h)->solve(); if (ok) *ok = r ? 1 : 0; return 0; }
int oracle_vector_size(void* h, int64_t* n) { *n = (int64_t)ORA(h)->x.size(); return 0; }
int oracle_get_x(void* h, double* x) { std::memcpy(x, ORA(h)->x.data(), ORA(h)->x.size() * 8); return 0; }
int oracle_get_b(void* h, double* b) { std::memcpy(b, ORA(h)->b.data(), ORA(h)->b.size() * 8); return 0; }
int oracle_get_hpp(void* h, double* blocks) {
  Oracle* o = ORA(h);
  for (size_t i = 0; i < o->hpp.size(); ++i) std::memcpy(blocks + i * 144, o->hpp[i].a, 144 * 8);
  return 0;
}
int oracle_get_hschur(void* h, double* blocks, double* bschur) {
  Oracle* o = ORA(h);
  if (blocks) for (size_t i = 0; i < o->hs.size(); ++i) std::memcpy(blocks + i * 144, o->hs[i].a, 144 * 8);
  if (bschur) std::memcpy(bschur, o->bschur.data(), o->bschur.size() * 8);
  return 0;
}
int oracle_get_hll(void* h, double* blocks) {
  Oracle* o = ORA(h);
  for (size_t i = 0; i < o->hll.size(); ++i) std::memcpy(blocks + i * 9, o->hll[i].a, 9 * 8);
  return 0;
}
int oracle_get_hpl(void* h, int64_t* lm_begin, int32_t* pose, double* blocks) {
  Oracle* o = ORA(h);
  if (lm_begin) std::memcpy(lm_begin, o->lm_begin.data(), o->lm_begin.size() * 8);
  if (pose) std::memcpy(pose, o->hpl_pose.data(), o->hpl_pose.size() * 4);
  if (blocks) for (size_t i = 0; i < o->hpl.size(); ++i) std::memcpy(blocks + i * 36, o->hpl[i].a, 36 * 8);
  return 0;
}
int oracle_oplus(void* h, const double* x) { ORA(h)->oplus(x ? x : ORA(h)->x.data()); return 0; }
int oracle_push(void* h) { ORA(h)->push(); return 0; }
int oracle_pop(void* h) { ORA(h)->pop(); return 0; }
int oracle_discard_top(void* h) { ORA(h)->discard_top(); return 0; }
int oracle_optimize(void* h, int iters, const volatile unsigned char* stop, const gpba_lm_params* P, gpba_lm_trace* tr) {
  gpba_lm_params d;
  oracle_default_lm_params(&d);
  ORA(h)->optimize(iters, stop, P ? *P : d, tr);
  return 0;
}
int oracle_download_state(void* h, double* kf_pose, double* kf_vel, double* pt_xyz) {
  Oracle* o = ORA(h);
  for (int k = 0; k < o->n_kf; ++k) {
    if (kf_pose) {
      double* q = kf_pose + 7 * k;
      q[0] = o->kf[k].Twb.q.x; q[1] = o->kf[k].Twb.q.y; q[2] = o->kf[k].Twb.q.z; q[3] = o->kf[k].Twb.q.w;
      q[4] = o->kf[k].Twb.t[0]; q[5] = o->kf[k].Twb.t[1]; q[6] = o->kf[k].Twb.t[2];
    }
    if (kf_vel) for (int i = 0; i < 6; ++i) kf_vel[6 * k + i] = o->kf[k].vel[i];
  }
  if (pt_xyz) for (int i = 0; i < o->n_pt; ++i) for (int c = 0; c < 3; ++c) pt_xyz[3 * (size_t)i + c] = o->pt[i][c];
  return 0;
}
int oracle_edge_chi2(void* h, double* chi2) { Oracle* o = ORA(h); for (int64_t i = 0; i < o->n_obs; ++i) chi2[i] = o->obs_chi2(i); return 0; }
int oracle_edge_errors(void* h, double* err3) { std::memcpy(err3, ORA(h)->obs_err.data(), ORA(h)->obs_err.size() * 8); return 0; }
int oracle_active_robust_chi2(void* h, double* chi2) { *chi2 = ORA(h)->robust_chi2(); return 0; }
int oracle_outlier_flags(void* h, const gpba_thresholds* th, uint8_t* flags) { ORA(h)->outlier_flags(*th, flags); return 0; }
int oracle_set_levels(void* h, const uint8_t* level) {
  Oracle* o = ORA(h);
  o->structure_fresh = false;
  for (int64_t i = 0; i < o->n_obs; ++i) { if (level[i]) o->obs_flags[i] |= GPBA_OBS_LEVEL1; else o->obs_flags[i] &= ~GPBA_OBS_LEVEL1; }
  return 0;
}
int oracle_set_robust_kernel(void* h, int enabled) {
  Oracle* o = ORA(h);
  o->structure_fresh = false;
  for (int64_t i = 0; i < o->n_obs; ++i) { if (!enabled) o->obs_flags[i] |= GPBA_OBS_NO_KERNEL; else o->obs_flags[i] &= ~GPBA_OBS_NO_KERNEL; }
  return 0;
}
int oracle_compute_errors_inactive(void* h) {
  Oracle* o = ORA(h);
  for (int64_t i = 0; i < o->n_obs; ++i) if (o->obs_flags[i] & GPBA_OBS_LEVEL1) o->compute_obs_error(i);
  return 0;
}
int oracle_rejection_rounds(void* h, int n_rounds, int iters, const gpba_thresholds* th, const gpba_lm_params* P,
                            uint8_t* flags, gpba_lm_trace* traces) {
  gpba_lm_params d;
  oracle_default_lm_params(&d);
  ORA(h)->rejection_rounds(n_rounds, iters, *th, P ? *P : d, flags, traces);
  return 0;
}

// ---- math probes used by tests/test_oracle_*.py (pin the Lie / GP layer against the numpy mirror)
static SE3 se3_from7(const double* p) { SE3 T; T.q = {p[0], p[1], p[2], p[3]}; T.t[0] = p[4]; T.t[1] = p[5]; T.t[2] = p[6]; return T; }
static void se3_to7(const SE3& T, double* p) { p[0] = T.q.x; p[1] = T.q.y; p[2] = T.q.z; p[3] = T.q.w; p[4] = T.t[0]; p[5] = T.t[1]; p[6] = T.t[2]; }
// keyframe states / extrinsics of the last EVALUATED state (what gpba_download_evaluated_state hands back on the device)
int oracle_download_evaluated_state(void* h, double* kf_pose, double* kf_vel, double* cam_Tbc) {
  Oracle* o = ORA(h);
  if (o->eval_kf.empty()) { o->eval_kf = o->kf; o->eval_Tbc = o->Tbc; }
  for (int k = 0; k < o->n_kf; ++k) {
    se3_to7(o->eval_kf[k].Twb, kf_pose + 7 * k);
    for (int i = 0; i < 6; ++i) kf_vel[6 * k + i] = o->eval_kf[k].vel[i];
  }
  if (cam_Tbc) for (int c = 0; c < o->n_cam; ++c) se3_to7(o->eval_Tbc[c], cam_Tbc + 7 * c);
  return 0;
}
static V6 v6(const double* p) { V6 v; for (int i = 0; i < 6; ++i) v[i] = p[i]; return v; }
// Optimizer::PoseGPOptimizationFromeLastFrame for every frame of the batch (oracle/pose_only.h)
int oracle_pose_optimize(const gpba_pose_batch* B, double* cur_pose_out, double* cur_vel_out, double* prev_pose_out, double* prev_vel_out,
                         uint8_t* outlier_out, int32_t* n_inliers_out, gpba_lm_trace* traces) {
  for (int f = 0; f < B->n_frames; ++f) {
    ora::PoseOnlyFrame F(B, f);
    const int inl = F.run(traces ? traces + (size_t)f * GPBA_POSE_ROUNDS : nullptr, outlier_out);
    if (n_inliers_out) n_inliers_out[f] = inl;
    if (cur_pose_out) se3_to7(F.s2.Twb, cur_pose_out + 7 * f);
    if (cur_vel_out) std::memcpy(cur_vel_out + 6 * f, F.s2.vel.a, 48);
    if (prev_pose_out) se3_to7(F.s1.Twb, prev_pose_out + 7 * f);
    if (prev_vel_out) std::memcpy(prev_vel_out + 6 * f, F.s1.vel.a, 48);
  }
  return 0;
}
// Test hooks of the pose-only restatement: the linear system of frame f at the initial estimate, and the robust chi2 after
// the tangent step `delta` (n = 12 or 24) -- tests/test_pose_only.py checks b against central differences of chi2.
int oracle_pose_system(const gpba_pose_batch* B, int f, double* H, double* b, double* chi2) {
  ora::PoseOnlyFrame F(B, f);
  const double c = F.compute_errors();
  F.build_system();
  if (chi2) *chi2 = c;
  if (H) std::memcpy(H, F.H.data(), sizeof(double) * F.H.size());
  if (b) std::memcpy(b, F.b.data(), sizeof(double) * F.b.size());
  return F.n;
}
double oracle_pose_chi2_at(const gpba_pose_batch* B, int f, const double* delta) {
  ora::PoseOnlyFrame F(B, f);
  if (!F.fix1) { ora::PoseOnlyFrame::oplus(F.s1, delta); ora::PoseOnlyFrame::oplus(F.s2, delta + 12); }
  else ora::PoseOnlyFrame::oplus(F.s2, delta);
  return F.compute_errors();
}
// Tracking::MCRansac's hypotheses, one Optimizer::OptimizeVel each (oracle/vel_ransac.h)
int oracle_vel_ransac(const gpba_vel_batch* B, double* vel_out, int32_t* inliers_out, uint8_t* mask_out, int32_t* best_out, gpba_lm_trace* traces) {
  int best = -1, best_inl = 0;
  for (int h = 0; h < B->n_hyp; ++h) {
    ora::VelHypothesis H(B);
    const int inl = H.run(B->samples + (size_t)h * B->set_size, vel_out ? vel_out + 6 * h : nullptr,
                          mask_out ? mask_out + (size_t)h * B->n_match : nullptr, traces ? traces + h : nullptr);
    if (inliers_out) inliers_out[h] = inl;
    if (inl > best_inl) { best_inl = inl; best = h; }
  }
  if (best_out) *best_out = best;
  return 0;
}
void oracle_se3_exp(const double* xi, double* out7) { se3_to7(se3_exp(v6(xi)), out7); }
void oracle_se3_log(const double* T7, double* xi) { V6 v = se3_log(se3_from7(T7)); std::memcpy(xi, v.a, 48); }
void oracle_se3_mul(const double* a, const double* b, double* out7) { se3_to7(se3_mul(se3_from7(a), se3_from7(b)), out7); }
void oracle_se3_inv(const double* a, double* out7) { se3_to7(se3_inv(se3_from7(a)), out7); }
void oracle_se3_adj(const double* a, double* out36) { M6 A = se3_Adj(se3_from7(a)); std::memcpy(out36, A.a, 288); }
void oracle_se3_matrix(const double* a, double* out12) {
  SE3 T = se3_from7(a);
  M3 R = se3_R(T);
  for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) out12[r * 4 + c] = R(r, c); out12[r * 4 + 3] = T.t[r]; }
}
void oracle_se3_act(const double* a, const double* p, double* out3) { V3 v; v[0] = p[0]; v[1] = p[1]; v[2] = p[2]; V3 r = se3_act(se3_from7(a), v); std::memcpy(out3, r.a, 24); }
void oracle_jac_pose3(const double* xi, int which, double* out36) {  // 0 Jl, 1 Jr, 2 Jl^-1, 3 Jr^-1, 4 se3Adj
  M6 J;
  switch (which) {
    case 0: J = LeftJacobianPose3(v6(xi)); break;
    case 1: J = RightJacobianPose3(v6(xi)); break;
    case 2: J = LeftJacobianPose3Inv(v6(xi)); break;
    case 3: J = RightJacobianPose3Inv(v6(xi)); break;
    default: J = se3Adj(v6(xi)); break;
  }
  std::memcpy(out36, J.a, 288);
}
void oracle_query_pose(const double* qc, const double* T1, const double* T2, const double* v1, const double* v2, double t1,
                       double t2, double t, double* out7, double* At1, double* Pt1) {
  GaussianProcess G;
  G.set_diag(qc);
  M6x12 A, P;
  SE3 T = G.QueryPose(se3_from7(T1), se3_from7(T2), v6(v1), v6(v2), t1, t2, t, &A, &P);
  se3_to7(T, out7);
  if (At1) std::memcpy(At1, A.a, 72 * 8);
  if (Pt1) std::memcpy(Pt1, P.a, 72 * 8);
}
// One reprojection edge: error (dim) and Jacobians [J_kf1 (dim x12) | J_kf2 (dim x12) | J_pt (dim x3)], dim = ur>=0 ? 3 : 2
void oracle_edge_eval(const double* qc, int gp, const double* T1, const double* v1, double t1, const double* T2,
                      const double* v2, double t2, double t, const double* Tbc7, const double* intr, double bf,
                      const double* Xw, const double* obs3, double* err, double* J1, double* J2, double* Jp) {
  GaussianProcess G;
  G.set_diag(qc);
  KfState f1, f2;
  if (gp) { f1.Twb = se3_from7(T1); f1.vel = v6(v1); f1.time = t1; }
  f2.Twb = se3_from7(T2); f2.vel = v6(v2); f2.time = t2;
  Pinhole cam = {intr[0], intr[1], intr[2], intr[3]};
  V3 X; X[0] = Xw[0]; X[1] = Xw[1]; X[2] = Xw[2];
  int dim = obs3[2] >= 0 ? 3 : 2;
  reproj_error(G, gp != 0, dim, gp ? &f1 : nullptr, f2, t, se3_from7(Tbc7), cam, bf, X, obs3, err);
  if (J2) {
    double j1[36] = {0}, j2[36] = {0}, jp[9] = {0};
    reproj_jacobian(G, gp != 0, dim, gp ? &f1 : nullptr, f2, t, se3_from7(Tbc7), cam, bf, X, j1, j2, jp);
    if (J1) std::memcpy(J1, j1, dim * 12 * 8);
    std::memcpy(J2, j2, dim * 12 * 8);
    if (Jp) std::memcpy(Jp, jp, dim * 3 * 8);
  }
}
// the fourth Jacobian block of EdgeMonoGPExtrinsic (wrt VertexExtrinsic), dim x 6
void oracle_edge_jext(const double* qc, int gp, const double* T1, const double* v1, double t1, const double* T2,
                      const double* v2, double t2, double t, const double* Tbc7, const double* intr, double bf,
                      const double* Xw, const double* obs3, double* Jext) {
  GaussianProcess G;
  G.set_diag(qc);
  KfState f1, f2;
  if (gp) { f1.Twb = se3_from7(T1); f1.vel = v6(v1); f1.time = t1; }
  f2.Twb = se3_from7(T2); f2.vel = v6(v2); f2.time = t2;
  Pinhole cam = {intr[0], intr[1], intr[2], intr[3]};
  V3 X; X[0] = Xw[0]; X[1] = Xw[1]; X[2] = Xw[2];
  int dim = obs3[2] >= 0 ? 3 : 2;
  double j1[36] = {0}, j2[36] = {0}, jp[9] = {0}, je[18] = {0};
  reproj_jacobian(G, gp != 0, dim, gp ? &f1 : nullptr, f2, t, se3_from7(Tbc7), cam, bf, X, j1, j2, jp, je);
  std::memcpy(Jext, je, dim * 6 * 8);
}
// EdgeExtrinsicPrior: error (3) and Jacobian wrt the rotation part of the tangent (3 x 3)
void oracle_ext_prior_eval(const double* q_ini, const double* Tbc7, double* err3, double* J9) {
  const Quat qi = quat_inv(Quat{q_ini[0], q_ini[1], q_ini[2], q_ini[3]});
  const V3 e = ext_prior_error(qi, se3_from7(Tbc7));
  const M3 J = inverse3_cofactor(RightJacobianSO3_orb(e));
  for (int i = 0; i < 3; ++i) { err3[i] = e[i]; for (int k = 0; k < 3; ++k) J9[3 * i + k] = J(i, k); }
}
void oracle_prior_eval(const double* T1, const double* v1, double t1, const double* T2, const double* v2, double t2,
                       double* err12, double* Ji144, double* Jj144) {
  KfState f1, f2;
  f1.Twb = se3_from7(T1); f1.vel = v6(v1); f1.time = t1;
  f2.Twb = se3_from7(T2); f2.vel = v6(v2); f2.time = t2;
  prior_error(f1, f2, err12);
  if (Ji144) { M12 Ji, Jj; prior_jacobian(f1, f2, &Ji, &Jj); std::memcpy(Ji144, Ji.a, 144 * 8); std::memcpy(Jj144, Jj.a, 144 * 8); }
}
// EdgeVelReproj on one match (vel_ransac.h): error (2) and Jacobian (2 x 6) wrt the body velocity
void oracle_vel_edge_eval(const double* Tlast7, const double* Tbc7, const double* intr, double dt, const double* vel,
                          const double* Xw, const double* obs2, double* err2, double* J12) {
  gpba_vel_batch B;
  std::memset(&B, 0, sizeof(B));
  const int32_t cam0 = 0;
  const double w = 1.0;
  B.n_cam = 1; B.cam_intr = intr; B.cam_Tbc = Tbc7; B.cam_dt = &dt;
  std::memcpy(B.last_pose, Tlast7, 56);
  B.n_match = 1; B.obs_u = obs2; B.obs_v = obs2 + 1; B.obs_inv_sigma2 = &w; B.obs_xw = Xw; B.obs_cam = &cam0;
  B.huber_delta = 1.0;
  VelHypothesis H(&B);
  H.error(0, v6(vel), err2);
  double J[2][6];
  H.jacobian(0, v6(vel), J);
  std::memcpy(J12, J, sizeof(J));
}
// PoseVelocity::Update as the pose-only path applies it (pose_only.h), same arithmetic as Oracle::oplus
void oracle_posevel_update(const double* T7, const double* v, const double* upd12, double* T7_out, double* v_out) {
  KfState s;
  s.Twb = se3_from7(T7); s.vel = v6(v); s.time = 0;
  PoseOnlyFrame::oplus(s, upd12);
  se3_to7(s.Twb, T7_out);
  std::memcpy(v_out, s.vel.a, 48);
}
// RightJacobianSO3 as ORB-SLAM3 writes it (G2oTypes.cc:573-590), used by EdgeExtrinsicPrior
void oracle_right_jacobian_so3_orb(const double* w3, double* out9) {
  V3 w; w[0] = w3[0]; w[1] = w3[1]; w[2] = w3[2];
  const M3 J = RightJacobianSO3_orb(w);
  for (int i = 0; i < 3; ++i) for (int k = 0; k < 3; ++k) out9[3 * i + k] = J(i, k);
}
void oracle_huber(double delta, double e, double* rho3) { Huber h; h.setDelta(delta); h.robustify(e, rho3); }
int oracle_ldlt_dense(int n, const double* A, const double* b, double* x) {
  std::vector<double> M(A, A + (size_t)n * n);
  return ldlt_dense(n, M, b, x) ? 1 : 0;
}

}  // extern "C"
