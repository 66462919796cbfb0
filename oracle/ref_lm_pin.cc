// TEST INFRASTRUCTURE, not product code.  The reference's own Levenberg-Marquardt controller
// (Thirdparty/g2o/g2o/core/optimization_algorithm_levenberg.cpp, compiled unmodified; see ref_shim_lm/lm_standins.h) run on
// top of the oracle's level-1 steps: a g2o::Solver whose buildStructure / buildSystem / setLambda / solve / restoreDiagonal
// are oracle_build_structure / oracle_build_system / ... and a SparseOptimizer whose computeActiveErrors / activeRobustChi2 /
// push / pop / discardTop / update are the oracle's.  The loop around solve() restates SparseOptimizer::optimize
// (sparse_optimizer.cpp:354-419, no batch statistics, not verbose).  Built by `make -C oracle _ref` into
// oracle/_ref/libg2o_ref_lm.so; tests/test_ref_pin.py requires that the trace equals oracle_optimize's on the same problem
// BIT FOR BIT: the linear algebra is the same code on both sides, so every difference is a difference in the controller
// (lambda schedule, gain ratio, accept / reject, stop criteria) -- row a22 of SURVEY 8 pinned against the reference's code.
#include <cmath>
#include <cstdint>
#include <limits>
#include <vector>
#include "../include/gpba.h"
#include "Thirdparty/g2o/g2o/core/optimization_algorithm_levenberg.h"
#include "Thirdparty/g2o/g2o/core/solver.h"

extern "C" {   // oracle/gpba_oracle.cc
int oracle_build_structure(void* h, gpba_structure_info* info);
int oracle_get_hpp_pattern(void* h, int32_t* rows, int32_t* cols);
int oracle_compute_errors(void* h, double* chi2);
int oracle_active_robust_chi2(void* h, double* chi2);
int oracle_build_system(void* h);
int oracle_set_lambda(void* h, double l, int backup);
int oracle_restore_diagonal(void* h);
int oracle_solve(void* h, int* ok);
int oracle_vector_size(void* h, int64_t* n);
int oracle_get_x(void* h, double* x);
int oracle_get_b(void* h, double* b);
int oracle_get_hpp(void* h, double* blocks);
int oracle_get_hll(void* h, double* blocks);
int oracle_oplus(void* h, const double* x);
int oracle_push(void* h);
int oracle_pop(void* h);
int oracle_discard_top(void* h);
}

namespace {

struct DiagVertex : g2o::OptimizableGraph::Vertex {   // one Hessian-indexed vertex: its diagonal block, refreshed per buildSystem
  int dim;
  bool marg;
  std::vector<double> H;   // dim x dim, row-major
  DiagVertex(int d, bool m) : dim(d), marg(m), H((size_t)d * d, 0.0) {}
  int dimension() const override { return dim; }
  const double& hessian(int i, int j) const override { return H[(size_t)i * dim + j]; }
  bool marginalized() const override { return marg; }
};

enum Event { EV_CHI2 = 0, EV_PUSH = 1, EV_POP = 2, EV_DISCARD = 3 };

struct OracleOptimizer : g2o::SparseOptimizer {
  void* h;
  const volatile unsigned char* stop = nullptr;
  g2o::OptimizableGraph::VertexContainer iv;
  std::vector<int> ev;          // what the controller did, in order
  std::vector<double> chi2s;    // every activeRobustChi2() it read
  explicit OracleOptimizer(void* h_) : h(h_) {}
  ~OracleOptimizer() override { for (auto* v : iv) delete v; }
  void computeActiveErrors() override { oracle_compute_errors(h, nullptr); }
  double activeRobustChi2() const override {
    double c; oracle_active_robust_chi2(h, &c);
    const_cast<OracleOptimizer*>(this)->ev.push_back(EV_CHI2);
    const_cast<OracleOptimizer*>(this)->chi2s.push_back(c);
    return c;
  }
  void push() override { ev.push_back(EV_PUSH); oracle_push(h); }
  void pop() override { ev.push_back(EV_POP); oracle_pop(h); }
  void discardTop() override { ev.push_back(EV_DISCARD); oracle_discard_top(h); }
  void update(const double* u) override { oracle_oplus(h, u); }
  bool terminate() override { return stop && *stop; }
  const g2o::OptimizableGraph::VertexContainer& indexMapping() const override { return iv; }
  const g2o::OptimizableGraph::VertexContainer& activeVertices() const override { return iv; }
};

struct OracleSolver : g2o::Solver {
  void* h;
  OracleOptimizer* opt = nullptr;
  gpba_structure_info info;
  int64_t n_pose = 0, n_lm = 0;
  std::vector<int32_t> hpp_r, hpp_c;
  std::vector<double> hpp, hll;
  bool dbg = false;
  explicit OracleSolver(void* h_) : h(h_) { std::memset(&info, 0, sizeof(info)); }
  bool init(g2o::SparseOptimizer* o, bool) override { _optimizer = o; return true; }
  bool buildStructure(bool) override {
    if (oracle_build_structure(h, &info) != 0) return false;
    int64_t n; oracle_vector_size(h, &n);
    resizeVector((size_t)n);
    n_lm = info.n_active_pt;
    n_pose = (n - 3 * n_lm) / 12;   // keyframes and free extrinsics (12-slots), gpba_oracle.cc build_structure
    hpp_r.resize(info.n_hpp); hpp_c.resize(info.n_hpp);
    oracle_get_hpp_pattern(h, hpp_r.data(), hpp_c.data());
    hpp.resize((size_t)info.n_hpp * 144); hll.resize((size_t)n_lm * 9);
    for (auto* v : opt->iv) delete v;
    opt->iv.clear();
    for (int64_t i = 0; i < n_pose; ++i) opt->iv.push_back(new DiagVertex(12, false));
    for (int64_t l = 0; l < n_lm; ++l) opt->iv.push_back(new DiagVertex(3, true));
    return true;
  }
  bool updateStructure(const std::vector<g2o::HyperGraph::Vertex*>&, const g2o::HyperGraph::EdgeSet&) override { return false; }
  bool buildSystem() override {
    oracle_build_system(h);
    oracle_get_b(h, _b);
    oracle_get_hpp(h, hpp.data()); oracle_get_hll(h, hll.data());
    for (int64_t k = 0; k < info.n_hpp; ++k)
      if (hpp_r[k] == hpp_c[k]) static_cast<DiagVertex*>(opt->iv[hpp_r[k]])->H.assign(&hpp[k * 144], &hpp[k * 144] + 144);
    for (int64_t l = 0; l < n_lm; ++l)
      static_cast<DiagVertex*>(opt->iv[n_pose + l])->H.assign(&hll[l * 9], &hll[l * 9] + 9);
    return true;
  }
  bool solve() override {
    int ok = 0;
    oracle_solve(h, &ok);
    oracle_get_x(h, _x);
    return ok != 0;
  }
  bool computeMarginals(g2o::SparseBlockMatrix<Eigen::MatrixXd>&, const std::vector<std::pair<int, int> >&) override { return false; }
  bool setLambda(double l, bool backup) override { oracle_set_lambda(h, l, backup ? 1 : 0); return true; }
  void restoreDiagonal() override { oracle_restore_diagonal(h); }
  bool supportsSchur() override { return true; }
  bool schur() override { return true; }
  void setSchur(bool) override {}
  void setWriteDebug(bool b) override { dbg = b; }
  bool writeDebug() const override { return dbg; }
  bool saveHessian(const std::string&) const override { return false; }
};

}  // namespace

extern "C" {

// SparseOptimizer::optimize(iters) with OptimizationAlgorithmLevenberg(solver), setUserLambdaInit(lambda_init) when
// lambda_init > 0 (Optimizer.cc:912, 1229: 1e-0 / 1e-5 ... as the caller sets it), on the oracle handle `h` (oracle_create).
// Fills the same trace oracle_optimize fills; chi2_log (may be NULL, capacity log_cap) receives every chi2 the controller read.
int ref_lm_optimize(void* h, int iters, double lambda_init, int max_trials, const volatile unsigned char* stop, gpba_lm_trace* tr,
                    double* chi2_log, int log_cap, int* log_n) {
  OracleSolver* solver = new OracleSolver(h);   // owned by the algorithm (optimization_algorithm_with_hessian.cpp:47-50)
  OracleOptimizer opt(h);
  opt.stop = stop;
  solver->opt = &opt;
  g2o::OptimizationAlgorithmLevenberg alg(solver);
  if (lambda_init > 0) alg.setUserLambdaInit(lambda_init);
  if (max_trials > 0) alg.setMaxTrialsAfterFailure(max_trials);
  alg.setOptimizer(&opt);
  std::memset(tr, 0, sizeof(*tr));
  tr->result = GPBA_RESULT_OK;
  // ---- sparse_optimizer.cpp:354-419
  int cjIterations = 0;
  bool ok = alg.init(false);
  if (!ok) return -1;
  g2o::OptimizationAlgorithm::SolverResult result = g2o::OptimizationAlgorithm::OK;
  for (int i = 0; i < iters && !opt.terminate() && ok; i++) {
    const size_t ev0 = opt.ev.size(), c0 = opt.chi2s.size();
    result = alg.solve(i, false);
    ok = (result == g2o::OptimizationAlgorithm::OK);
    ++cjIterations;
    if (i < GPBA_MAX_ITERS) {   // what the controller did, read off its calls
      const double iniChi = opt.chi2s[c0];
      bool accepted = false;
      for (size_t e = ev0; e < opt.ev.size(); ++e) if (opt.ev[e] == EV_DISCARD) accepted = true;
      tr->levenberg_iterations[i] = alg.levenbergIteration();
      tr->chi2_before[i] = iniChi;
      tr->chi2_after[i] = accepted ? opt.chi2s.back() : iniChi;
      tr->lambda[i] = alg.currentLambda();
      tr->total_trials += alg.levenbergIteration();
      tr->last_trial_chi2 = opt.chi2s.back();
    }
  }
  tr->n_iters = cjIterations;
  tr->result = (int32_t)result;
  if (log_n) *log_n = (int)opt.chi2s.size();
  if (chi2_log) for (int k = 0; k < (int)opt.chi2s.size() && k < log_cap; ++k) chi2_log[k] = opt.chi2s[k];
  return result == g2o::OptimizationAlgorithm::Fail ? 0 : cjIterations;
}

}  // extern "C"
