// TEST DOUBLE -- NOT A PRODUCT PATH AND NOT A FALLBACK.  The entry points of include/gpba.h that the reference-side binding
// (adapter/g2o_gpba_solver.h) calls, implemented on the CPU oracle, so that the binding's HOST logic -- flattening, the
// extrinsic hand-over, write-back of estimates, the stale-error hand-back, the level-1 solver seam -- can be executed inside
// the reference's real g2o::SparseOptimizer in this container, which has no GPU.  It is built into
// oracle/_ref/libgpba_abi_double.so, linked only into oracle/_ref/libadapter_on_double.so, loaded only by
// tests/test_whole_path_reference.py.  The product library amc-slam_b200/libgpba.so is unaffected: it has no CPU path and
// returns GPBA_ERR_NO_DEVICE without a device (tests/test_abi.py).  What runs on the device is tested by the GPU suite.
#include <cmath>
#include <cstring>
#include <limits>
#include <vector>
#include "../include/gpba.h"

extern "C" {   // oracle/gpba_oracle.cc
void* oracle_create(const gpba_problem* p);
void oracle_destroy(void* h);
void oracle_reset_state(void* h, const double* kf_pose, const double* kf_vel, const double* pt_xyz);
int oracle_set_extrinsics(void* h, const uint8_t* free_, const double* prior_q, const double* prior_info);
int oracle_get_extrinsics(void* h, double* Tbc7);
int oracle_build_structure(void* h, gpba_structure_info* info);
int oracle_compute_errors(void* h, double* chi2);
int oracle_build_system(void* h);
int oracle_set_lambda(void* h, double l, int backup);
int oracle_restore_diagonal(void* h);
int oracle_solve(void* h, int* ok);
int oracle_vector_size(void* h, int64_t* n);
int oracle_get_x(void* h, double* x);
int oracle_get_b(void* h, double* b);
int oracle_optimize(void* h, int iters, const volatile unsigned char* stop, const gpba_lm_params* P, gpba_lm_trace* tr);
int oracle_download_state(void* h, double* kf_pose, double* kf_vel, double* pt_xyz);
int oracle_download_evaluated_state(void* h, double* kf_pose, double* kf_vel, double* cam_Tbc);
int oracle_edge_errors(void* h, double* err3);
}

struct gpba_handle {
  void* o;
  int64_t n_obs;
  std::vector<uint8_t> level1;   // edges inactive at creation (the adapter does not change levels through the C ABI)
};

extern "C" {

int gpba_create(const gpba_problem* prob, int /*device*/, gpba_handle** out) {
  gpba_handle* h = new gpba_handle();
  h->o = oracle_create(prob);
  h->n_obs = prob->n_obs;
  h->level1.assign((size_t)prob->n_obs, 0);
  if (prob->obs_flags) for (int64_t i = 0; i < prob->n_obs; ++i) h->level1[i] = (prob->obs_flags[i] & GPBA_OBS_LEVEL1) ? 1 : 0;
  *out = h;
  return GPBA_OK;
}
int gpba_destroy(gpba_handle* h) { if (h) { oracle_destroy(h->o); delete h; } return GPBA_OK; }
const char* gpba_last_error(void) { return "test double on the CPU oracle"; }
int gpba_reset_state(gpba_handle* h, const double* kf_pose, const double* kf_vel, const double* pt_xyz) { oracle_reset_state(h->o, kf_pose, kf_vel, pt_xyz); return GPBA_OK; }
int gpba_set_extrinsics(gpba_handle* h, const gpba_extrinsics* e) { return oracle_set_extrinsics(h->o, e->free_mask, e->prior_R, e->prior_info); }
int gpba_get_extrinsics(gpba_handle* h, double* cam_Tbc) { return oracle_get_extrinsics(h->o, cam_Tbc); }
int gpba_build_structure(gpba_handle* h, gpba_structure_info* info) { return oracle_build_structure(h->o, info); }
int gpba_compute_errors(gpba_handle* h, double* robust_chi2) { return oracle_compute_errors(h->o, robust_chi2); }
int gpba_build_system(gpba_handle* h) { return oracle_build_system(h->o); }
int gpba_set_lambda(gpba_handle* h, double lambda, int backup) { return oracle_set_lambda(h->o, lambda, backup); }
int gpba_restore_diagonal(gpba_handle* h) { return oracle_restore_diagonal(h->o); }
int gpba_solve(gpba_handle* h, int* ok) { return oracle_solve(h->o, ok); }
int gpba_vector_size(gpba_handle* h, int64_t* n) { return oracle_vector_size(h->o, n); }
int gpba_get_x(gpba_handle* h, double* x) { return oracle_get_x(h->o, x); }
int gpba_get_b(gpba_handle* h, double* b) { return oracle_get_b(h->o, b); }
int gpba_optimize(gpba_handle* h, int iters, const volatile unsigned char* stop_flag, const gpba_lm_params* params, gpba_lm_trace* trace) {
  oracle_optimize(h->o, iters, stop_flag, params, trace);
  return GPBA_OK;
}
int gpba_download_state(gpba_handle* h, double* kf_pose, double* kf_vel, double* pt_xyz) { return oracle_download_state(h->o, kf_pose, kf_vel, pt_xyz); }
int gpba_download_evaluated_state(gpba_handle* h, double* kf_pose, double* kf_vel, double* cam_Tbc) {
  return oracle_download_evaluated_state(h->o, kf_pose, kf_vel, cam_Tbc);
}
int gpba_edge_errors(gpba_handle* h, double* err3) {   // NaN for inactive edges, as include/gpba.h:231-235 specifies
  oracle_edge_errors(h->o, err3);
  for (int64_t i = 0; i < h->n_obs; ++i)
    if (h->level1[i]) for (int d = 0; d < 3; ++d) err3[3 * i + d] = std::numeric_limits<double>::quiet_NaN();
  return GPBA_OK;
}

}  // extern "C"
