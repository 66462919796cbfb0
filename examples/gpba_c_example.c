/* examples/gpba_c_example.c -- the drop-in boundary used from plain C: a two-keyframe, one-camera-pair toy problem goes
 * through gpba_create / gpba_build_structure / gpba_optimize / gpba_download_state, and a one-frame batch through
 * gpba_pose_optimize.  Build:  gcc -Iinclude examples/gpba_c_example.c -Lamc-slam_b200 -lgpba -Wl,-rpath,amc-slam_b200 -lm
 * Exit code 0: optimised on the GPU; 3: no CUDA device (the library has no CPU fallback and says so); 1: anything else. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "gpba.h"

#define N_KF 3
#define N_PT 40
#define N_CAM 2 /* one asynchronous camera + the reference camera */

static double frand(unsigned* s) { *s = *s * 1664525u + 1013904223u; return (double)(*s >> 8) / 16777216.0; }

int main(void) {
  /* cameras: identical intrinsics, the async camera 10 cm to the left of the body frame */
  double cam_intr[N_CAM * 4] = {500, 500, 480, 300, 500, 500, 480, 300};
  double cam_Tbc[N_CAM * 7] = {0, 0, 0, 1, 0.1, 0, 0, 0, 0, 0, 1, 0, 0, 0};
  /* keyframes move 0.4 m per 0.1 s along x; identity rotation; the first one is fixed */
  double kf_pose[N_KF * 7], kf_vel[N_KF * 6], kf_time[N_KF];
  uint8_t kf_fixed[N_KF] = {1, 0, 0};
  unsigned seed = 7;
  for (int k = 0; k < N_KF; ++k) {
    double* p = kf_pose + 7 * k;
    p[0] = p[1] = p[2] = 0; p[3] = 1; p[4] = 0.4 * k + (k ? 0.02 * (frand(&seed) - 0.5) : 0); p[5] = p[6] = 0;
    double* v = kf_vel + 6 * k;
    v[0] = 4; v[1] = v[2] = v[3] = v[4] = v[5] = 0;
    kf_time[k] = 0.1 * k;
  }
  /* records: for keyframes 1 and 2 one GP record (async camera, captured between the keyframes) and one synchronous */
  int32_t rec_kf1[4] = {0, -1, 1, -1}, rec_kf2[4] = {1, 1, 2, 2}, rec_cam[4] = {0, 1, 0, 1};
  double rec_t[4] = {0.05, 0.1, 0.15, 0.2};
  /* points in front of the rig, every point seen by all four records: exact projections + a little noise */
  double pt_xyz[N_PT * 3], obs_u[N_PT * 4], obs_v[N_PT * 4], obs_w[N_PT * 4];
  int32_t obs_rec[N_PT * 4], obs_pt[N_PT * 4];
  int64_t n_obs = 0;
  for (int j = 0; j < N_PT; ++j) {
    double X = 2.0 * (frand(&seed) - 0.5) * 4, Y = 2.0 * (frand(&seed) - 0.5) * 2, Z = 6 + 10 * frand(&seed);
    pt_xyz[3 * j] = X; pt_xyz[3 * j + 1] = Y; pt_xyz[3 * j + 2] = Z;
    for (int r = 0; r < 4; ++r) {
      const double tx = 0.4 * (rec_t[r] / 0.1) + (rec_cam[r] == 0 ? 0.1 : 0.0); /* true camera centre (pure translation) */
      obs_u[n_obs] = 500 * (X - tx) / Z + 480 + 0.3 * (frand(&seed) - 0.5);
      obs_v[n_obs] = 500 * Y / Z + 300 + 0.3 * (frand(&seed) - 0.5);
      obs_w[n_obs] = 1.0; obs_rec[n_obs] = r; obs_pt[n_obs] = j;
      ++n_obs;
    }
    pt_xyz[3 * j] += 0.05 * (frand(&seed) - 0.5); /* perturbed initial estimate */
  }
  int32_t prior_kf1[2] = {0, 1}, prior_kf2[2] = {1, 2}, velp_kf[2] = {1, 2};

  gpba_problem P;
  memset(&P, 0, sizeof(P));
  P.n_cam = N_CAM; P.cam_intr = cam_intr; P.cam_Tbc = cam_Tbc; P.bf = 0;
  P.n_kf = N_KF; P.kf_pose = kf_pose; P.kf_vel = kf_vel; P.kf_time = kf_time; P.kf_fixed = kf_fixed;
  P.n_pt = N_PT; P.pt_xyz = pt_xyz;
  P.n_rec = 4; P.rec_kf1 = rec_kf1; P.rec_kf2 = rec_kf2; P.rec_cam = rec_cam; P.rec_t = rec_t;
  P.n_obs = n_obs; P.obs_u = obs_u; P.obs_v = obs_v; P.obs_ur = NULL; P.obs_inv_sigma2 = obs_w; P.obs_rec = obs_rec; P.obs_pt = obs_pt;
  P.n_prior = 2; P.prior_kf1 = prior_kf1; P.prior_kf2 = prior_kf2; P.n_velp = 2; P.velp_kf = velp_kf;
  const double qc[6] = {0.02, 0.02, 0.02, 0.002, 0.002, 0.002};
  memcpy(P.qc, qc, sizeof(qc));
  P.huber_mono = (double)(float)sqrt(5.991); P.huber_stereo = (double)(float)sqrt(7.815); P.huber_prior = 0;
  P.lambda_init = 1.0; P.linear_solver = GPBA_SOLVER_DENSE_CHOL;

  gpba_handle* h = NULL;
  int rc = gpba_create(&P, -1, &h);
  if (rc == GPBA_ERR_NO_DEVICE) { printf("no CUDA device: %s\n", gpba_last_error()); return 3; }
  if (rc != GPBA_OK) { printf("gpba_create failed (%d): %s\n", rc, gpba_last_error()); return 1; }
  gpba_structure_info info;
  gpba_lm_trace tr;
  double chi0 = 0;
  if (gpba_build_structure(h, &info) != GPBA_OK || gpba_compute_errors(h, &chi0) != GPBA_OK ||
      gpba_optimize(h, 10, NULL, NULL, &tr) != GPBA_OK) { printf("optimize failed: %s\n", gpba_last_error()); gpba_destroy(h); return 1; }
  double kp[N_KF * 7], kv[N_KF * 6], px[N_PT * 3];
  gpba_download_state(h, kp, kv, px);
  gpba_destroy(h);
  printf("free keyframes %d, Hschur blocks %lld; chi2 %.3f -> %.3f in %d LM iterations; keyframe 2 at x = %.4f m\n",
         (int)info.n_free_kf, (long long)info.n_hschur, chi0, tr.chi2_after[tr.n_iters - 1], (int)tr.n_iters, kp[7 * 2 + 4]);
  return (tr.n_iters >= 1 && tr.chi2_after[tr.n_iters - 1] < chi0) ? 0 : 1;
}
