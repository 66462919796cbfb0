// gpba_device.cuh -- device-side views shared by the kernels and the host driver of libgpba.
#pragma once
#include <stdint.h>
#include "gpba_math.cuh"

namespace gpba {

// Per-camera constants derived from MultiKeyFrame::mTbc / VertexExtrinsic and the Pinhole intrinsics.
struct CamConst {
  double fx, fy, cx, cy;
  double Rcb[9], tcb[3];  // Tcb = Tbc^-1
  double Rbc[9], tbc[3];
  double qbc[4];          // Tbc quaternion xyzw (for se3 products in K0)
  double AdjTbc[36];      // Adj(T_bc) = [[R_bc, t_bc^ R_bc], [0, R_bc]]: J_ext = J1 Adj(T_bc) (src/G2oTypes.cc:311-313)
};

// Record table (output of K0), one row per (KF_prev, KF_cur, cam, t) record:
//   [0..8]  R_cw   [9..11] t_cw      so that X_c = R_cw X_w + t_cw with T_cw = (T_wb(t) T_bc)^-1
//   [12..227] M (6 x 36, row-major): [M_T1 | M_V1 | M_T2 | M_V2 | A_c 0]  (SURVEY Appendix A.3); [0 | 0 | I | 0 | 0 0] for a
//             synchronous record.  The three 12-column slices belong to the record's three pose-like vertices: previous
//             keyframe, current keyframe, and (GP records of a camera whose VertexExtrinsic is free) the extrinsic, whose
//             12-slot holds the 6-dim tangent in front of six padding dimensions; A_c = Adj(T_bc).
#define GPBA_REC_MS 36
#define GPBA_REC_STRIDE (12 + 6 * GPBA_REC_MS)
#define GPBA_REC_M 12
#define GPBA_REC_LITE_STRIDE 12

#define GPBA_NO_SLOT 0xFFFFu
#define GPBA_TILE_OBS 1024     // observations per tile (a tile ends where a landmark ends)
#define GPBA_WIN_ROWS 224      // record rows a tile can stage: 224 x 96 B = 21 KB (a 1024-observation tile of C4 touches ~126)

struct DevView {
  // ---- static problem
  int n_cam, n_kf, n_pt, n_rec, n_prior, n_velp;
  const CamConst* cam;
  const double* kf_time;
  const int* kf_h;        // hessian index of each KF or -1
  const int* rec_kf1; const int* rec_kf2; const int* rec_cam; const double* rec_t;
  const int* prior_kf1; const int* prior_kf2; const int* velp_kf;
  double qc_inv[6];
  double bf;
  double hub_mono_delta, hub_mono_dsqr;     // delta <= 0: no kernel
  double hub_stereo_delta, hub_stereo_dsqr;
  double hub_prior_delta, hub_prior_dsqr;
  // ---- active observations, sorted by (landmark order, insertion order)
  int64_t n_aobs;
  const double* o_u; const double* o_v; const double* o_ur; const double* o_w;
  const int* o_rec; const int* o_lm;          // o_lm: sorted landmark index
  const uint8_t* o_flags;
  const int64_t* o_orig;                      // original observation index
  // ---- landmarks (sorted order)
  int n_lm;
  const int* lm_pt;                           // sorted landmark -> point index
  const int64_t* lm_obs_begin;                // [n_lm+1] into sorted obs
  // ---- landmark-aligned observation tiles of the streaming kernels (K1 / K2a): tile t covers the sorted landmarks
  //      [tile_lm[t], tile_lm[t+1]) and stages the record rows [tile_rlo[t], tile_rlo[t] + tile_rcnt[t]) in shared memory
  int n_tiles;
  const int* tile_lm; const int* tile_rlo; const int* tile_rcnt;
  // ---- record-major permutation (K2b)
  const int64_t* rperm;                       // [n_aobs] sorted-obs indices grouped by record
  int n_rseg;
  const int* rseg_rec; const int64_t* rseg_begin;  // [n_rseg], [n_rseg+1]
  // ---- Hessian storage
  int n_pose;                                 // free keyframes + free extrinsics (the extrinsics follow the keyframes)
  int n_pose_kf;                              // free keyframes
  const int* ext_h;                           // [n_cam] hessian index of the camera's extrinsic or -1 (fixed / absent)
  const unsigned char* ext_prior_on;          // [n_cam] EdgeExtrinsicPrior active
  const double* ext_prior_qinv;               // [n_cam][4] R_ini^-1 (xyzw)
  const double* ext_prior_info;               // [n_cam][9]
  int n_hpp, n_hs;
  const int* rec_hpp11; const int* rec_hpp12; const int* rec_hpp22;  // Hpp block index per record (-1 if n/a); hpp12 < 0 => none, bit30 set => transposed
  const int* rec_hpp13; const int* rec_hpp23; const int* rec_hpp33;  // blocks with the record's extrinsic (never transposed: it follows the keyframes)
  const int* prior_hpp11; const int* prior_hpp12; const int* prior_hpp22;
  const int* pose_hpp_diag;                   // [n_pose]
  const int* hs_from_hpp;                     // [n_hs] index into Hpp or -1
  const int* hs_diag_pose;                    // [n_hs] pose index if diagonal block else -1
};

}  // namespace gpba
