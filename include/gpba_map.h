/*
 * gpba_map.h -- persistent SoA mirror of the AMC-SLAM map for the GP-BA path.
 *
 * The reference rebuilds a g2o graph with one `new` per vertex and edge on every BA call by chasing
 * MultiKeyFrame / MapPoint pointers: Optimizer::LocalGPBA (src/Optimizer.cc:713-1211 window selection and graph
 * construction, :1263-1430 outlier erasure and write-back) and Optimizer::BundleAdjustment (:61-315, :324-366).
 * Once the solve runs on the GPU that O(N_obs) walk dominates the call.  This mirror is updated by the map's own
 * mutation points (one call per Map::AddKeyFrame, MapPoint::AddObservation, EraseObservation, SetBadFlag, SetPose,
 * SetWorldPos ...) and turns "run LocalGPBA for keyframe K" into: select the window (same rules as :718-834), emit
 * the flattened `gpba_problem` arrays of include/gpba.h in one pass over contiguous per-point observation lists,
 * hand them to gpba_create / gpba_optimize, and apply the result (poses, points, erased observations) back.
 *
 * Host-side only (no device work); one map is used by one thread at a time -- the caller holds the map mutex the
 * reference takes (Map::mMutexMapUpdate, Optimizer.cc:1351).  Ids are the reference's mnId values.
 * Deviations from the reference, all in iteration order only: MapPoint::GetObservations() is a std::map keyed by
 * keyframe POINTER (src/MapPoint.cc), so the reference visits a point's keyframes in allocation-address order; the
 * mirror visits them in ascending keyframe id.  This changes the insertion order of edges (summation order) and which
 * keyframes fill the 50 fixed slots when more than 50 qualify, nothing else.  GPObs observations (MapPoint::
 * GetGPObservations) are never created by the reference's tracking code and are not mirrored.
 */
#ifndef GPBA_MAP_H
#define GPBA_MAP_H
#include "gpba.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gpba_map gpba_map;
typedef struct gpba_window gpba_window;

typedef struct gpba_map_config {
  int32_t n_cam;            /* MultiKeyFrame::nCamera: async cameras + the reference camera (last)   */
  const double* cam_intr;   /* [n_cam][4] fx fy cx cy                                                */
  const double* cam_Tbc;    /* [n_cam][7] MultiKeyFrame::mTbc                                        */
  double bf;
  double qc[6];             /* GaussianProcess::mQc diagonal                                         */
} gpba_map_config;

/* All functions return GPBA_OK or GPBA_ERR_INVALID; the message of the last failure on this thread: */
const char* gpba_map_last_error(void);

int gpba_map_create(const gpba_map_config* cfg, gpba_map** out);
void gpba_map_destroy(gpba_map* m);

/* ---- mutation hooks ------------------------------------------------------------------------------------------- */
/* Map::AddKeyFrame (MultiKeyFrame ctor: mnId, mPrevKF, mTimeStamp, mvTimeStamps; src/KeyFrame.cc:132-139).
 * prev_id < 0: no previous keyframe.  Links prev->mNextKF as the reference does. */
int gpba_map_add_keyframe(gpba_map* m, int64_t id, int64_t prev_id, const double pose_Twb[7], const double vel[6],
                          double time, const double* cam_time /* [n_cam] */);
/* MultiKeyFrame::SetPose / SetVelocity (vel may be NULL: unchanged) */
int gpba_map_set_keyframe_state(gpba_map* m, int64_t id, const double pose_Twb[7], const double vel[6]);
/* KeyFrameCulling + MultiKeyFrame::SetBadFlag: when the keyframe has both neighbours they are linked to each other
 * (src/LocalMapping.cc:873-876), its observations are erased from the points (SetBadFlag -> EraseObservation). */
int gpba_map_set_keyframe_bad(gpba_map* m, int64_t id);
/* Map::AddMapPoint / MapPoint::SetWorldPos / MapPoint::SetBadFlag */
int gpba_map_add_point(gpba_map* m, int64_t id, const double xyz[3]);
int gpba_map_set_point(gpba_map* m, int64_t id, const double xyz[3]);
int gpba_map_set_point_bad(gpba_map* m, int64_t id);
/* MapPoint::AddObservation(pKF, idx, cam) + MultiKeyFrame::AddMapPoint: keypoint (u, v[, ur]) of camera `cam` of keyframe
 * `kf` observes point `pt`; inv_sigma2 = mvInvLevelSigma2[octave] / uncertainty2; close = mvTrackDepth[cam] < 10 m.
 * A second observation of the same (kf, cam, pt) replaces the first (the reference's vector<int> slot is overwritten). */
int gpba_map_add_observation(gpba_map* m, int64_t kf, int32_t cam, int64_t pt, double u, double v, double ur,
                             double inv_sigma2, int32_t close_flag);
/* n calls of gpba_map_add_observation in array order (initial fill of the mirror from an existing map, Atlas load).
 * ur / close_flag may be NULL (monocular / not close).  Stops at the first invalid entry and reports its index. */
int gpba_map_add_observations(gpba_map* m, int64_t n, const int64_t* kf, const int32_t* cam, const int64_t* pt,
                              const double* u, const double* v, const double* ur, const double* inv_sigma2,
                              const uint8_t* close_flag, int64_t* n_done);
/* MapPoint::EraseObservation(pKF, cam) + MultiKeyFrame::EraseMapPointMatch */
int gpba_map_erase_observation(gpba_map* m, int64_t kf, int32_t cam, int64_t pt);
/* MultiKeyFrame::UpdateConnections (src/KeyFrame.cc:455-550) at the places the reference calls it (LocalMapping.cc:261,694,
 * LoopClosing.cc:840,952,995, Tracking.cc:1613-1614): counts, for every other keyframe, the map points it shares with
 * `kf` (one count per keypoint of `kf`, so a point matched in two cameras counts twice, as in the reference); keyframes
 * with >= 15 shared points (or the best one if none reaches 15) become the ordered covisible list of `kf` and get the
 * symmetric AddConnection + UpdateBestCovisibles (:250-287) -- including the reference's quirk that a list rebuilt by
 * UpdateBestCovisibles holds ALL counted keyframes, not only those above the threshold.  gpba_map_set_keyframe_bad erases
 * the connections like SetBadFlag does (:663-666, EraseConnection :763-777).  Ties are ordered by keyframe id where the
 * reference orders them by pointer. */
int gpba_map_update_connections(gpba_map* m, int64_t kf);
/* MultiKeyFrame::GetVectorCovisibleKeyFrames: ids in decreasing weight; *n_out = list length (may exceed capacity). */
int gpba_map_covisibles(const gpba_map* m, int64_t kf, int64_t* ids, int32_t* weights, int32_t capacity, int32_t* n_out);
/* counts: [0] keyframes (not bad) [1] points (not bad) [2] observations */
int gpba_map_stats(const gpba_map* m, int64_t out[3]);

/* ---- flattening ----------------------------------------------------------------------------------------------- */
/* Optimizer::LocalGPBA(pKF, ..., bLarge, ...) window selection and graph construction (src/Optimizer.cc:718-1211).
 * covisible = pKF->GetVectorCovisibleKeyFrames() ids in that order; n_covisible < 0: use the mirror's own covisibility
 * list of kf_id (gpba_map_update_connections).  At most one of them joins the window (maxCovKF = 0 at :784). */
int gpba_map_local_window(gpba_map* m, int64_t kf_id, int32_t large, const int64_t* covisible, int32_t n_covisible,
                          gpba_window** out);
/* Optimizer::BundleAdjustment(vpKFs = all keyframes, vpMP = all points, ...) graph construction (:85-315);
 * init_kf_id = pMap->GetInitKFid() (the fixed vertex). */
int gpba_map_global_window(gpba_map* m, int64_t init_kf_id, gpba_window** out);
void gpba_window_destroy(gpba_window* w);

/* The flattened problem (arrays owned by the window; valid until gpba_window_destroy). */
const gpba_problem* gpba_window_problem(const gpba_window* w);
/* optimize() iteration count of the first stage: 10 (opt_it1, :1221) for local windows, caller's choice for global. */
int32_t gpba_window_iterations(const gpba_window* w);
/* Index maps from problem indices back to the map.  kf_role: 0 temporal window (vpOptimizableKFs), 1 covisible
 * (lpOptVisKFs), 2 fixed (lFixedKeyFrames).  Any pointer may be NULL. */
int gpba_window_ids(const gpba_window* w, int64_t* kf_id /* [n_kf] */, int32_t* kf_role /* [n_kf] */,
                    int64_t* pt_id /* [n_pt] */, int64_t* obs_kf /* [n_obs] */, int32_t* obs_cam /* [n_obs] */,
                    int64_t* obs_pt /* [n_obs] */);
/* cam_obs[c] of :1010,1128: GP edges per async camera (the extrinsic self-calibration gate, :1225-1240) */
int gpba_window_cam_obs(const gpba_window* w, int32_t* cam_obs /* [n_cam] */);

/* ---- write-back ----------------------------------------------------------------------------------------------- */
/* LocalGPBA tail (:1349-1430): if (2 err < err_end or NaN) and !large, nothing is applied and *applied = 0 ("FAIL
 * LOCAL-GP BA").  Otherwise the flagged observations (gpba_outlier_flags) are erased from the mirror, the free keyframes
 * get the optimised pose (float-rounded like SetPose(...cast<float>()); velocities only if kf_vel != NULL -- the
 * reference writes them in BundleAdjustment (:336-339) but not in LocalGPBA) and the points their position.
 * n_erased / erased_obs (indices into the window's observation arrays, capacity n_obs) tell the caller which
 * MapPoint::EraseObservation / EraseMapPointMatch calls to replay on its own objects.  flags may be NULL. */
int gpba_window_apply(gpba_map* m, const gpba_window* w, const double* kf_pose, const double* kf_vel,
                      const double* pt_xyz, const uint8_t* flags, float err, float err_end, int32_t* applied,
                      int64_t* n_erased, int64_t* erased_obs);

/* LocalGPBA's extrinsic write-back (:1419-1428): for every camera with cam_obs[c] >= min_obs (extrin_thresh = 50) the
 * mirror's Tbc becomes the calibrated estimate (gpba_get_extrinsics), float-rounded like v->estimate().cast<float>();
 * later windows are flattened with it.  n_updated (optional): how many cameras were written. */
int gpba_map_apply_extrinsics(gpba_map* m, const gpba_window* w, const double* cam_Tbc /* [n_cam][7] */, int32_t min_obs,
                              int32_t* n_updated);
/* MultiKeyFrame::mTbc as the mirror holds it: [n_cam][7] qx qy qz qw tx ty tz */
int gpba_map_extrinsics(const gpba_map* m, double* cam_Tbc);

#ifdef __cplusplus
}
#endif
#endif
