// adapter/tracking_gpba.h -- reference-side bindings of the two frame-rate entry points of libgpba.so
// (include/gpba.h: gpba_pose_optimize, gpba_vel_ransac).
//
// Header-only C++11, to be compiled INSIDE the AMC-SLAM tree like adapter/g2o_gpba_solver.h (it needs the reference's
// MultiFrame / MapPoint / Sophus headers, which do not exist in this repository's container: SURVEY.md 0.6).  Everything
// below reads exactly the fields the reference reads while it builds its g2o graphs:
//
//   gpba::PoseGPOptimizationFromeLastFrame(pFrame, fix)   body of Optimizer::PoseGPOptimizationFromeLastFrame
//                                                         (src/Optimizer.cc:369-686): same arguments, same return value
//                                                         (nInitialCorrespondences - nBad), same side effects
//                                                         (pFrame->mvbOutlier, SetPose, SetVelocity).
//   gpba::OptimizeVelBatch(...)                           the loop of Tracking::MCRansac (src/Tracking.cc:1939-2002) over
//                                                         Optimizer::OptimizeVel (src/Optimizer.cc:2364-2447): the caller
//                                                         keeps its std::mt19937 and passes the sample sets.
//
// Call sites: src/Tracking.cc:1863, 1912, 2036 (pose-only) and :2029 (MCRansac).
#pragma once

#include <unordered_set>
#include <vector>

#include "Frame.h"          // MultiFrame: N, nCamera, mvpMapPoints, mvKeysUn, mmpKeyToCam, mvTimeStamps, mvuRight, mmpGlobalToLocalID,
                            //             mvInvLevelSigma2, mvpCamera, mvbOutlier, mpPrevFrame, mTimeStamp, mbf, mpGP, static mTbc
#include "MapPoint.h"       // GetWorldPos, mvTrackDepth, mGlobalMutex
#include "gpba.h"

namespace gpba {

inline void put_se3(const Sophus::SE3d& T, double* out) {   // qx qy qz qw tx ty tz
  const Eigen::Quaterniond q = T.unit_quaternion();
  out[0] = q.x(); out[1] = q.y(); out[2] = q.z(); out[3] = q.w();
  out[4] = T.translation()[0]; out[5] = T.translation()[1]; out[6] = T.translation()[2];
}
inline Sophus::SE3d get_se3(const double* p) {
  return Sophus::SE3d(Eigen::Quaterniond(p[3], p[0], p[1], p[2]), Eigen::Vector3d(p[4], p[5], p[6]));
}

// Camera block shared by both entry points: intrinsics (GeometricCamera::getParameter, CameraModels/GeometricCamera.h:81),
// MultiFrame::mTbc (include/Frame.h:220), mbf, diag(mQc).
struct RigArrays {
  std::vector<double> intr, Tbc;
  void fill(ORB_SLAM3::MultiFrame* pF) {
    const int n = pF->nCamera;
    intr.resize(4 * n); Tbc.resize(7 * n);
    for (int c = 0; c < n; ++c) {
      for (int k = 0; k < 4; ++k) intr[4 * c + k] = pF->mvpCamera[c]->getParameter(k);
      put_se3(ORB_SLAM3::MultiFrame::mTbc[c].cast<double>(), &Tbc[7 * c]);
    }
  }
};

// ---------------------------------------------------------------------------------------------------------------------
inline int PoseGPOptimizationFromeLastFrame(ORB_SLAM3::MultiFrame* pFrame, bool fix) {
  using namespace ORB_SLAM3;
  MultiFrame* pPrev = pFrame->mpPrevFrame;
  RigArrays rig;
  rig.fill(pFrame);
  const int nCam = pFrame->nCamera;
  std::vector<double> u, v, ur, w, xw;
  std::vector<int32_t> cam;
  std::vector<uint8_t> flags, outlier;
  std::vector<int> index;   // match -> keypoint index i of the frame
  {
    std::unique_lock<std::mutex> lock(MapPoint::mGlobalMutex);          // Optimizer.cc:417
    for (int i = 0; i < pFrame->N; ++i) {                               // :419-519
      MapPoint* pMP = pFrame->mvpMapPoints[i];
      if (!pMP) continue;
      const cv::KeyPoint& kp = pFrame->mvKeysUn[i];
      const int c = pFrame->mmpKeyToCam[i];
      const Eigen::Vector2d obs(kp.pt.x, kp.pt.y);
      const float unc2 = pFrame->mvpCamera[c]->uncertainty2(obs);
      const float invSigma2 = pFrame->mvInvLevelSigma2[kp.octave] / unc2;
      const Eigen::Vector3d X = pMP->GetWorldPos().cast<double>();
      u.push_back(kp.pt.x); v.push_back(kp.pt.y);
      ur.push_back(c == nCam - 1 ? pFrame->mvuRight[pFrame->mmpGlobalToLocalID[i]] : -1.f);
      w.push_back(invSigma2);
      xw.push_back(X[0]); xw.push_back(X[1]); xw.push_back(X[2]);
      cam.push_back(c);
      flags.push_back((pMP->mvTrackDepth[c] < 10.f ? GPBA_OBS_CLOSE : 0) | (pFrame->mvbOutlier[i] ? GPBA_OBS_LEVEL1 : 0));   // :590, :447-450
      index.push_back(i);
    }
  }
  const int64_t obs_begin[2] = {0, (int64_t)u.size()};
  double prev_pose[7], cur_pose[7], prev_vel[6], cur_vel[6];
  put_se3(pPrev->GetPoseW().cast<double>(), prev_pose);                 // PoseVelocity(MultiFrame*), G2oTypes.cc:33-39
  put_se3(pFrame->GetPoseW().cast<double>(), cur_pose);
  const Eigen::Matrix<double, 6, 1> v1 = pPrev->GetVelocity().cast<double>(), v2 = pFrame->GetVelocity().cast<double>();
  for (int k = 0; k < 6; ++k) { prev_vel[k] = v1[k]; cur_vel[k] = v2[k]; }
  const double prev_time = pPrev->mTimeStamp, cur_time = pFrame->mTimeStamp;
  std::vector<double> cam_time(pFrame->mvTimeStamps.begin(), pFrame->mvTimeStamps.end());
  const uint8_t prev_fixed = fix ? 1 : 0;

  gpba_pose_batch B;
  std::memset(&B, 0, sizeof(B));
  B.n_cam = nCam; B.cam_intr = rig.intr.data(); B.cam_Tbc = rig.Tbc.data(); B.bf = pFrame->mbf;
  for (int k = 0; k < 6; ++k) B.qc[k] = pFrame->mpGP->mQc(k, k);
  B.n_frames = 1;
  B.prev_pose = prev_pose; B.prev_vel = prev_vel; B.prev_time = &prev_time; B.prev_fixed = &prev_fixed;
  B.cur_pose = cur_pose; B.cur_vel = cur_vel; B.cur_time = &cur_time; B.cam_time = cam_time.data();
  B.obs_begin = obs_begin; B.obs_u = u.data(); B.obs_v = v.data(); B.obs_ur = ur.data(); B.obs_inv_sigma2 = w.data();
  B.obs_xw = xw.data(); B.obs_cam = cam.data(); B.obs_flags = flags.data();
  B.huber_mono = (float)sqrt(5.991); B.huber_stereo = (float)sqrt(7.815);   // :414-415

  outlier.resize(u.size());
  double pose[7], vel[6];
  int32_t inliers = 0;
  if (gpba_pose_optimize(&B, -1, pose, vel, nullptr, nullptr, outlier.data(), &inliers, nullptr) != GPBA_OK) {
    std::cerr << "gpba_pose_optimize: " << gpba_last_error() << std::endl;
    return 0;
  }
  for (size_t k = 0; k < index.size(); ++k) pFrame->mvbOutlier[index[k]] = outlier[k] != 0;   // :596-664
  pFrame->SetPose(get_se3(pose).inverse().cast<float>());                                     // :678
  Eigen::Matrix<double, 6, 1> vout;
  for (int k = 0; k < 6; ++k) vout[k] = vel[k];
  pFrame->SetVelocity(vout.cast<float>());                                                    // :679
  return inliers;                                                                             // nInitialCorrespondences - nBad
}

// ---------------------------------------------------------------------------------------------------------------------
// All hypotheses of one MCRansac call.  samples[h] = the `min_set` indices (into `indices`, i.e. positions in
// vMatchedFeatures) drawn for hypothesis h.  On return vel[h], inliers[h] and, for the winning hypothesis, vbBestInliers
// (Tracking.cc:1967-1979).  Returns the index of the winning hypothesis or -1.
inline int OptimizeVelBatch(ORB_SLAM3::MultiFrame* pF1, ORB_SLAM3::MultiFrame* pF2, const std::vector<int>& indices,
                            const std::vector<std::vector<int> >& samples, std::vector<Eigen::Matrix<double, 6, 1> >& vel,
                            std::vector<int>& inliers, std::vector<bool>& vbBestInliers, double threshold = 2.0) {
  using namespace ORB_SLAM3;
  RigArrays rig;
  rig.fill(pF1);
  const int nCam = pF1->nCamera, nm = (int)indices.size(), nh = (int)samples.size();
  if (nh == 0) return -1;
  const int set_size = (int)samples[0].size();
  std::vector<double> u(nm), v(nm), w(nm), xw(3 * (size_t)nm), cam_dt(nCam);
  std::vector<int32_t> cam(nm), flat((size_t)nh * set_size);
  {
    std::unique_lock<std::mutex> lock(MapPoint::mGlobalMutex);                                // Optimizer.cc:2388
    for (int k = 0; k < nm; ++k) {                                                            // :2390-2420
      const int index = indices[k];
      const cv::KeyPoint& kp = pF1->mvKeysUn[index];
      const Eigen::Vector3d X = pF1->mvpMapPoints[index]->GetWorldPos().cast<double>();
      u[k] = kp.pt.x; v[k] = kp.pt.y; w[k] = pF1->mvInvLevelSigma2[kp.octave];
      xw[3 * k] = X[0]; xw[3 * k + 1] = X[1]; xw[3 * k + 2] = X[2];
      cam[k] = pF1->mmpKeyToCam[index];
    }
  }
  for (int c = 0; c < nCam; ++c) cam_dt[c] = pF1->mvTimeStamps[c] - pF2->mTimeStamp;          // :2399
  for (int h = 0; h < nh; ++h)
    for (int s = 0; s < set_size; ++s) flat[(size_t)h * set_size + s] = samples[h][s];

  gpba_vel_batch B;
  std::memset(&B, 0, sizeof(B));
  B.n_cam = nCam; B.cam_intr = rig.intr.data(); B.cam_Tbc = rig.Tbc.data(); B.cam_dt = cam_dt.data();
  put_se3(pF2->GetPoseW().cast<double>(), B.last_pose);                                       // :2398
  const Eigen::Matrix<double, 6, 1> v0 = pF1->GetVelocity().cast<double>();                    // :2381
  for (int k = 0; k < 6; ++k) B.vel_init[k] = v0[k];
  B.n_match = nm; B.obs_u = u.data(); B.obs_v = v.data(); B.obs_inv_sigma2 = w.data(); B.obs_xw = xw.data(); B.obs_cam = cam.data();
  B.n_hyp = nh; B.set_size = set_size; B.samples = flat.data();
  B.huber_delta = 5.991; B.threshold = threshold; B.iterations = 40;                          // :2410, :2423

  std::vector<double> vout(6 * (size_t)nh);
  std::vector<int32_t> inl(nh);
  std::vector<uint8_t> mask((size_t)nh * nm);
  int32_t best = -1;
  if (gpba_vel_ransac(&B, -1, vout.data(), inl.data(), mask.data(), &best, nullptr) != GPBA_OK) {
    std::cerr << "gpba_vel_ransac: " << gpba_last_error() << std::endl;
    return -1;
  }
  vel.resize(nh); inliers.assign(inl.begin(), inl.end());
  for (int h = 0; h < nh; ++h)
    for (int k = 0; k < 6; ++k) vel[h][k] = vout[6 * (size_t)h + k];
  vbBestInliers.assign(nm, false);
  if (best >= 0)
    for (int k = 0; k < nm; ++k) vbBestInliers[k] = mask[(size_t)best * nm + k] != 0;
  return best;
}

}  // namespace gpba
