// -*- C++ -*-
// TEST INFRASTRUCTURE (oracle/ref_shim_tracking): MultiFrame and MapPoint with the members adapter/tracking_gpba.h reads, declared
// with the TYPES the reference declares them with (include/Frame.h:89-96, 166, 218-393; include/MapPoint.h:135, 199, 243) --
// float poses and velocities (Sophus::SE3f, Eigen::VectorXf), unordered_map key tables, std::vector<bool> outlier flags -- so
// that compiling the adapter against them checks its conversions.  Used only by the compile check in oracle/Makefile
// (target adapter_tracking_check); the pinned libraries use the lighter stand-ins of oracle/ref_shim/KeyFrame.h.
#pragma once
#include <mutex>
#include <unordered_map>
#include <vector>
#include <opencv2/core/core.hpp>
#include <Eigen/Core>
#include "sophus/se3.hpp"
#include "GaussianProcess.h"

namespace ORB_SLAM3 {

class GeometricCamera {
 public:
  virtual ~GeometricCamera() {}
  virtual Eigen::Vector2d project(const Eigen::Vector3d& v3D) = 0;
  virtual Eigen::Matrix<double, 2, 3> projectJac(const Eigen::Vector3d& v3D) = 0;
  virtual float getParameter(const int i) = 0;
  virtual float uncertainty2(const Eigen::Matrix<double, 2, 1>& p2D) = 0;
};

class MapPoint {
 public:
  Eigen::Vector3f GetWorldPos() { return mWorldPos; }
  Eigen::Vector3f mWorldPos;
  std::vector<float> mvTrackDepth;
  static std::mutex mGlobalMutex;
};

class MultiFrame {
 public:
  void SetPose(const Sophus::SE3<float>& Tbw) { mTbw = Tbw; }
  void SetVelocity(Eigen::VectorXf Twist) { mVel = Twist; }
  Eigen::VectorXf GetVelocity() const { return mVel; }
  inline Sophus::SE3f GetPoseW() const { return mTbw.inverse(); }
  Sophus::SE3f mTbw;
  Eigen::VectorXf mVel;
  GaussianProcess* mpGP = nullptr;
  double mTimeStamp = 0;
  std::vector<double> mvTimeStamps;
  static int nCamera;
  float mbf = 0;
  int N = 0;
  std::vector<cv::KeyPoint> mvKeysUn;
  std::unordered_map<size_t, int> mmpKeyToCam;
  std::unordered_map<size_t, int> mmpGlobalToLocalID;
  std::vector<MapPoint*> mvpMapPoints;
  std::vector<float> mvuRight;
  std::vector<bool> mvbOutlier;
  MultiFrame* mpPrevFrame = nullptr;
  std::vector<float> mvInvLevelSigma2;
  std::vector<GeometricCamera*> mvpCamera;
  static std::vector<Sophus::SE3<float>> mTbc;
};

}  // namespace ORB_SLAM3
