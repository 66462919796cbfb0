// -*- C++ -*-
// TEST INFRASTRUCTURE (oracle/ref_shim): what src/G2oTypes.cc reads from the map classes -- MultiKeyFrame
// (include/KeyFrame.h:212,222,325,455,538: GetPoseInverse, GetVelocity, mTimeStamp, mTbc, mvpCamera, mbf), MultiFrame
// (include/Frame.h:96,166,220,232,259,393) and the camera interface (include/CameraModels/GeometricCamera.h:62,71).  The
// real headers pull in OpenCV, DBoW2 and the whole map.  The reference stores the extrinsics mTbc as SE3f and casts them to
// double at every use; here they are held in double so that the pin is not limited by float rounding of the test inputs
// (`.cast<double>()` is then the identity).
#pragma once
#include <vector>
#include <Eigen/Core>
#include "sophus/se3.hpp"
#include "GaussianProcess.h"

namespace ORB_SLAM3 {

class GeometricCamera {
 public:
  virtual ~GeometricCamera() {}
  virtual Eigen::Vector2d project(const Eigen::Vector3d& v3D) = 0;
  virtual Eigen::Matrix<double, 2, 3> projectJac(const Eigen::Vector3d& v3D) = 0;
  virtual float getParameter(const int i) { (void)i; return 0.f; }   // GeometricCamera.h:101 (float intrinsics), read by adapter/
};

class MultiKeyFrame {
 public:
  Sophus::SE3d GetPoseInverse() { return Twb; }
  Eigen::Matrix<double, 6, 1> GetVelocity() { return Vel; }
  Sophus::SE3d Twb;
  Eigen::Matrix<double, 6, 1> Vel;
  double mTimeStamp = 0;
  double mbf = 0;
  std::vector<GeometricCamera*> mvpCamera;
  static std::vector<Sophus::SE3d> mTbc;
};

class MultiFrame {
 public:
  Sophus::SE3d GetPoseW() const { return Twb; }
  Eigen::Matrix<double, 6, 1> GetVelocity() const { return Vel; }
  Sophus::SE3d Twb;
  Eigen::Matrix<double, 6, 1> Vel;
  double mTimeStamp = 0;
  double mbf = 0;
  std::vector<GeometricCamera*> mvpCamera;
  static std::vector<Sophus::SE3d> mTbc;
};

}  // namespace ORB_SLAM3
