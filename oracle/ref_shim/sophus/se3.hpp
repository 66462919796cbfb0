// -*- C++ -*-
// TEST INFRASTRUCTURE (oracle/ref_shim): stand-in for Sophus::SE3, see so3.hpp beside this file.  Follows the documented
// algorithms of Thirdparty/Sophus/sophus/se3.hpp in own code: Adj (se3.hpp:103-111), log with V^-1 (se3.hpp:223-256),
// product (se3.hpp:304-309), exp with V (se3.hpp:761-782); tangent order [upsilon (translation), omega (rotation)].
#pragma once
#include "so3.hpp"

namespace Sophus {

template <class S> class SE3 {
 public:
  typedef Eigen::Matrix<S, 6, 1> Tangent;
  typedef Eigen::Matrix<S, 3, 1> Point;
  typedef Eigen::Matrix<S, 4, 4> Transformation;
  typedef Eigen::Matrix<S, 6, 6> Adjoint;

  SE3() {}
  SE3(const SO3<S>& so3, const Point& t) : so3_(so3), t_(t) {}
  SE3(const Eigen::Matrix<S, 3, 3>& R, const Point& t) : so3_(R), t_(t) {}
  SE3(const Eigen::Quaternion<S>& q, const Point& t) : so3_(q), t_(t) {}
  Eigen::Quaternion<S> unit_quaternion() const { return so3_.unit_quaternion(); }

  const SO3<S>& so3() const { return so3_; }
  SO3<S>& so3() { return so3_; }
  const Point& translation() const { return t_; }
  Point& translation() { return t_; }
  Eigen::Matrix<S, 3, 3> rotationMatrix() const { return so3_.matrix(); }
  template <class U> SE3<U> cast() const { return SE3<U>(so3_.template cast<U>(), t_.template cast<U>()); }

  SE3 inverse() const {
    const SO3<S> inv = so3_.inverse();
    return SE3(inv, inv * (t_ * S(-1)));
  }
  SE3 operator*(const SE3& o) const { return SE3(so3_ * o.so3_, t_ + so3_ * o.t_); }
  Point operator*(const Point& p) const { return so3_ * p + t_; }
  Adjoint Adj() const {
    const Eigen::Matrix<S, 3, 3> R = so3_.matrix();
    Adjoint res;
    res.template block<3, 3>(0, 0) = R;
    res.template block<3, 3>(3, 3) = R;
    res.template block<3, 3>(0, 3) = SO3<S>::hat(t_) * R;
    return res;
  }
  Transformation matrix() const {
    Transformation T;
    T.template block<3, 3>(0, 0) = so3_.matrix();
    T.template block<3, 1>(0, 3) = t_;
    T(3, 3) = S(1);
    return T;
  }
  Eigen::Matrix<S, 3, 4> matrix3x4() const {
    Eigen::Matrix<S, 3, 4> T;
    T.template block<3, 3>(0, 0) = so3_.matrix();
    T.template block<3, 1>(0, 3) = t_;
    return T;
  }
  Tangent log() const {
    Tangent u;
    const auto ot = so3_.logAndTheta();
    const S theta = ot.theta;
    u.template tail<3>() = ot.tangent;
    const Eigen::Matrix<S, 3, 3> Omega = SO3<S>::hat(ot.tangent), I = Eigen::Matrix<S, 3, 3>::Identity();
    if (std::fabs(theta) < Constants<S>::epsilon()) {
      const Eigen::Matrix<S, 3, 3> V_inv = I - S(0.5) * Omega + S(1. / 12.) * (Omega * Omega);
      u.template head<3>() = V_inv * t_;
    } else {
      const S half = S(0.5) * theta;
      const Eigen::Matrix<S, 3, 3> V_inv =
          I - S(0.5) * Omega + (S(1) - theta * std::cos(half) / (S(2) * std::sin(half))) / (theta * theta) * (Omega * Omega);
      u.template head<3>() = V_inv * t_;
    }
    return u;
  }
  static SE3 exp(const Tangent& a) {
    const Eigen::Matrix<S, 3, 1> omega = a.template tail<3>();
    S theta;
    const SO3<S> so3 = SO3<S>::expAndTheta(omega, &theta);
    const Eigen::Matrix<S, 3, 3> Omega = SO3<S>::hat(omega), Omega_sq = Omega * Omega;
    Eigen::Matrix<S, 3, 3> V;
    if (theta < Constants<S>::epsilon()) {
      V = so3.matrix();
    } else {
      const S theta_sq = theta * theta;
      V = Eigen::Matrix<S, 3, 3>::Identity() + (S(1) - std::cos(theta)) / theta_sq * Omega +
          (theta - std::sin(theta)) / (theta_sq * theta) * Omega_sq;
    }
    return SE3(so3, V * a.template head<3>());
  }
 private:
  SO3<S> so3_;
  Point t_;
};
typedef SE3<double> SE3d;
typedef SE3<float> SE3f;

}  // namespace Sophus
