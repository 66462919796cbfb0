// -*- C++ -*-
// TEST INFRASTRUCTURE (oracle/ref_shim): stand-in for Sophus::SO3 on top of the Eigen stand-in, so that the reference's
// sources compile here.  The vendored Sophus (Thirdparty/Sophus/sophus/so3.hpp) needs the real Eigen (Quaternion, Map,
// MatrixBase traits) and cannot be used; this file follows its documented algorithms -- unit quaternion storage,
// normalisation in every quaternion constructor (so3.hpp:481-488, 297-303), atan-based log (so3.hpp:247-291), exp with the
// small-angle series below 1e-10 (so3.hpp:583-620), rotation of a point by two cross products (so3.hpp:358-369) -- in own
// code.  It is therefore NOT a pin of the Lie-group layer; that one is pinned against scipy expm/logm (tests/test_oracle_math.py).
#pragma once
#include "../Eigen/Core"

namespace Sophus {

template <class S> struct Constants {
  static S epsilon() { return S(1e-10); }
  static S pi() { return S(3.141592653589793238462643383279502884); }
};
template <> struct Constants<float> {
  static float epsilon() { return 1e-5f; }
  static float pi() { return 3.141592653589793238462643383279502884f; }
};

template <class S> class SO3 {
 public:
  typedef Eigen::Matrix<S, 3, 1> Tangent;
  typedef Eigen::Matrix<S, 3, 1> Point;
  typedef Eigen::Matrix<S, 3, 3> Transformation;

  SO3() : x_(0), y_(0), z_(0), w_(1) {}
  // stand-in only: quaternion coefficients, normalised as every Sophus quaternion constructor does
  static SO3 fromQuaternion(S x, S y, S z, S w) {
    SO3 q;
    const S len = std::sqrt(x * x + y * y + z * z + w * w);
    q.x_ = x / len; q.y_ = y / len; q.z_ = z / len; q.w_ = w / len;
    return q;
  }
  // from a rotation matrix: Eigen's quaternion-from-matrix rule (trace branch, else the largest diagonal entry)
  SO3(const Transformation& R) {
    S t = R(0, 0) + R(1, 1) + R(2, 2);
    S q[4];
    if (t > S(0)) {
      t = std::sqrt(t + S(1));
      q[3] = S(0.5) * t;
      t = S(0.5) / t;
      q[0] = (R(2, 1) - R(1, 2)) * t; q[1] = (R(0, 2) - R(2, 0)) * t; q[2] = (R(1, 0) - R(0, 1)) * t;
    } else {
      int i = 0;
      if (R(1, 1) > R(0, 0)) i = 1;
      if (R(2, 2) > R(i, i)) i = 2;
      const int j = (i + 1) % 3, k = (j + 1) % 3;
      t = std::sqrt(R(i, i) - R(j, j) - R(k, k) + S(1));
      q[i] = S(0.5) * t;
      t = S(0.5) / t;
      q[3] = (R(k, j) - R(j, k)) * t; q[j] = (R(j, i) + R(i, j)) * t; q[k] = (R(k, i) + R(i, k)) * t;
    }
    *this = fromQuaternion(q[0], q[1], q[2], q[3]);
  }
  Eigen::Quaternion<S> unit_quaternion() const { return Eigen::Quaternion<S>(w_, x_, y_, z_); }
  explicit SO3(const Eigen::Quaternion<S>& q) { *this = fromQuaternion(q.x(), q.y(), q.z(), q.w()); }
  S qx() const { return x_; }
  S qy() const { return y_; }
  S qz() const { return z_; }
  S qw() const { return w_; }

  template <class U> SO3<U> cast() const { return SO3<U>::fromQuaternion((U)x_, (U)y_, (U)z_, (U)w_); }
  SO3 inverse() const { return fromQuaternion(-x_, -y_, -z_, w_); }
  Transformation matrix() const {
    const S tx = 2 * x_, ty = 2 * y_, tz = 2 * z_;
    const S twx = tx * w_, twy = ty * w_, twz = tz * w_, txx = tx * x_, txy = ty * x_, txz = tz * x_, tyy = ty * y_,
            tyz = tz * y_, tzz = tz * z_;
    Transformation R;
    R(0, 0) = 1 - (tyy + tzz); R(0, 1) = txy - twz; R(0, 2) = txz + twy;
    R(1, 0) = txy + twz; R(1, 1) = 1 - (txx + tzz); R(1, 2) = tyz - twx;
    R(2, 0) = txz - twy; R(2, 1) = tyz + twx; R(2, 2) = 1 - (txx + tyy);
    return R;
  }
  Transformation Adj() const { return matrix(); }
  SO3 operator*(const SO3& b) const {
    const SO3& a = *this;
    return fromQuaternion(a.w_ * b.x_ + a.x_ * b.w_ + a.y_ * b.z_ - a.z_ * b.y_,
                          a.w_ * b.y_ + a.y_ * b.w_ + a.z_ * b.x_ - a.x_ * b.z_,
                          a.w_ * b.z_ + a.z_ * b.w_ + a.x_ * b.y_ - a.y_ * b.x_,
                          a.w_ * b.w_ - a.x_ * b.x_ - a.y_ * b.y_ - a.z_ * b.z_);
  }
  Point operator*(const Point& p) const {
    S ux = y_ * p(2) - z_ * p(1), uy = z_ * p(0) - x_ * p(2), uz = x_ * p(1) - y_ * p(0);
    ux += ux; uy += uy; uz += uz;
    Point r;
    r(0) = p(0) + w_ * ux + (y_ * uz - z_ * uy);
    r(1) = p(1) + w_ * uy + (z_ * ux - x_ * uz);
    r(2) = p(2) + w_ * uz + (x_ * uy - y_ * ux);
    return r;
  }
  struct TangentAndTheta { Tangent tangent; S theta; };
  TangentAndTheta logAndTheta() const {
    TangentAndTheta J;
    const S squared_n = x_ * x_ + y_ * y_ + z_ * z_;
    S two_atan_nbyw_by_n;
    if (squared_n < Constants<S>::epsilon() * Constants<S>::epsilon()) {
      const S squared_w = w_ * w_;
      two_atan_nbyw_by_n = S(2) / w_ - S(2.0 / 3.0) * squared_n / (w_ * squared_w);
      J.theta = S(2) * squared_n / w_;
    } else {
      const S n = std::sqrt(squared_n);
      if (std::fabs(w_) < Constants<S>::epsilon())
        two_atan_nbyw_by_n = (w_ > S(0) ? Constants<S>::pi() : -Constants<S>::pi()) / n;
      else
        two_atan_nbyw_by_n = S(2) * std::atan(n / w_) / n;
      J.theta = two_atan_nbyw_by_n * n;
    }
    J.tangent = Tangent(two_atan_nbyw_by_n * x_, two_atan_nbyw_by_n * y_, two_atan_nbyw_by_n * z_);
    return J;
  }
  Tangent log() const { return logAndTheta().tangent; }
  static SO3 expAndTheta(const Tangent& omega, S* theta) {
    const S theta_sq = omega.squaredNorm();
    S imag, real;
    if (theta_sq < Constants<S>::epsilon() * Constants<S>::epsilon()) {
      *theta = S(0);
      const S theta_po4 = theta_sq * theta_sq;
      imag = S(0.5) - S(1.0 / 48.0) * theta_sq + S(1.0 / 3840.0) * theta_po4;
      real = S(1) - S(1.0 / 8.0) * theta_sq + S(1.0 / 384.0) * theta_po4;
    } else {
      *theta = std::sqrt(theta_sq);
      const S half = S(0.5) * (*theta);
      imag = std::sin(half) / (*theta);
      real = std::cos(half);
    }
    SO3 q;   // Sophus stores these coefficients without re-normalising
    q.x_ = imag * omega(0); q.y_ = imag * omega(1); q.z_ = imag * omega(2); q.w_ = real;
    return q;
  }
  static SO3 exp(const Tangent& omega) { S th; return expAndTheta(omega, &th); }
  static Transformation hat(const Tangent& o) {
    Transformation O;
    O(0, 1) = -o(2); O(0, 2) = o(1);
    O(1, 0) = o(2); O(1, 2) = -o(0);
    O(2, 0) = -o(1); O(2, 1) = o(0);
    return O;
  }
 private:
  S x_, y_, z_, w_;
};
typedef SO3<double> SO3d;
typedef SO3<float> SO3f;

}  // namespace Sophus
