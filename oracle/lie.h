// oracle/lie.h -- TEST INFRASTRUCTURE ONLY (CPU restatement; never linked into libgpba.so).
//
// PARITY, per layer: the reference holds no golden vectors / known-answer tests for this path
//  and as a whole cannot be compiled here (no Eigen3/OpenCV/Boost, fact 0.6).
//  * The Pose3utils part of this file (Q, J_l/J_r and inverses, SO3 J_l/J_l^-1, se3Adj) IS pinned against the reference's
//    own src/Pose3utils.cc, compiled unmodified into oracle/_ref (stand-in Eigen/Sophus headers, oracle/ref_shim/):
//    tests/test_ref_pin.py, 1e-11.
//  * The Sophus part (SO3/SE3 exp, log, Adj, products) is NOT pinned by oracle/_ref -- the vendored Sophus needs the real
//    Eigen, so the reference side uses a stand-in there; it is pinned by (a) the Sophus property tests restated in
//    tests/test_oracle_math.py, (b) an independent scipy expm/logm + series mirror (oracle/numpy_mirror.py) and (c) g2o's
//    own central-difference Jacobian scheme.
//
// Fixed-size dense helpers + SO(3)/SE(3) restated from the vendored Sophus and from Pose3utils:
//   Thirdparty/Sophus/sophus/so3.hpp  expAndTheta :583-619, logAndTheta :247-291, product :324-338,
//                                     normalize :297-303, point action :355-366, inverse :229-231
//   Thirdparty/Sophus/sophus/se3.hpp  exp :761-783, log :223-255, Adj :103-111, inverse :208-211,
//                                     product :304-309, point action :321-325
//   Thirdparty/Sophus/sophus/common.hpp:94   Constants<double>::epsilon() = 1e-10
//   src/Pose3utils.cc                 Q :5-22, J_l/J_r and inverses :24-46, SO3 J_l/J_l^-1 :48-73, se3Adj :111-118
// Quaternion -> rotation matrix follows Eigen's QuaternionBase::toRotationMatrix (third party,
// un-vendored, version unpinned: find_package(Eigen3 3.1.0), CMakeLists.txt:48).
#pragma once
#include <cmath>
#include <cfloat>
#include <cstring>

namespace ora {

template <int R, int C>
struct Mat {
  double a[R * C];
  double& operator()(int r, int c) { return a[r * C + c]; }
  double operator()(int r, int c) const { return a[r * C + c]; }
  double& operator[](int i) { return a[i]; }
  double operator[](int i) const { return a[i]; }
  static Mat Zero() {
    Mat m;
    for (int i = 0; i < R * C; ++i) m.a[i] = 0.0;
    return m;
  }
  static Mat Identity() {
    Mat m = Zero();
    for (int i = 0; i < (R < C ? R : C); ++i) m(i, i) = 1.0;
    return m;
  }
  template <int R2, int C2>
  Mat<R2, C2> block(int r0, int c0) const {
    Mat<R2, C2> m;
    for (int r = 0; r < R2; ++r)
      for (int c = 0; c < C2; ++c) m(r, c) = (*this)(r0 + r, c0 + c);
    return m;
  }
  template <int R2, int C2>
  void set_block(int r0, int c0, const Mat<R2, C2>& b) {
    for (int r = 0; r < R2; ++r)
      for (int c = 0; c < C2; ++c) (*this)(r0 + r, c0 + c) = b(r, c);
  }
};

template <int R, int K, int C>
inline Mat<R, C> operator*(const Mat<R, K>& A, const Mat<K, C>& B) {
  Mat<R, C> m;
  for (int r = 0; r < R; ++r)
    for (int c = 0; c < C; ++c) {
      double s = 0.0;
      for (int k = 0; k < K; ++k) s += A(r, k) * B(k, c);
      m(r, c) = s;
    }
  return m;
}
template <int R, int C>
inline Mat<R, C> operator+(const Mat<R, C>& A, const Mat<R, C>& B) {
  Mat<R, C> m;
  for (int i = 0; i < R * C; ++i) m.a[i] = A.a[i] + B.a[i];
  return m;
}
template <int R, int C>
inline Mat<R, C> operator-(const Mat<R, C>& A, const Mat<R, C>& B) {
  Mat<R, C> m;
  for (int i = 0; i < R * C; ++i) m.a[i] = A.a[i] - B.a[i];
  return m;
}
template <int R, int C>
inline Mat<R, C> operator-(const Mat<R, C>& A) {
  Mat<R, C> m;
  for (int i = 0; i < R * C; ++i) m.a[i] = -A.a[i];
  return m;
}
template <int R, int C>
inline Mat<R, C> operator*(double s, const Mat<R, C>& A) {
  Mat<R, C> m;
  for (int i = 0; i < R * C; ++i) m.a[i] = s * A.a[i];
  return m;
}
template <int R, int C>
inline Mat<R, C> operator*(const Mat<R, C>& A, double s) {
  return s * A;
}
template <int R, int C>
inline Mat<C, R> transpose(const Mat<R, C>& A) {
  Mat<C, R> m;
  for (int r = 0; r < R; ++r)
    for (int c = 0; c < C; ++c) m(c, r) = A(r, c);
  return m;
}

typedef Mat<3, 1> V3;
typedef Mat<6, 1> V6;
typedef Mat<3, 3> M3;
typedef Mat<6, 6> M6;

inline double dot3(const V3& a, const V3& b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
inline V3 cross3(const V3& a, const V3& b) {
  V3 c;
  c[0] = a[1] * b[2] - a[2] * b[1];
  c[1] = a[2] * b[0] - a[0] * b[2];
  c[2] = a[0] * b[1] - a[1] * b[0];
  return c;
}
inline V3 head3(const V6& v) { V3 r; r[0] = v[0]; r[1] = v[1]; r[2] = v[2]; return r; }
inline V3 tail3(const V6& v) { V3 r; r[0] = v[3]; r[1] = v[4]; r[2] = v[5]; return r; }

// General inverse by LU with partial pivoting: what Eigen's .inverse() does for dynamic sizes and
// fixed sizes > 4 (PartialPivLU).  Call sites: Adj().inverse() G2oTypes.cc:110,352; Hll^-1
// block_solver.hpp:389; mathematically unique.
template <int N>
inline Mat<N, N> inverse(const Mat<N, N>& A) {
  Mat<N, N> lu = A;
  Mat<N, N> inv = Mat<N, N>::Identity();
  for (int k = 0; k < N; ++k) {
    int p = k;
    double best = std::fabs(lu(k, k));
    for (int r = k + 1; r < N; ++r)
      if (std::fabs(lu(r, k)) > best) { best = std::fabs(lu(r, k)); p = r; }
    if (p != k)
      for (int c = 0; c < N; ++c) {
        double t = lu(k, c); lu(k, c) = lu(p, c); lu(p, c) = t;
        t = inv(k, c); inv(k, c) = inv(p, c); inv(p, c) = t;
      }
    double d = 1.0 / lu(k, k);
    for (int r = k + 1; r < N; ++r) {
      double f = lu(r, k) * d;
      if (f == 0.0) continue;
      for (int c = k; c < N; ++c) lu(r, c) -= f * lu(k, c);
      for (int c = 0; c < N; ++c) inv(r, c) -= f * inv(k, c);
    }
  }
  for (int k = N - 1; k >= 0; --k) {
    double d = 1.0 / lu(k, k);
    for (int c = 0; c < N; ++c) inv(k, c) *= d;
    for (int r = 0; r < k; ++r) {
      double f = lu(r, k);
      if (f == 0.0) continue;
      for (int c = 0; c < N; ++c) inv(r, c) -= f * inv(k, c);
    }
  }
  return inv;
}

// ---------------------------------------------------------------- SO(3) / SE(3)
const double kSophusEps = 1e-10;  // common.hpp:94

inline M3 hat(const V3& w) {  // SO3::hat; also ORB_SLAM3::Skew (G2oTypes.cc:592-597)
  M3 m = M3::Zero();
  m(0, 1) = -w[2]; m(0, 2) = w[1];
  m(1, 0) = w[2];  m(1, 2) = -w[0];
  m(2, 0) = -w[1]; m(2, 1) = w[0];
  return m;
}

struct Quat { double x, y, z, w; };

inline Quat quat_normalized(Quat q) {  // so3.hpp:297-303
  double len = std::sqrt(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
  q.x /= len; q.y /= len; q.z /= len; q.w /= len;
  return q;
}
inline Quat quat_mul(const Quat& a, const Quat& b) {  // so3.hpp:324-338 (result re-normalised by the SO3 ctor :480-487)
  Quat r;
  r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
  r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
  r.y = a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z;
  r.z = a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x;
  return quat_normalized(r);
}
inline Quat quat_inv(const Quat& q) {  // so3.hpp:229-231: SO3(conjugate) -> normalising ctor
  Quat r = {-q.x, -q.y, -q.z, q.w};
  return quat_normalized(r);
}
inline V3 quat_rot(const Quat& q, const V3& p) {  // so3.hpp:355-366
  V3 qv; qv[0] = q.x; qv[1] = q.y; qv[2] = q.z;
  V3 uv = cross3(qv, p);
  uv = uv + uv;
  V3 c = cross3(qv, uv);
  V3 r;
  for (int i = 0; i < 3; ++i) r[i] = p[i] + q.w * uv[i] + c[i];
  return r;
}
inline M3 quat_to_R(const Quat& q) {  // Eigen::QuaternionBase::toRotationMatrix
  const double tx = 2.0 * q.x, ty = 2.0 * q.y, tz = 2.0 * q.z;
  const double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
  const double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
  const double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
  M3 R;
  R(0, 0) = 1.0 - (tyy + tzz); R(0, 1) = txy - twz;         R(0, 2) = txz + twy;
  R(1, 0) = txy + twz;         R(1, 1) = 1.0 - (txx + tzz); R(1, 2) = tyz - twx;
  R(2, 0) = txz - twy;         R(2, 1) = tyz + twx;         R(2, 2) = 1.0 - (txx + tyy);
  return R;
}

inline Quat so3_exp(const V3& omega, double* theta) {  // so3.hpp:583-619
  double theta_sq = dot3(omega, omega);
  double imag, real;
  if (theta_sq < kSophusEps * kSophusEps) {
    *theta = 0.0;
    double theta_po4 = theta_sq * theta_sq;
    imag = 0.5 - (1.0 / 48.0) * theta_sq + (1.0 / 3840.0) * theta_po4;
    real = 1.0 - (1.0 / 8.0) * theta_sq + (1.0 / 384.0) * theta_po4;
  } else {
    *theta = std::sqrt(theta_sq);
    double half = 0.5 * (*theta);
    imag = std::sin(half) / (*theta);
    real = std::cos(half);
  }
  Quat q = {imag * omega[0], imag * omega[1], imag * omega[2], real};
  return q;
}

inline V3 so3_log(const Quat& q, double* theta) {  // so3.hpp:247-291
  double squared_n = q.x * q.x + q.y * q.y + q.z * q.z;
  double w = q.w;
  double two_atan_nbyw_by_n;
  if (squared_n < kSophusEps * kSophusEps) {
    double squared_w = w * w;
    two_atan_nbyw_by_n = 2.0 / w - (2.0 / 3.0) * squared_n / (w * squared_w);
    *theta = 2.0 * squared_n / w;
  } else {
    double n = std::sqrt(squared_n);
    if (std::fabs(w) < kSophusEps) {
      two_atan_nbyw_by_n = (w > 0.0 ? M_PI : -M_PI) / n;
    } else {
      two_atan_nbyw_by_n = 2.0 * std::atan(n / w) / n;
    }
    *theta = two_atan_nbyw_by_n * n;
  }
  V3 t;
  t[0] = two_atan_nbyw_by_n * q.x; t[1] = two_atan_nbyw_by_n * q.y; t[2] = two_atan_nbyw_by_n * q.z;
  return t;
}

struct SE3 {
  Quat q;
  V3 t;
  static SE3 Identity() { SE3 T; T.q = {0, 0, 0, 1}; T.t = V3::Zero(); return T; }
};

inline SE3 se3_mul(const SE3& a, const SE3& b) {  // se3.hpp:304-309
  SE3 r;
  r.q = quat_mul(a.q, b.q);
  r.t = a.t + quat_rot(a.q, b.t);
  return r;
}
inline SE3 se3_inv(const SE3& a) {  // se3.hpp:208-211
  SE3 r;
  r.q = quat_inv(a.q);
  r.t = quat_rot(r.q, -a.t);  // invR * (translation * -1)
  return r;
}
inline V3 se3_act(const SE3& a, const V3& p) { return quat_rot(a.q, p) + a.t; }  // se3.hpp:321-325
inline M3 se3_R(const SE3& a) { return quat_to_R(a.q); }

inline SE3 se3_exp(const V6& a) {  // se3.hpp:761-783
  V3 omega = tail3(a);
  double theta;
  SE3 r;
  r.q = so3_exp(omega, &theta);
  M3 Omega = hat(omega);
  M3 Omega_sq = Omega * Omega;
  M3 V;
  if (theta < kSophusEps) {
    V = quat_to_R(r.q);
  } else {
    double theta_sq = theta * theta;
    V = M3::Identity() + ((1.0 - std::cos(theta)) / theta_sq) * Omega +
        ((theta - std::sin(theta)) / (theta_sq * theta)) * Omega_sq;
  }
  r.t = V * head3(a);
  return r;
}

inline V6 se3_log(const SE3& T) {  // se3.hpp:223-255
  double theta;
  V3 omega = so3_log(T.q, &theta);
  M3 Omega = hat(omega);
  M3 V_inv;
  if (std::fabs(theta) < kSophusEps) {
    V_inv = M3::Identity() - 0.5 * Omega + (1.0 / 12.0) * (Omega * Omega);
  } else {
    double half = 0.5 * theta;
    V_inv = M3::Identity() - 0.5 * Omega +
            ((1.0 - theta * std::cos(half) / (2.0 * std::sin(half))) / (theta * theta)) * (Omega * Omega);
  }
  V3 up = V_inv * T.t;
  V6 r;
  r[0] = up[0]; r[1] = up[1]; r[2] = up[2];
  r[3] = omega[0]; r[4] = omega[1]; r[5] = omega[2];
  return r;
}

inline M6 se3_Adj(const SE3& T) {  // se3.hpp:103-111
  M3 R = se3_R(T);
  M6 A = M6::Zero();
  A.set_block(0, 0, R);
  A.set_block(3, 3, R);
  A.set_block(0, 3, hat(T.t) * R);
  return A;
}

// ---------------------------------------------------------------- Pose3utils (src/Pose3utils.cc)
inline M3 LeftJacobianRot3(const V3& omega) {  // :48-59
  double theta2 = dot3(omega, omega);
  if (theta2 <= DBL_EPSILON) return M3::Identity();
  const double theta = std::sqrt(theta2);
  V3 dir = (1.0 / theta) * omega;
  for (int i = 0; i < 3; ++i) dir[i] = omega[i] / theta;
  const double s = std::sin(theta);
  M3 A = hat(omega);
  for (int i = 0; i < 9; ++i) A.a[i] = A.a[i] / theta;
  M3 ddT = dir * transpose(dir);
  return (s / theta) * M3::Identity() + (1.0 - s / theta) * ddT + ((1.0 - std::cos(theta)) / theta) * A;
}

inline M3 LeftJacobianRot3Inv(const V3& omega) {  // :61-73
  double theta2 = dot3(omega, omega);
  if (theta2 <= DBL_EPSILON) return M3::Identity();
  const double theta = std::sqrt(theta2);
  V3 dir;
  for (int i = 0; i < 3; ++i) dir[i] = omega[i] / theta;
  const double theta_2 = theta / 2.0;
  const double cot_theta_2 = 1.0 / std::tan(theta_2);
  M3 A = hat(omega);
  for (int i = 0; i < 9; ++i) A.a[i] = A.a[i] / theta;
  M3 ddT = dir * transpose(dir);
  return (theta_2 * cot_theta_2) * M3::Identity() + (1.0 - theta_2 * cot_theta_2) * ddT - theta_2 * A;
}

inline M3 LeftJacobianPose3Q(const V6& xi) {  // :5-22
  const V3 omega = tail3(xi), rho = head3(xi);
  const double theta = std::sqrt(dot3(omega, omega));
  const M3 X = hat(omega), Y = hat(rho);
  const M3 XY = X * Y, YX = Y * X, XYX = X * YX;
  if (std::fabs(theta) > 1e-5) {
    const double st = std::sin(theta), ct = std::cos(theta);
    const double t2 = theta * theta, t3 = t2 * theta, t4 = t3 * theta, t5 = t4 * theta;
    return 0.5 * Y + ((theta - st) / t3) * (XY + YX + XYX) -
           ((1.0 - 0.5 * t2 - ct) / t4) * (X * XY + YX * X - 3.0 * XYX) -
           (0.5 * ((1.0 - 0.5 * t2 - ct) / t4 - 3.0 * (theta - st - t3 / 6.0) / t5)) * (XYX * X + X * XYX);
  } else {
    return 0.5 * Y + (1.0 / 6.0) * (XY + YX + XYX) - (1.0 / 24.0) * (X * XY + YX * X - 3.0 * XYX) -
           (0.5 * (1.0 / 24.0 + 3.0 / 120.0)) * (XYX * X + X * XYX);
  }
}

inline M6 LeftJacobianPose3(const V6& xi) {  // :24-30
  const M3 Q = LeftJacobianPose3Q(xi);
  const M3 J = LeftJacobianRot3(tail3(xi));
  M6 r = M6::Zero();
  r.set_block(0, 0, J); r.set_block(0, 3, Q); r.set_block(3, 3, J);
  return r;
}
inline M6 RightJacobianPose3(const V6& xi) { return LeftJacobianPose3(-xi); }  // :32-34
inline M6 LeftJacobianPose3Inv(const V6& xi) {  // :36-42
  const M3 Q = LeftJacobianPose3Q(xi);
  const M3 Jinv = LeftJacobianRot3Inv(tail3(xi));
  M6 r = M6::Zero();
  r.set_block(0, 0, Jinv); r.set_block(0, 3, -(Jinv * Q * Jinv)); r.set_block(3, 3, Jinv);
  return r;
}
inline M6 RightJacobianPose3Inv(const V6& xi) { return LeftJacobianPose3Inv(-xi); }  // :44-46

inline M6 se3Adj(const V6& v) {  // :111-118
  M6 A = M6::Zero();
  A.set_block(0, 0, hat(tail3(v)));
  A.set_block(0, 3, hat(head3(v)));
  A.set_block(3, 3, hat(tail3(v)));
  return A;
}

}  // namespace ora
