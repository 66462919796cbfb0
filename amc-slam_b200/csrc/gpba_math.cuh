// gpba_math.cuh -- device-side SO(3)/SE(3), Pose3utils and GP-interpolation math for libgpba.
//
// Same functions the reference evaluates on the CPU (cited per function, paths relative to the
// AMC-SLAM tree), written for registers: fixed-size row-major matrices, no heap, thresholds kept
// identical so both sides take the same branch.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace gpba {

#define GPBA_HD __host__ __device__ __forceinline__
#define GPBA_D __device__ __forceinline__

template <int R, int C>
struct Mat {
  double a[R * C];
  GPBA_HD double& operator()(int r, int c) { return a[r * C + c]; }
  GPBA_HD double operator()(int r, int c) const { return a[r * C + c]; }
  GPBA_HD double& operator[](int i) { return a[i]; }
  GPBA_HD double operator[](int i) const { return a[i]; }
};
typedef Mat<3, 1> V3;
typedef Mat<6, 1> V6;
typedef Mat<3, 3> M3;
typedef Mat<6, 6> M6;

template <int R, int C>
GPBA_HD Mat<R, C> zeros() {
  Mat<R, C> m;
#pragma unroll
  for (int i = 0; i < R * C; ++i) m.a[i] = 0.0;
  return m;
}
template <int N>
GPBA_HD Mat<N, N> eye() {
  Mat<N, N> m = zeros<N, N>();
#pragma unroll
  for (int i = 0; i < N; ++i) m(i, i) = 1.0;
  return m;
}
template <int R, int K, int C>
GPBA_HD Mat<R, C> mul(const Mat<R, K>& A, const Mat<K, C>& B) {
  Mat<R, C> m;
#pragma unroll
  for (int r = 0; r < R; ++r)
#pragma unroll
    for (int c = 0; c < C; ++c) {
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < K; ++k) s = fma(A(r, k), B(k, c), s);
      m(r, c) = s;
    }
  return m;
}
template <int R, int C>
GPBA_HD Mat<R, C> add(const Mat<R, C>& A, const Mat<R, C>& B) {
  Mat<R, C> m;
#pragma unroll
  for (int i = 0; i < R * C; ++i) m.a[i] = A.a[i] + B.a[i];
  return m;
}
template <int R, int C>
GPBA_HD Mat<R, C> sub(const Mat<R, C>& A, const Mat<R, C>& B) {
  Mat<R, C> m;
#pragma unroll
  for (int i = 0; i < R * C; ++i) m.a[i] = A.a[i] - B.a[i];
  return m;
}
template <int R, int C>
GPBA_HD Mat<R, C> scale(double s, const Mat<R, C>& A) {
  Mat<R, C> m;
#pragma unroll
  for (int i = 0; i < R * C; ++i) m.a[i] = s * A.a[i];
  return m;
}
template <int R, int C>
GPBA_HD Mat<C, R> transpose(const Mat<R, C>& A) {
  Mat<C, R> m;
#pragma unroll
  for (int r = 0; r < R; ++r)
#pragma unroll
    for (int c = 0; c < C; ++c) m(c, r) = A(r, c);
  return m;
}
GPBA_HD double dot3(const V3& a, const V3& b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
GPBA_HD V3 cross3(const V3& a, const V3& b) {
  V3 c;
  c[0] = a[1] * b[2] - a[2] * b[1];
  c[1] = a[2] * b[0] - a[0] * b[2];
  c[2] = a[0] * b[1] - a[1] * b[0];
  return c;
}
GPBA_HD V3 v3(double x, double y, double z) { V3 v; v[0] = x; v[1] = y; v[2] = z; return v; }
GPBA_HD V3 head3(const V6& v) { return v3(v[0], v[1], v[2]); }
GPBA_HD V3 tail3(const V6& v) { return v3(v[3], v[4], v[5]); }
GPBA_HD V6 neg6(const V6& v) { V6 r; for (int i = 0; i < 6; ++i) r[i] = -v[i]; return r; }

GPBA_HD M3 hat(const V3& w) {  // Sophus SO3::hat, ORB_SLAM3::Skew (src/G2oTypes.cc:592-597)
  M3 m = zeros<3, 3>();
  m(0, 1) = -w[2]; m(0, 2) = w[1];
  m(1, 0) = w[2];  m(1, 2) = -w[0];
  m(2, 0) = -w[1]; m(2, 1) = w[0];
  return m;
}
template <int R, int C, int R2, int C2>
GPBA_HD void set_block(Mat<R, C>& M, int r0, int c0, const Mat<R2, C2>& B) {
#pragma unroll
  for (int r = 0; r < R2; ++r)
#pragma unroll
    for (int c = 0; c < C2; ++c) M(r0 + r, c0 + c) = B(r, c);
}

// ------------------------------------------------------------------ SO(3) / SE(3) (Thirdparty/Sophus/sophus)
#define GPBA_SOPHUS_EPS 1e-10  // common.hpp:94

struct Quat { double x, y, z, w; };
struct SE3 { Quat q; V3 t; };

GPBA_HD Quat quat_normalized(Quat q) {  // so3.hpp:297-303
  double len = sqrt(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
  q.x /= len; q.y /= len; q.z /= len; q.w /= len;
  return q;
}
GPBA_HD Quat quat_mul(const Quat& a, const Quat& b) {  // so3.hpp:324-338 (+ normalising ctor :480-487)
  Quat r;
  r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
  r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
  r.y = a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z;
  r.z = a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x;
  return quat_normalized(r);
}
GPBA_HD Quat quat_inv(const Quat& q) {  // so3.hpp:229-231
  Quat r = {-q.x, -q.y, -q.z, q.w};
  return quat_normalized(r);
}
GPBA_HD V3 quat_rot(const Quat& q, const V3& p) {  // so3.hpp:355-366
  V3 qv = v3(q.x, q.y, q.z);
  V3 uv = cross3(qv, p);
  uv = add(uv, uv);
  V3 c = cross3(qv, uv);
  return v3(p[0] + q.w * uv[0] + c[0], p[1] + q.w * uv[1] + c[1], p[2] + q.w * uv[2] + c[2]);
}
GPBA_HD M3 quat_to_R(const Quat& q) {  // Eigen QuaternionBase::toRotationMatrix (so3.hpp:310-312)
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
GPBA_HD Quat so3_exp(const V3& omega, double* theta) {  // so3.hpp:583-619
  double theta_sq = dot3(omega, omega);
  double imag, real;
  if (theta_sq < GPBA_SOPHUS_EPS * GPBA_SOPHUS_EPS) {
    *theta = 0.0;
    double theta_po4 = theta_sq * theta_sq;
    imag = 0.5 - (1.0 / 48.0) * theta_sq + (1.0 / 3840.0) * theta_po4;
    real = 1.0 - (1.0 / 8.0) * theta_sq + (1.0 / 384.0) * theta_po4;
  } else {
    *theta = sqrt(theta_sq);
    double half = 0.5 * (*theta);
    double s, c;
    sincos(half, &s, &c);
    imag = s / (*theta);
    real = c;
  }
  Quat q = {imag * omega[0], imag * omega[1], imag * omega[2], real};
  return q;
}
GPBA_HD V3 so3_log(const Quat& q, double* theta) {  // so3.hpp:247-291
  double squared_n = q.x * q.x + q.y * q.y + q.z * q.z;
  double w = q.w;
  double f;
  if (squared_n < GPBA_SOPHUS_EPS * GPBA_SOPHUS_EPS) {
    double squared_w = w * w;
    f = 2.0 / w - (2.0 / 3.0) * squared_n / (w * squared_w);
    *theta = 2.0 * squared_n / w;
  } else {
    double n = sqrt(squared_n);
    if (fabs(w) < GPBA_SOPHUS_EPS) f = (w > 0.0 ? 3.14159265358979323846 : -3.14159265358979323846) / n;
    else f = 2.0 * atan(n / w) / n;
    *theta = f * n;
  }
  return v3(f * q.x, f * q.y, f * q.z);
}
GPBA_HD SE3 se3_mul(const SE3& a, const SE3& b) {  // se3.hpp:304-309
  SE3 r;
  r.q = quat_mul(a.q, b.q);
  r.t = add(a.t, quat_rot(a.q, b.t));
  return r;
}
GPBA_HD SE3 se3_inv(const SE3& a) {  // se3.hpp:208-211
  SE3 r;
  r.q = quat_inv(a.q);
  r.t = quat_rot(r.q, v3(-a.t[0], -a.t[1], -a.t[2]));
  return r;
}
GPBA_HD V3 se3_act(const SE3& a, const V3& p) { return add(quat_rot(a.q, p), a.t); }  // se3.hpp:321-325
GPBA_HD SE3 se3_exp(const V6& a) {  // se3.hpp:761-783
  V3 omega = tail3(a);
  double theta;
  SE3 r;
  r.q = so3_exp(omega, &theta);
  M3 Omega = hat(omega);
  M3 Omega_sq = mul(Omega, Omega);
  M3 V;
  if (theta < GPBA_SOPHUS_EPS) {
    V = quat_to_R(r.q);
  } else {
    double theta_sq = theta * theta;
    double s, c;
    sincos(theta, &s, &c);
    V = add(add(eye<3>(), scale((1.0 - c) / theta_sq, Omega)), scale((theta - s) / (theta_sq * theta), Omega_sq));
  }
  r.t = mul(V, head3(a));
  return r;
}
GPBA_HD V6 se3_log(const SE3& T) {  // se3.hpp:223-255
  double theta;
  V3 omega = so3_log(T.q, &theta);
  M3 Omega = hat(omega);
  M3 Osq = mul(Omega, Omega);
  M3 V_inv;
  if (fabs(theta) < GPBA_SOPHUS_EPS) {
    V_inv = add(sub(eye<3>(), scale(0.5, Omega)), scale(1.0 / 12.0, Osq));
  } else {
    double half = 0.5 * theta;
    double s, c;
    sincos(half, &s, &c);
    V_inv = add(sub(eye<3>(), scale(0.5, Omega)), scale((1.0 - theta * c / (2.0 * s)) / (theta * theta), Osq));
  }
  V3 up = mul(V_inv, T.t);
  V6 r;
  r[0] = up[0]; r[1] = up[1]; r[2] = up[2]; r[3] = omega[0]; r[4] = omega[1]; r[5] = omega[2];
  return r;
}
GPBA_HD M6 se3_Adj(const SE3& T) {  // se3.hpp:103-111
  M3 R = quat_to_R(T.q);
  M6 A = zeros<6, 6>();
  set_block(A, 0, 0, R);
  set_block(A, 3, 3, R);
  set_block(A, 0, 3, mul(hat(T.t), R));
  return A;
}

// ------------------------------------------------------------------ Pose3utils (src/Pose3utils.cc)
GPBA_HD M3 LeftJacobianRot3(const V3& omega) {  // :48-59
  double theta2 = dot3(omega, omega);
  if (theta2 <= 2.220446049250313e-16) return eye<3>();
  const double theta = sqrt(theta2);
  V3 dir = v3(omega[0] / theta, omega[1] / theta, omega[2] / theta);
  double s, c;
  sincos(theta, &s, &c);
  M3 A = hat(omega);
#pragma unroll
  for (int i = 0; i < 9; ++i) A.a[i] = A.a[i] / theta;
  M3 r;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      r(i, j) = (i == j ? s / theta : 0.0) + (1.0 - s / theta) * (dir[i] * dir[j]) + ((1.0 - c) / theta) * A(i, j);
  return r;
}
GPBA_HD M3 LeftJacobianRot3Inv(const V3& omega) {  // :61-73
  double theta2 = dot3(omega, omega);
  if (theta2 <= 2.220446049250313e-16) return eye<3>();
  const double theta = sqrt(theta2);
  V3 dir = v3(omega[0] / theta, omega[1] / theta, omega[2] / theta);
  const double theta_2 = theta / 2.0;
  const double cot_theta_2 = 1.0 / tan(theta_2);
  M3 A = hat(omega);
#pragma unroll
  for (int i = 0; i < 9; ++i) A.a[i] = A.a[i] / theta;
  M3 r;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      r(i, j) = (i == j ? theta_2 * cot_theta_2 : 0.0) + (1.0 - theta_2 * cot_theta_2) * (dir[i] * dir[j]) - theta_2 * A(i, j);
  return r;
}
GPBA_HD M3 LeftJacobianPose3Q(const V6& xi) {  // :5-22
  const V3 omega = tail3(xi), rho = head3(xi);
  const double theta = sqrt(dot3(omega, omega));
  const M3 X = hat(omega), Y = hat(rho);
  const M3 XY = mul(X, Y), YX = mul(Y, X), XYX = mul(X, YX);
  const M3 t1 = add(add(XY, YX), XYX);
  const M3 t2 = sub(add(mul(X, XY), mul(YX, X)), scale(3.0, XYX));
  const M3 t3 = add(mul(XYX, X), mul(X, XYX));
  double c1, c2, c3;
  if (fabs(theta) > 1e-5) {
    double st, ct;
    sincos(theta, &st, &ct);
    const double th2 = theta * theta, th3 = th2 * theta, th4 = th3 * theta, th5 = th4 * theta;
    c1 = (theta - st) / th3;
    c2 = (1.0 - 0.5 * th2 - ct) / th4;
    c3 = 0.5 * ((1.0 - 0.5 * th2 - ct) / th4 - 3.0 * (theta - st - th3 / 6.0) / th5);
  } else {
    c1 = 1.0 / 6.0; c2 = 1.0 / 24.0; c3 = 0.5 * (1.0 / 24.0 + 3.0 / 120.0);
  }
  return sub(sub(add(scale(0.5, Y), scale(c1, t1)), scale(c2, t2)), scale(c3, t3));
}
GPBA_HD M6 LeftJacobianPose3(const V6& xi) {  // :24-30
  M6 r = zeros<6, 6>();
  M3 J = LeftJacobianRot3(tail3(xi));
  set_block(r, 0, 0, J);
  set_block(r, 0, 3, LeftJacobianPose3Q(xi));
  set_block(r, 3, 3, J);
  return r;
}
GPBA_HD M6 RightJacobianPose3(const V6& xi) { return LeftJacobianPose3(neg6(xi)); }  // :32-34
GPBA_HD M6 LeftJacobianPose3Inv(const V6& xi) {  // :36-42
  M6 r = zeros<6, 6>();
  M3 Jinv = LeftJacobianRot3Inv(tail3(xi));
  M3 Q = LeftJacobianPose3Q(xi);
  set_block(r, 0, 0, Jinv);
  set_block(r, 0, 3, scale(-1.0, mul(mul(Jinv, Q), Jinv)));
  set_block(r, 3, 3, Jinv);
  return r;
}
GPBA_HD M6 RightJacobianPose3Inv(const V6& xi) { return LeftJacobianPose3Inv(neg6(xi)); }  // :44-46
GPBA_HD M6 se3Adj(const V6& v) {  // :111-118
  M6 A = zeros<6, 6>();
  set_block(A, 0, 0, hat(tail3(v)));
  set_block(A, 0, 3, hat(head3(v)));
  set_block(A, 3, 3, hat(tail3(v)));
  return A;
}

// ------------------------------------------------------------------ GP interpolation
// Closed form of At1 / Pt1: the reference's Qi(t-t1) Transition(t,t2)^T QiInv(t2-t1)
// and Transition(t1,t) - Pt Transition(t1,t2) (src/GaussianProcess.cc:27-30) reduce to scalar
// multiples of identity blocks, independent of Qc.
struct GpWeights { double l11, l12, p11, p12; };
GPBA_HD GpWeights gp_weights(double t1, double t2, double t) {
  const double D = t2 - t1;
  const double s = (t - t1) / D;
  GpWeights w;
  w.p11 = 3.0 * s * s - 2.0 * s * s * s;
  w.p12 = D * (s * s * s - s * s);
  w.l11 = 1.0 - w.p11;
  w.l12 = D * (s - 2.0 * s * s + s * s * s);
  return w;
}

GPBA_HD SE3 load_se3(const double* p) {
  SE3 T;
  T.q.x = p[0]; T.q.y = p[1]; T.q.z = p[2]; T.q.w = p[3];
  T.t[0] = p[4]; T.t[1] = p[5]; T.t[2] = p[6];
  return T;
}
GPBA_HD void store_se3(const SE3& T, double* p) {
  p[0] = T.q.x; p[1] = T.q.y; p[2] = T.q.z; p[3] = T.q.w;
  p[4] = T.t[0]; p[5] = T.t[1]; p[6] = T.t[2];
}
GPBA_HD V6 load_v6(const double* p) { V6 v; for (int i = 0; i < 6; ++i) v[i] = p[i]; return v; }

// RobustKernelHuber with float dsqr (robust_kernel_impl.h:84, .cpp:78-91): returns rho, sets *rho1 = rho'
GPBA_HD double huber(double e, double delta, double dsqr, double* rho1) {
  if (e <= dsqr) { *rho1 = 1.0; return e; }
  double sqrte = sqrt(e);
  *rho1 = delta / sqrte;
  return 2 * sqrte * delta - dsqr;
}

}  // namespace gpba
