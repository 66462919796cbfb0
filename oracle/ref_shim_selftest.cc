// TEST INFRASTRUCTURE: exercises the stand-in headers under oracle/ref_shim/ (Eigen, Sophus) through the same expression
// forms the reference's sources use, so that tests/test_ref_shim.py can compare them with numpy / scipy.  The pin of the
// oracle against the reference's code (oracle/ref_pin.cc) is only as good as this arithmetic.  Needs nothing from
// /root/reference; built by `make -C oracle selftest` into oracle/_ref/libref_shim_selftest.so.
#include <Eigen/Dense>
#include "sophus/se3.hpp"

using Eigen::MatrixXd;

namespace {
MatrixXd load(const double* p, int r, int c) {
  MatrixXd m(r, c);
  for (int i = 0; i < r; ++i) for (int j = 0; j < c; ++j) m(i, j) = p[i * c + j];
  return m;
}
template <class M> void store(const M& m, double* o) {
  for (int i = 0; i < m.rows(); ++i) for (int j = 0; j < m.cols(); ++j) o[i * m.cols() + j] = m(i, j);
}
Sophus::SE3d from7(const double* p) {
  return Sophus::SE3d(Sophus::SO3d::fromQuaternion(p[0], p[1], p[2], p[3]), Eigen::Vector3d(p[4], p[5], p[6]));
}
void to7(const Sophus::SE3d& T, double* p) {
  p[0] = T.so3().qx(); p[1] = T.so3().qy(); p[2] = T.so3().qz(); p[3] = T.so3().qw();
  for (int i = 0; i < 3; ++i) p[4 + i] = T.translation()(i);
}
}  // namespace

extern "C" {

// out = a * A * B^T + (C - D) / b - (-E), every operand n x n (dynamic types)
void shim_expr_dynamic(int n, const double* A, const double* B, const double* C, const double* D, const double* E, double a,
                       double b, double* out) {
  const MatrixXd r = a * load(A, n, n) * load(B, n, n).transpose() + (load(C, n, n) - load(D, n, n)) / b - (-load(E, n, n));
  store(r, out);
}
void shim_inverse(int n, const double* A, double* out) { store(load(A, n, n).inverse(), out); }
// the fixed-size forms of G2oTypes.cc: 6x6 inverse, 2x3 * 3x6 product, comma initialisers with blocks, block write-through,
// mixed fixed / dynamic vectors, Map<const>.  out (12 x 12) is assembled the way EdgeGaussianPrior::linearizeOplus does.
void shim_fixed_forms(const double* A66, const double* B66, const double* v6, double dt, double* out144, double* out12) {
  Eigen::Matrix<double, 6, 6> A, B;
  for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) { A(i, j) = A66[6 * i + j]; B(i, j) = B66[6 * i + j]; }
  Eigen::Map<const Eigen::Matrix<double, 6, 1>> v(v6);
  Eigen::Matrix<double, 12, 12> J;
  J.block<6, 6>(0, 0) = -A * B.inverse();
  J.block<6, 6>(6, 0) = -0.5 * B * J.block<6, 6>(0, 0);
  J.block<12, 6>(0, 6) << -dt * Eigen::Matrix<double, 6, 6>::Identity(), -Eigen::Matrix<double, 6, 6>::Identity();
  store(J, out144);
  Eigen::VectorXd x(12);
  x.block<6, 1>(0, 0).setZero();
  x.block<6, 1>(6, 0) = v;
  const Eigen::VectorXd y = A * v;                       // fixed * Map -> dynamic
  Eigen::Matrix<double, 12, 1> z;
  z.head<6>() = y;
  z.tail<6>() = (Eigen::Matrix<double, 6, 6>() << A.block<3, 3>(0, 0), B.block<3, 3>(0, 3), Eigen::Matrix3d::Zero(),
                 A.block<3, 3>(3, 3)).finished() * x.tail<6>();
  z(3) += x.dot(x) + v.norm();
  store(z, out12);
}
// row-wise scalar comma initialiser, col() assignment, 3x4 / 4x6 shapes, transpose of a non-square matrix
void shim_small_forms(const double* w3, double* skew9, double* cols18, double* t12) {
  Eigen::Matrix3d W;
  W << 0.0, -w3[2], w3[1], w3[2], 0.0, -w3[0], -w3[1], w3[0], 0.0;
  store(W, skew9);
  Eigen::Matrix<double, 3, 6> D = Eigen::Matrix<double, 3, 6>::Zero();
  for (int i = 0; i < 6; ++i) D.col(i) = (W * Eigen::Vector3d(w3[0], w3[1], w3[2]) + Eigen::Vector3d(i, 2 * i, 3 * i)) * (i + 1.0) / 2;
  store(D, cols18);
  Eigen::Matrix<double, 3, 4> T;
  T << W, Eigen::Vector3d(w3[0], w3[1], w3[2]);
  store(T.transpose(), t12);
}
void shim_svd3(const double* A9, double* U9, double* V9, double* s3) {
  Eigen::Matrix3d A;
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) A(i, j) = A9[3 * i + j];
  Eigen::JacobiSVD<Eigen::Matrix3d> svd(A, Eigen::ComputeFullU | Eigen::ComputeFullV);
  store(svd.matrixU(), U9); store(svd.matrixV(), V9); store(svd.singularValues(), s3);
}
void shim_quat(const double* R9, double* q4, double* Rback9, const double* p3, double* rot3) {
  Eigen::Matrix3d R;
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) R(i, j) = R9[3 * i + j];
  const Eigen::Quaterniond q(R);
  store(q.coeffs(), q4); store(q.toRotationMatrix(), Rback9);
  store(q * Eigen::Vector3d(p3[0], p3[1], p3[2]), rot3);
}
void shim_se3(const double* xi6, double* T7, double* M16, double* Adj36, double* log6, double* inv7) {
  Eigen::Matrix<double, 6, 1> xi;
  for (int i = 0; i < 6; ++i) xi(i) = xi6[i];
  const Sophus::SE3d T = Sophus::SE3d::exp(xi);
  to7(T, T7); store(T.matrix(), M16); store(T.Adj(), Adj36); store(T.log(), log6); to7(T.inverse(), inv7);
}
void shim_se3_mul_act(const double* a7, const double* b7, const double* p3, double* ab7, double* ap3) {
  to7(from7(a7) * from7(b7), ab7);
  store(from7(a7) * Eigen::Vector3d(p3[0], p3[1], p3[2]), ap3);
}

// what the g2o core uses beyond the edge sources: pivoted LDLT (linear_solver_dense.h:105-110), llt().solve and determinant
// (base_vertex.hpp:38-41), symmetric eigenvalues (optimizable_graph.cpp:870-880)
int shim_ldlt(int n, const double* A, const double* b, double* x) {
  Eigen::LDLT<MatrixXd> ch;
  ch.compute(load(A, n, n));
  Eigen::VectorXd::ConstMapType bvec(b, n);
  Eigen::VectorXd::MapType xvec(x, n);
  xvec = ch.solve(bvec);
  return ch.isPositive() ? 1 : 0;
}
double shim_llt_det(int n, const double* A, const double* b, double* x) {
  const MatrixXd M = load(A, n, n);
  Eigen::VectorXd bv(n);
  for (int i = 0; i < n; ++i) bv(i) = b[i];
  store(M.llt().solve(bv), x);
  return M.determinant();
}
void shim_eigenvalues(int n, const double* A, double* ev) {
  Eigen::SelfAdjointEigenSolver<MatrixXd> es;
  es.compute(load(A, n, n), Eigen::EigenvaluesOnly);
  store(es.eigenvalues(), ev);
}
// Map as a view over foreign memory, the way block_solver.hpp / base_*_edge.hpp use it: segment += on a mapped vector
// (matrix_operations.h:38-41), noalias() += on a mapped block, diagonal().array() += lambda and diagonal() = backup
// (block_solver.hpp:576-602), assignment through a Map of a fixed type (base_multi_edge.hpp), copy of a Map = the same view.
void shim_map_views(int n, double* vec, double* mat, const double* A, double lambda, double* diag_backup) {
  Eigen::Map<Eigen::VectorXd> y(vec, n);
  Eigen::Map<const Eigen::VectorXd> yc(vec, n);
  const MatrixXd M = load(A, n, n);
  const Eigen::VectorXd x0 = yc;                         // a value copy of the mapped data
  y.segment(0, n) += M * x0;                             // vec <- vec + A vec
  y.segment<2>(1) += Eigen::Vector2d(10.0, 20.0);
  Eigen::Map<MatrixXd> H(mat, n, n);
  Eigen::Map<MatrixXd> H2(H);                            // same memory
  H2.noalias() += M.transpose() * M;                     // mat <- mat + A^T A
  store(Eigen::VectorXd(H.diagonal()), diag_backup);
  H.diagonal().array() += lambda;
  H.block(0, 0, 2, 2) = MatrixXd(Eigen::Matrix2d::Identity() * 7.0);
  Eigen::Map<Eigen::Matrix<double, 1, 2>> last2(mat + (size_t)n * n - 2);   // a fixed-size view: the last two entries of the last row
  last2 = Eigen::Matrix<double, 1, 2>(Eigen::Vector2d(-1.0, -2.0).transpose());
}

}  // extern "C"
