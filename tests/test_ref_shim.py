"""The stand-in headers under oracle/ref_shim/ (Eigen, Sophus) against numpy / scipy.

oracle/_ref compiles the reference's own sources against these stand-ins (tests/test_ref_pin.py); the pin is only as good as
their arithmetic, so every expression form the reference's three source files use is exercised here through
oracle/ref_shim_selftest.cc: dynamic and fixed products / sums / scalar forms, inverses, comma initialisers with blocks, block
write-through, head / tail / col views, Map<const>, the 3x3 SVD, quaternions, and the Sophus stand-in's exp / log / Adj /
inverse / product / action against scipy's expm / logm.  Needs no reference and no GPU.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
import scipy.linalg as sl

ORACLE = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle")


@pytest.fixture(scope="module")
def S():
    subprocess.check_call(["make", "-C", ORACLE, "-s", "selftest"])
    return C.CDLL(os.path.join(ORACLE, "_ref", "libref_shim_selftest.so"))


def p(a):
    return a.ctypes.data_as(C.c_void_p)


def d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def hat(w):
    return np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])


def hat6(xi):
    M = np.zeros((4, 4)); M[:3, :3] = hat(xi[3:]); M[:3, 3] = xi[:3]
    return M


def quat_R(q):
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def T4(T7):
    M = np.eye(4); M[:3, :3] = quat_R(T7[:4]); M[:3, 3] = T7[4:]
    return M


def test_dynamic_expression_and_inverse(S):
    rng = np.random.default_rng(0)
    for n in (1, 2, 3, 6, 12, 24):
        A, B, Cc, D, E = (rng.normal(size=(n, n)) for _ in range(5))
        out = np.zeros((n, n))
        S.shim_expr_dynamic(n, p(d(A)), p(d(B)), p(d(Cc)), p(d(D)), p(d(E)), C.c_double(1.7), C.c_double(-0.3), p(out))
        np.testing.assert_allclose(out, 1.7 * A @ B.T + (Cc - D) / -0.3 + E, rtol=0, atol=1e-13 * n)
        M = A + n * np.eye(n)
        S.shim_inverse(n, p(d(M)), p(out))
        np.testing.assert_allclose(out, np.linalg.inv(M), rtol=0, atol=1e-13)
    # pivoting: a leading zero
    M = np.array([[0.0, 2.0, 1.0], [1.0, 0.0, 0.0], [3.0, 1.0, 0.0]]); out = np.zeros((3, 3))
    S.shim_inverse(3, p(d(M)), p(out))
    np.testing.assert_allclose(out, np.linalg.inv(M), atol=1e-14)


def test_fixed_forms_blocks_comma_map(S):
    rng = np.random.default_rng(1)
    for _ in range(20):
        A = rng.normal(size=(6, 6)); B = rng.normal(size=(6, 6)) + 4 * np.eye(6); v = rng.normal(size=6); dt = rng.uniform(0.1, 1)
        J = np.zeros((12, 12)); z = np.zeros(12)
        S.shim_fixed_forms(p(d(A)), p(d(B)), p(d(v)), C.c_double(dt), p(J), p(z))
        E = np.zeros((12, 12))
        E[:6, :6] = -A @ np.linalg.inv(B)
        E[6:, :6] = -0.5 * B @ E[:6, :6]
        E[:6, 6:] = -dt * np.eye(6); E[6:, 6:] = -np.eye(6)
        np.testing.assert_allclose(J, E, atol=1e-12)
        x = np.concatenate([np.zeros(6), v])
        K = np.block([[A[:3, :3], B[:3, 3:]], [np.zeros((3, 3)), A[3:, 3:]]])
        ze = np.concatenate([A @ v, K @ x[6:]])
        ze[3] += x @ x + np.linalg.norm(v)
        np.testing.assert_allclose(z, ze, atol=1e-12)


def test_small_forms(S):
    rng = np.random.default_rng(2)
    w = rng.normal(size=3)
    sk = np.zeros((3, 3)); cols = np.zeros((3, 6)); t = np.zeros((4, 3))
    S.shim_small_forms(p(d(w)), p(sk), p(cols), p(t))
    assert np.array_equal(sk, hat(w))
    exp_cols = np.stack([(hat(w) @ w + np.array([i, 2 * i, 3 * i])) * (i + 1.0) / 2 for i in range(6)], axis=1)
    np.testing.assert_allclose(cols, exp_cols, atol=1e-14)
    np.testing.assert_allclose(t, np.hstack([hat(w), w[:, None]]).T, atol=0)


def test_svd3(S):
    rng = np.random.default_rng(3)
    for i in range(20):
        A = rng.normal(size=(3, 3)) if i % 2 else sl.expm(hat(rng.normal(size=3))) + 1e-6 * rng.normal(size=(3, 3))
        U = np.zeros((3, 3)); V = np.zeros((3, 3)); s = np.zeros(3)
        S.shim_svd3(p(d(A)), p(U), p(V), p(s))
        np.testing.assert_allclose(U @ np.diag(s) @ V.T, A, atol=1e-12)
        np.testing.assert_allclose(U.T @ U, np.eye(3), atol=1e-12); np.testing.assert_allclose(V.T @ V, np.eye(3), atol=1e-12)
        np.testing.assert_allclose(np.sort(s), np.sort(np.linalg.svd(A, compute_uv=False)), atol=1e-12)


def test_quaternion(S):
    rng = np.random.default_rng(4)
    for i in range(40):
        w = rng.normal(size=3) * (1.0 if i % 4 else 3.0)
        R = sl.expm(hat(w))
        q = np.zeros(4); Rb = np.zeros((3, 3)); pt = rng.normal(size=3); r = np.zeros(3)
        S.shim_quat(p(d(R)), p(q), p(Rb), p(d(pt)), p(r))
        np.testing.assert_allclose(np.linalg.norm(q), 1.0, atol=1e-13)
        np.testing.assert_allclose(Rb, R, atol=1e-13); np.testing.assert_allclose(quat_R(q), R, atol=1e-13)
        np.testing.assert_allclose(r, R @ pt, atol=1e-13)
    # 180 degrees about each axis: the branches of the matrix -> quaternion rule
    for ax in range(3):
        w = np.zeros(3); w[ax] = np.pi
        R = sl.expm(hat(w)); q = np.zeros(4); Rb = np.zeros((3, 3)); r = np.zeros(3)
        S.shim_quat(p(d(R)), p(q), p(Rb), p(d(np.ones(3))), p(r))
        np.testing.assert_allclose(Rb, R, atol=1e-13)


def test_sophus_standin_vs_scipy(S):
    rng = np.random.default_rng(5)
    for i in range(60):
        xi = np.concatenate([rng.normal(size=3) * 2, rng.normal(size=3) * [1.0, 1e-3, 1e-12, 2.8][i % 4]])   # 1e-12: below Sophus' epsilon, series branch
        # (between 1e-10 and ~1e-7 Sophus' closed form loses digits to (1 - cos theta) / theta^2; the stand-in follows it, so
        # that range is compared with the oracle in tests/test_ref_pin.py, not with scipy)
        T7 = np.zeros(7); M = np.zeros((4, 4)); Ad = np.zeros((6, 6)); lg = np.zeros(6); inv7 = np.zeros(7)
        S.shim_se3(p(d(xi)), p(T7), p(M), p(Ad), p(lg), p(inv7))
        E = sl.expm(hat6(xi))
        np.testing.assert_allclose(M, E, atol=1e-12); np.testing.assert_allclose(T4(T7), E, atol=1e-12)
        np.testing.assert_allclose(sl.expm(hat6(lg)), E, atol=1e-9)          # beyond pi the log is the wrapped twist
        if np.linalg.norm(xi[3:]) < 3.0:
            np.testing.assert_allclose(lg, xi, atol=1e-9)
        np.testing.assert_allclose(T4(inv7), np.linalg.inv(E), atol=1e-12)
        R, t = E[:3, :3], E[:3, 3]
        np.testing.assert_allclose(Ad, np.block([[R, hat(t) @ R], [np.zeros((3, 3)), R]]), atol=1e-12)
        xi2 = rng.normal(size=6); T7b = np.zeros(7)
        S.shim_se3(p(d(xi2)), p(T7b), p(M), p(Ad), p(lg), p(inv7))
        ab = np.zeros(7); pt = rng.normal(size=3); ap = np.zeros(3)
        S.shim_se3_mul_act(p(T7), p(T7b), p(d(pt)), p(ab), p(ap))
        np.testing.assert_allclose(T4(ab), E @ sl.expm(hat6(xi2)), atol=1e-12)
        np.testing.assert_allclose(ap, R @ pt + t, atol=1e-12)


def test_ldlt_llt_determinant_eigenvalues(S):
    S.shim_llt_det.restype = C.c_double
    rng = np.random.default_rng(6)
    for n in (1, 3, 7, 24, 60):
        B = rng.normal(size=(n, n)); A = B @ B.T + 0.5 * np.eye(n); b = rng.normal(size=n)
        x = np.zeros(n)
        assert S.shim_ldlt(n, p(d(A)), p(d(b)), p(x)) == 1
        np.testing.assert_allclose(x, np.linalg.solve(A, b), rtol=1e-9, atol=1e-11)
        x2 = np.zeros(n)
        det = S.shim_llt_det(n, p(d(A)), p(d(b)), p(x2))
        np.testing.assert_allclose(x2, np.linalg.solve(A, b), rtol=1e-9, atol=1e-11)
        np.testing.assert_allclose(det, np.linalg.det(A), rtol=1e-9)
        ev = np.zeros(n); S.shim_eigenvalues(n, p(d(A)), p(ev))
        np.testing.assert_allclose(ev, np.linalg.eigvalsh(A), rtol=1e-9, atol=1e-11)
    # an indefinite matrix: still solved (pivoted), reported as not positive (g2o's LinearSolverDense then returns false)
    A = np.diag([2.0, -1.0, 3.0]) + 0.1; b = np.ones(3); x = np.zeros(3)
    assert S.shim_ldlt(3, p(d(A)), p(d(b)), p(x)) == 0
    np.testing.assert_allclose(x, np.linalg.solve(A, b), atol=1e-12)


def test_map_views_write_through(S):
    rng = np.random.default_rng(7)
    n = 5
    A = rng.normal(size=(n, n)); v = rng.normal(size=n); M = rng.normal(size=(n, n))
    vec = d(v.copy()); mat = d(M.copy()); backup = np.zeros(n)
    S.shim_map_views(n, p(vec), p(mat), p(d(A)), C.c_double(0.25), p(backup))
    ev = v + A @ v; ev[1] += 10.0; ev[2] += 20.0
    np.testing.assert_allclose(vec, ev, atol=1e-13)
    E = M + A.T @ A
    np.testing.assert_allclose(backup, np.diag(E), atol=1e-13)
    E[np.arange(n), np.arange(n)] += 0.25
    E[:2, :2] = 7.0 * np.eye(2)
    E[n - 1, n - 2:] = [-1.0, -2.0]
    np.testing.assert_allclose(mat, E, atol=1e-13)
