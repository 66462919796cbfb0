"""Pins the CPU oracle's Lie / GP / edge layer (the reference has no tests of its own for this path).

Sources of truth, none of which is the oracle itself:
  * scipy expm/logm + power series (oracle/numpy_mirror.py);
  * the property tests of the vendored Sophus (Thirdparty/Sophus/test/core/tests.hpp:46 adjoint,
    :209 exp(log), :221 group action; element sets test_se3.cpp:32-50);
  * g2o's central-difference Jacobian scheme (base_multi_edge.hpp:62-126);
  * the float-typed constants of SURVEY Appendix C.
"""
import numpy as np
import pytest

import numpy_mirror as M


def sophus_se3_elements():
    """Element set of Thirdparty/Sophus/test/core/test_se3.cpp:32-50 as (omega, translation)."""
    pi = np.pi
    return [((0.2, 0.5, 0.0), (0, 0, 0)), ((0.2, 0.5, -1.0), (10, 0, 0)), ((0, 0, 0), (0, 100, 5)),
            ((0, 0, 0.00001), (0, 0, 0)), ((0, 0, 0.00001), (0, -0.00000001, 0.0000000001)),
            ((0, 0, 0.00001), (0.01, 0, 0)), ((pi, 0, 0), (4, -5, 0)),
            ((0.2, 0.5, 0.0), (0, 0, 0)), ((0.3, 0.5, 0.1), (2, 0, -7))]


def rand_xi(rng, scale_t=2.0, scale_r=1.0):
    return np.concatenate([rng.normal(size=3) * scale_t, rng.normal(size=3) * scale_r])


def test_se3_exp_log_vs_scipy(oracle_mod):
    O = oracle_mod
    rng = np.random.default_rng(0)
    for i in range(50):
        xi = rand_xi(rng, 2.0, 0.8 if i % 2 else 1e-3)
        T = O.se3_matrix(O.se3_exp(xi))
        np.testing.assert_allclose(T, M.exp_se3(xi), atol=1e-12)
        np.testing.assert_allclose(O.se3_log(O.se3_exp(xi)), xi, atol=1e-10)
        np.testing.assert_allclose(O.se3_log(O.se3_exp(xi)), M.log_se3(M.exp_se3(xi)), atol=1e-9)


def test_sophus_properties(oracle_mod):
    O = oracle_mod
    rng = np.random.default_rng(1)
    elems = []
    for om, t in sophus_se3_elements():
        q = O.se3_exp(np.array([0, 0, 0, *om]))
        q[4:] = t
        elems.append(q)
    for T in elems:
        Tm = O.se3_matrix(T)
        # tests.hpp:209  exp(log(T)) == T
        np.testing.assert_allclose(O.se3_matrix(O.se3_exp(O.se3_log(T))), Tm, atol=1e-9)
        # tests.hpp:46   hat(Ad_T x) == T hat(x) T^-1
        x = rand_xi(rng)
        lhs = M.hat6(O.se3_adj(T) @ x)
        rhs = Tm @ M.hat6(x) @ np.linalg.inv(Tm)
        np.testing.assert_allclose(lhs, rhs, atol=1e-8 * max(1.0, np.abs(rhs).max()))
        # tests.hpp:221  group action == matrix action; inverse; product
        p = rng.normal(size=3) * 5
        np.testing.assert_allclose(O.se3_act(T, p), (Tm @ np.append(p, 1))[:3], atol=1e-10 * max(1, np.abs(Tm).max()))
        np.testing.assert_allclose(O.se3_matrix(O.se3_inv(T)), np.linalg.inv(Tm), atol=1e-9 * max(1, np.abs(Tm).max()))
        for U in elems[:4]:
            np.testing.assert_allclose(O.se3_matrix(O.se3_mul(T, U)), Tm @ O.se3_matrix(U), atol=1e-9 * max(1, np.abs(Tm).max()))


def test_pose3_jacobians_vs_series(oracle_mod):
    O = oracle_mod
    rng = np.random.default_rng(2)
    for i in range(40):
        xi = rand_xi(rng, 1.5, [1.0, 0.1, 1e-3, 1e-7][i % 4])
        Jl = M.Jl_series(xi)
        np.testing.assert_allclose(O.jac_pose3(xi, 0), Jl, atol=2e-9)
        np.testing.assert_allclose(O.jac_pose3(xi, 1), M.Jr_series(xi), atol=2e-9)
        np.testing.assert_allclose(O.jac_pose3(xi, 2), np.linalg.inv(Jl), atol=2e-8)
        np.testing.assert_allclose(O.jac_pose3(xi, 3), np.linalg.inv(M.Jr_series(xi)), atol=2e-8)
        np.testing.assert_allclose(O.jac_pose3(xi, 4), M.ad(xi), atol=0)


def test_gp_query_pose(oracle_mod):
    O = oracle_mod
    rng = np.random.default_rng(3)
    qc = np.array([0.02, 0.02, 0.02, 0.002, 0.002, 0.002])
    for _ in range(20):
        T1 = O.se3_exp(rand_xi(rng, 3, 0.5))
        v1 = np.array([4, 0, 0, 0, 0, 0.1]) + rng.normal(size=6) * 0.2
        v2 = v1 + rng.normal(size=6) * 0.05
        t1 = rng.uniform(0, 10); t2 = t1 + rng.uniform(0.05, 0.3)
        T2 = O.se3_mul(T1, O.se3_exp((t2 - t1) * v1 + rng.normal(size=6) * 0.01))
        t = t1 + rng.uniform(0.1, 0.9) * (t2 - t1)
        Tq, At1, Pt1 = O.query_pose(qc, T1, T2, v1, v2, t1, t2, t)
        l11, l12, p11, p12 = M.gp_weights(t1, t2, t)  # Qc cancels
        np.testing.assert_allclose(Pt1, np.hstack([p11 * np.eye(6), p12 * np.eye(6)]), atol=1e-9)
        np.testing.assert_allclose(At1, np.hstack([l11 * np.eye(6), l12 * np.eye(6)]), atol=1e-9)
        ref = M.query_pose(O.se3_matrix(T1), O.se3_matrix(T2), v1, v2, t1, t2, t)
        np.testing.assert_allclose(O.se3_matrix(Tq), ref, atol=1e-9)


def _edge_setup(rng, O):
    qc = np.array([0.02, 0.02, 0.02, 0.002, 0.002, 0.002])
    T1 = O.se3_exp(rand_xi(rng, 3, 0.4))
    v1 = np.array([4, 0, 0, 0, 0, 0.1]) + rng.normal(size=6) * 0.2
    v2 = v1 + rng.normal(size=6) * 0.05
    t1, t2 = 1.0, 1.1
    T2 = O.se3_mul(T1, O.se3_exp((t2 - t1) * v1 + rng.normal(size=6) * 0.01))
    t = 1.0 + rng.uniform(0.01, 0.09)
    Tbc = O.se3_exp(np.array([0.4, 0.1, 1.2, -1.2, 1.2, -1.2]) + rng.normal(size=6) * 0.05)
    intr = np.array([500.0, 501.25, 480.0, 300.0])
    Xc = np.array([rng.uniform(-3, 3), rng.uniform(-2, 2), rng.uniform(4, 30)])
    Tq, _, _ = O.query_pose(qc, T1, T2, v1, v2, t1, t2, t)
    Xw = O.se3_act(O.se3_mul(Tq, Tbc), Xc)
    return qc, T1, v1, t1, T2, v2, t2, t, Tbc, intr, Xw


@pytest.mark.parametrize("stereo", [False, True])
def test_edge_error_and_jacobian(oracle_mod, stereo):
    """Error vs the scipy mirror; Jacobian vs central differences (g2o's numeric scheme). Velocity and
    point blocks are exact derivatives; pose blocks carry the reference's first-order approximation
    (-0.5 ad(v2), src/G2oTypes.cc:351-357) so they only agree to ~1e-4 relative."""
    O = oracle_mod
    rng = np.random.default_rng(4 + stereo)
    for _ in range(6):
        qc, T1, v1, t1, T2, v2, t2, t, Tbc, intr, Xw = _edge_setup(rng, O)
        bf = 501.7
        obs = np.array([470.0, 310.0, 455.0 if stereo else -1.0])
        err, J1, J2, Jp = O.edge_eval(qc, 1, T1, v1, t1, T2, v2, t2, t, Tbc, intr, bf, Xw, obs)
        uv, z = M.reproj(O.se3_matrix(T1), O.se3_matrix(T2), v1, v2, t1, t2, t, O.se3_matrix(Tbc), intr, Xw)
        np.testing.assert_allclose(err[:2], obs[:2] - uv, atol=1e-7)
        if stereo:
            np.testing.assert_allclose(err[2], obs[2] - (uv[0] - bf / z), atol=1e-7)

        def f(T1_, v1_, T2_, v2_, X_):
            return O.edge_eval(qc, 1, T1_, v1_, t1, T2_, v2_, t2, t, Tbc, intr, bf, X_, obs, jac=False)[0]
        d = 1e-6
        num1 = np.zeros_like(J1); num2 = np.zeros_like(J2); nump = np.zeros_like(Jp)
        for k in range(6):
            e = np.zeros(6); e[k] = d
            num1[:, k] = (f(O.se3_mul(T1, O.se3_exp(e)), v1, T2, v2, Xw) - f(O.se3_mul(T1, O.se3_exp(-e)), v1, T2, v2, Xw)) / (2 * d)
            num2[:, k] = (f(T1, v1, O.se3_mul(T2, O.se3_exp(e)), v2, Xw) - f(T1, v1, O.se3_mul(T2, O.se3_exp(-e)), v2, Xw)) / (2 * d)
            num1[:, 6 + k] = (f(T1, v1 + e, T2, v2, Xw) - f(T1, v1 - e, T2, v2, Xw)) / (2 * d)
            num2[:, 6 + k] = (f(T1, v1, T2, v2 + e, Xw) - f(T1, v1, T2, v2 - e, Xw)) / (2 * d)
        for k in range(3):
            e = np.zeros(3); e[k] = d
            nump[:, k] = (f(T1, v1, T2, v2, Xw + e) - f(T1, v1, T2, v2, Xw - e)) / (2 * d)
        scale = max(np.abs(J1).max(), np.abs(J2).max())
        np.testing.assert_allclose(Jp, nump, atol=2e-5 * max(1, np.abs(Jp).max()))
        np.testing.assert_allclose(J1[:, 6:], num1[:, 6:], atol=2e-5 * scale)
        np.testing.assert_allclose(J2[:, 6:], num2[:, 6:], atol=2e-5 * scale)
        np.testing.assert_allclose(J1[:, :6], num1[:, :6], atol=2e-3 * scale)
        np.testing.assert_allclose(J2[:, :6], num2[:, :6], atol=2e-3 * scale)


@pytest.mark.parametrize("stereo", [False, True])
def test_edge_jacobian_formulas_vs_numpy_mirror(oracle_mod, stereo):
    """Second, independent pin of the analytic Jacobians (all blocks, including the pose blocks that a numeric derivative
    can only confirm to 1e-3 because of the reference's first-order term): the formulas of src/G2oTypes.cc:262-311 / :329-396
    composed in numpy from scipy logm/expm and the power series of the SE(3) Jacobians.  1e-10 relative to the largest entry."""
    O = oracle_mod
    rng = np.random.default_rng(40 + stereo)
    worst = 0.0
    for _ in range(12):
        qc, T1, v1, t1, T2, v2, t2, t, Tbc, intr, Xw = _edge_setup(rng, O)
        bf = 501.7
        obs = np.array([470.0, 310.0, 455.0 if stereo else -1.0])
        _, J1, J2, Jp = O.edge_eval(qc, 1, T1, v1, t1, T2, v2, t2, t, Tbc, intr, bf, Xw, obs)
        Ma, Mb, Mp = M.gp_edge_jacobians(O.se3_matrix(T1), O.se3_matrix(T2), v1, v2, t1, t2, t, O.se3_matrix(Tbc), intr, bf, Xw, stereo)
        scale = max(np.abs(Ma).max(), np.abs(Mb).max())
        for got, ref in ((J1, Ma), (J2, Mb), (Jp, Mp)):
            assert got.shape == ref.shape
            worst = max(worst, np.abs(got - ref).max() / max(scale, np.abs(ref).max()))
    assert worst < 1e-10, worst


def test_sync_edge_jacobian(oracle_mod):
    O = oracle_mod
    rng = np.random.default_rng(9)
    qc, T1, v1, t1, T2, v2, t2, t, Tbc, intr, Xw = _edge_setup(rng, O)
    Xw = O.se3_act(O.se3_mul(T2, Tbc), np.array([1.0, -0.5, 12.0]))
    obs = np.array([500.0, 280.0, -1.0])
    err, _, J2, Jp = O.edge_eval(qc, 0, T1, v1, t1, T2, v2, t2, t2, Tbc, intr, 0.0, Xw, obs)
    uv, _ = M.reproj(None, O.se3_matrix(T2), v1, v2, t1, t2, t2, O.se3_matrix(Tbc), intr, Xw, gp=False)
    np.testing.assert_allclose(err, obs[:2] - uv, atol=1e-8)
    d = 1e-6
    for k in range(6):
        e = np.zeros(6); e[k] = d
        fp = O.edge_eval(qc, 0, T1, v1, t1, O.se3_mul(T2, O.se3_exp(e)), v2, t2, t2, Tbc, intr, 0.0, Xw, obs, jac=False)[0]
        fm = O.edge_eval(qc, 0, T1, v1, t1, O.se3_mul(T2, O.se3_exp(-e)), v2, t2, t2, Tbc, intr, 0.0, Xw, obs, jac=False)[0]
        np.testing.assert_allclose(J2[:, k], (fp - fm) / (2 * d), atol=1e-4 * np.abs(J2).max())
    assert np.all(J2[:, 6:] == 0)  # EdgeMono: velocity block is zero (G2oTypes.cc:466)


def test_prior_edge(oracle_mod):
    O = oracle_mod
    rng = np.random.default_rng(5)
    qc, T1, v1, t1, T2, v2, t2, *_ = _edge_setup(rng, O)
    e, Ji, Jj = O.prior_eval(T1, v1, t1, T2, v2, t2)
    xi = M.log_se3(np.linalg.inv(O.se3_matrix(T1)) @ O.se3_matrix(T2))
    Jri = np.linalg.inv(M.Jr_series(xi))
    np.testing.assert_allclose(e, np.concatenate([xi - (t2 - t1) * v1, Jri @ v2 - v1]), atol=1e-8)
    # exact-derivative blocks (velocity columns) by central differences
    d = 1e-6
    for k in range(6):
        dv = np.zeros(6); dv[k] = d
        np.testing.assert_allclose(Ji[:, 6 + k], (O.prior_eval(T1, v1 + dv, t1, T2, v2, t2)[0] - O.prior_eval(T1, v1 - dv, t1, T2, v2, t2)[0]) / (2 * d), atol=1e-6)
        np.testing.assert_allclose(Jj[:, 6 + k], (O.prior_eval(T1, v1, t1, T2, v2 + dv, t2)[0] - O.prior_eval(T1, v1, t1, T2, v2 - dv, t2)[0]) / (2 * d), atol=1e-6)
        # top-left (xi wrt pose) is exact too
        np.testing.assert_allclose(Jj[:6, k], (O.prior_eval(T1, v1, t1, O.se3_mul(T2, O.se3_exp(dv)), v2, t2)[0][:6] - O.prior_eval(T1, v1, t1, O.se3_mul(T2, O.se3_exp(-dv)), v2, t2)[0][:6]) / (2 * d), atol=1e-5)


def test_huber_float_constants(oracle_mod):
    """SURVEY Appendix C known-answer values for the float-typed Huber parameters."""
    O = oracle_mod
    delta = float(np.float32(np.sqrt(5.991)))
    assert delta == 2.4476518630981445
    dsqr = float(np.float32(delta * delta))
    assert dsqr == 5.990999698638916
    assert list(O.huber(delta, dsqr)) == [dsqr, 1.0, 0.0]                 # e <= dsqr: inlier
    e = np.nextafter(dsqr, 10.0)
    r = O.huber(delta, e)
    assert r[0] == 2 * np.sqrt(e) * delta - dsqr and r[1] == delta / np.sqrt(e)
    d2 = float(np.float32(np.sqrt(7.815)))
    assert d2 == 2.7955322265625 and float(np.float32(d2 * d2)) == 7.815000534057617
    assert float(np.float32(21.026 * 21.026)) == 442.0926818847656
    assert float(np.float32(5.991)) == 5.991000175476074 and float(np.float32(1.5) * np.float32(5.991)) == 8.986499786376953


def test_ldlt_dense(oracle_mod):
    O = oracle_mod
    rng = np.random.default_rng(6)
    for n in (5, 36, 108):
        A = rng.normal(size=(n, n)); A = A @ A.T + 0.1 * np.eye(n)
        b = rng.normal(size=n)
        ok, x = O.ldlt_dense(A, b)
        assert ok
        np.testing.assert_allclose(x, np.linalg.solve(A, b), rtol=1e-8, atol=1e-10)
    A[3, 3] = -50.0  # indefinite -> isPositive() false -> solve() returns false (linear_solver_dense.h:108-112)
    ok, _ = O.ldlt_dense(A, b)
    assert not ok
