"""Pins the oracle's block solver / LM layer against plain dense numpy algebra on small maps."""
import numpy as np
import pytest

from pygpba import synth
from pygpba.problem import LmParams, SOLVER_PCG, SOLVER_DENSE_CHOL


def dense_normal_equations(O, P):
    """Assemble the full (poses + landmarks) H, b with numpy from per-edge Jacobians (edge probes only)."""
    o = O.Oracle(P)
    info = o.build_structure()
    o.compute_errors()
    npz, nl = info.n_free_kf, info.n_active_pt
    kf_h = -np.ones(P.n_kf, int); kf_h[P.kf_fixed == 0] = np.arange(npz)
    n = 12 * npz + 3 * nl
    H = np.zeros((n, n)); b = np.zeros(n)
    qc = P.qc
    QcInv = np.diag(1.0 / qc)
    hub = O.huber
    for i in range(P.n_obs):
        r = P.obs_rec[i]; k1, k2, c = P.rec_kf1[r], P.rec_kf2[r], P.rec_cam[r]
        gp = k1 >= 0
        ur = -1.0 if P.obs_ur is None else P.obs_ur[i]
        obs = np.array([P.obs_u[i], P.obs_v[i], ur])
        kk1 = k1 if gp else k2
        e, J1, J2, Jp = O.edge_eval(qc, gp, P.kf_pose[kk1], P.kf_vel[kk1], P.kf_time[kk1], P.kf_pose[k2], P.kf_vel[k2],
                                    P.kf_time[k2], P.rec_t[r], P.cam_Tbc[c], P.cam_intr[c], P.bf, P.pt_xyz[P.obs_pt[i]], obs)
        w = P.obs_inv_sigma2[i]
        chi2 = w * e @ e
        rho1 = hub(P.huber_stereo if ur >= 0 else P.huber_mono, chi2)[1]
        cols, Js = [], []
        if gp and kf_h[k1] >= 0:
            cols.append(np.arange(12) + 12 * kf_h[k1]); Js.append(J1)
        if kf_h[k2] >= 0:
            cols.append(np.arange(12) + 12 * kf_h[k2]); Js.append(J2)
        cols.append(12 * npz + 3 * P.obs_pt[i] + np.arange(3)); Js.append(Jp)
        cols = np.concatenate(cols); J = np.hstack(Js)
        H[np.ix_(cols, cols)] += rho1 * w * J.T @ J
        b[cols] += -rho1 * w * J.T @ e
    for k in P.velp_kf:
        if kf_h[k] >= 0:
            H[12 * kf_h[k] + 8, 12 * kf_h[k] + 8] += QcInv[2, 2]
            b[12 * kf_h[k] + 8] += -QcInv[2, 2] * P.kf_vel[k][2]
    for k1, k2 in zip(P.prior_kf1, P.prior_kf2):
        e, Ji, Jj = O.prior_eval(P.kf_pose[k1], P.kf_vel[k1], P.kf_time[k1], P.kf_pose[k2], P.kf_vel[k2], P.kf_time[k2])
        dt = P.kf_time[k2] - P.kf_time[k1]
        Om = np.block([[12 / dt ** 3 * QcInv, -6 / dt ** 2 * QcInv], [-6 / dt ** 2 * QcInv, 4 / dt * QcInv]])
        rho1 = hub(P.huber_prior, e @ Om @ e)[1] if P.huber_prior > 0 else 1.0
        cols, Js = [], []
        if kf_h[k1] >= 0:
            cols.append(np.arange(12) + 12 * kf_h[k1]); Js.append(Ji)
        if kf_h[k2] >= 0:
            cols.append(np.arange(12) + 12 * kf_h[k2]); Js.append(Jj)
        if not cols:
            continue
        cols = np.concatenate(cols); J = np.hstack(Js)
        H[np.ix_(cols, cols)] += rho1 * J.T @ Om @ J
        b[cols] += -rho1 * J.T @ Om @ e
    return o, info, H, b


@pytest.mark.parametrize("name", ["tiny", "tiny_global"])
def test_build_system_and_schur_vs_dense(oracle_mod, name):
    O = oracle_mod
    P = synth.make_problem(name)
    o, info, H, b = dense_normal_equations(O, P)
    o.build_system()
    npz = info.n_free_kf
    # blocks vs dense assembly
    r, c = o.hpp_pattern()
    for blk, i, j in zip(o.hpp(), r, c):
        np.testing.assert_allclose(blk, H[12 * i:12 * i + 12, 12 * j:12 * j + 12], rtol=1e-9, atol=1e-6)
    for l, blk in enumerate(o.hll()):
        s = 12 * npz + 3 * l
        np.testing.assert_allclose(blk, H[s:s + 3, s:s + 3], rtol=1e-9, atol=1e-6)
    beg, pose, blks = o.hpl()
    for l in range(info.n_active_pt):
        for s in range(beg[l], beg[l + 1]):
            np.testing.assert_allclose(blks[s], H[12 * pose[s]:12 * pose[s] + 12, 12 * npz + 3 * l:12 * npz + 3 * l + 3], rtol=1e-9, atol=1e-6)
    np.testing.assert_allclose(o.b(), b, rtol=1e-9, atol=1e-6)
    # Hpp pattern: every nonzero pose-pose block of the dense matrix is in the pattern and vice versa
    dense_pat = {(i, j) for i in range(npz) for j in range(i, npz) if np.any(H[12 * i:12 * i + 12, 12 * j:12 * j + 12] != 0)}
    assert dense_pat == set(zip(r.tolist(), c.tolist()))
    # Schur solve == dense solve of the damped full system
    lam = 0.37
    o.set_lambda(lam)
    assert o.solve()
    x = o.x()
    xd = np.linalg.solve(H + lam * np.eye(len(b)), b)
    np.testing.assert_allclose(x, xd, rtol=1e-7, atol=1e-9)
    # Hschur pattern == structural fill of Hpp - Hpl Hll^-1 Hlp
    Hs, bs = o.hschur()
    rs, cs = o.hschur_pattern()
    Hd = H + lam * np.eye(len(b))
    S = Hd[:12 * npz, :12 * npz] - Hd[:12 * npz, 12 * npz:] @ np.linalg.solve(Hd[12 * npz:, 12 * npz:], Hd[12 * npz:, :12 * npz])
    for blk, i, j in zip(Hs, rs, cs):
        np.testing.assert_allclose(blk, S[12 * i:12 * i + 12, 12 * j:12 * j + 12], rtol=1e-8, atol=1e-6)
    struct = {(i, j) for i in range(npz) for j in range(i, npz) if np.abs(S[12 * i:12 * i + 12, 12 * j:12 * j + 12]).max() > 1e-9}
    assert struct <= set(zip(rs.tolist(), cs.tolist()))
    o.restore_diagonal()
    np.testing.assert_allclose(o.hll()[0], H[12 * npz:12 * npz + 3, 12 * npz:12 * npz + 3], rtol=1e-9, atol=1e-6)


def test_sparse_solver_matches_dense(oracle_mod):
    O = oracle_mod
    P = synth.make_problem("loop")
    xs = []
    for solver in (SOLVER_DENSE_CHOL, SOLVER_PCG):
        P.linear_solver = solver
        o = O.Oracle(P)
        o.build_structure(); o.compute_errors(); o.build_system(); o.set_lambda(1e-5)
        assert o.solve()
        xs.append(o.x())
    np.testing.assert_allclose(xs[0], xs[1], rtol=1e-6, atol=1e-9)


def test_lm_l2_equals_l1_stepping(oracle_mod):
    """optimize() == driving the Solver-shaped calls by hand with the LM rules of
    optimization_algorithm_levenberg.cpp:61-169 (checks the two API levels against each other)."""
    O = oracle_mod
    P = synth.make_problem("tiny")
    tr = O.Oracle(P).optimize(6).summary()
    o = O.Oracle(P)
    lam, ni, chis = P.lambda_init, 2.0, []
    o.build_structure()
    for it in range(6):
        cur = o.compute_errors()
        o.build_system()
        q = 0
        while True:
            o.push(); o.set_lambda(lam); ok = o.solve(); x = o.x(); b = o.b(); o.oplus(); o.restore_diagonal()
            tmp = o.compute_errors()
            rho = (cur - tmp) / (np.dot(x, lam * x + b) + 1e-3)
            if rho > 0 and np.isfinite(tmp):
                lam *= max(1 / 3, min(1 - (2 * rho - 1) ** 3, 2 / 3)); ni = 2.0; cur = tmp; o.discard_top()
            else:
                lam *= ni; ni *= 2; o.pop()
            q += 1
            if not (rho < 0 and q < 10):
                break
        chis.append(cur)
    np.testing.assert_allclose(chis, tr["chi2_after"][:6], rtol=1e-12)


def test_rejection_rounds_find_injected_outliers(oracle_mod):
    O = oracle_mod
    P = synth.make_problem("c1", n_pt=300, outliers=0.3, seed=21)
    o = O.Oracle(P)
    flags, traces = o.rejection_rounds(4, 10)
    truth = P.truth["is_outlier"]
    # the chi2 test must recover the injected outliers almost perfectly
    assert (flags[truth] == 1).mean() > 0.97
    assert (flags[~truth] == 0).mean() > 0.85  # chi2(2) 95% quantile: >=5% of true inliers are flagged by design
    assert all(t.n_iters >= 1 for t in traces)
