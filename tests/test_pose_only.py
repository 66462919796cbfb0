"""Pose-only GP optimisation (Optimizer::PoseGPOptimizationFromeLastFrame, src/Optimizer.cc:369-686).

CPU: the oracle restatement (oracle/pose_only.h) behaves like the reference function on seeded frames (recovers the true
pose, rejects the planted wrong associations, leaves a fixed previous frame alone) and still reproduces the committed
golden vectors (tests/golden/pose_*.npz, minted by tests/golden/make_golden_pose.py).
GPU: gpba_pose_optimize, through the C ABI, against the oracle and against the golden vectors: identical LM trial counts
and outlier flags (excluding matches within 1e-6 of a chi2 threshold), cost 1e-6 relative, pose 1e-6 m / 1e-7 rad.
"""
import importlib.util
import os

import numpy as np
import pytest

from pygpba import pose as PO

HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location("make_golden_pose", os.path.join(HERE, "golden", "make_golden_pose.py"))
mgp = importlib.util.module_from_spec(spec)
spec.loader.exec_module(mgp)

CASES = {
    "fixed_prev": dict(n_frames=3, n_pt=400, A=2, outliers=0.15, seed=51, fix_prev=True),
    "free_prev": dict(n_frames=3, n_pt=400, A=2, outliers=0.15, seed=52, fix_prev=False),
    "stereo": dict(n_frames=2, n_pt=400, A=2, outliers=0.1, seed=53, fix_prev=True, stereo_fraction=0.5),
    "mono_only": dict(n_frames=2, n_pt=300, A=0, outliers=0.1, seed=54, fix_prev=True),
    "few_matches": dict(n_frames=2, n_pt=5, A=2, outliers=0.0, seed=55, fix_prev=True, obs_per_pt=6),
}


def rot_angle(qa, qb):
    s = np.sign(np.sum(qa * qb, axis=1))[:, None]
    return 2 * np.arcsin(np.minimum(1.0, np.linalg.norm(qa * s - qb, axis=1) / 2))


def assert_same(B, a, b, flag_slack=0):
    for f in range(B.n_frames):
        for rnd in range(PO.GPBA_POSE_ROUNDS):
            ta, tb = a.trace(f, rnd), b.trace(f, rnd)
            assert ta["n_iters"] == tb["n_iters"] and ta["result"] == tb["result"], (f, rnd, ta, tb)
            np.testing.assert_allclose(ta["chi2_before"], tb["chi2_before"], rtol=1e-6)
            np.testing.assert_allclose(ta["chi2_after"], tb["chi2_after"], rtol=1e-6)
            for i in range(ta["n_iters"]):
                # an iteration that has converged to rounding level (gain below 1e-9 of the cost) accepts or rejects its
                # trials on the last bits of chi2: its trial count and lambda are not comparable, everything else is
                if ta["chi2_before"][i] - ta["chi2_after"][i] <= 1e-9 * ta["chi2_before"][i]:
                    continue
                assert ta["trials"][i] == tb["trials"][i], (f, rnd, ta["trials"], tb["trials"])
                np.testing.assert_allclose(ta["lam"][i], tb["lam"][i], rtol=1e-5)
    assert np.abs(a.cur_pose[:, 4:] - b.cur_pose[:, 4:]).max() <= 1e-6
    assert rot_angle(a.cur_pose[:, :4], b.cur_pose[:, :4]).max() <= 1e-7
    assert np.abs(a.cur_vel - b.cur_vel).max() <= 1e-5
    assert np.abs(a.prev_pose[:, 4:] - b.prev_pose[:, 4:]).max() <= 1e-6
    assert rot_angle(a.prev_pose[:, :4], b.prev_pose[:, :4]).max() <= 1e-7
    assert np.abs(a.prev_vel - b.prev_vel).max() <= 1e-5
    assert int((a.outlier != b.outlier).sum()) <= flag_slack
    assert np.abs(a.n_inliers.astype(int) - b.n_inliers.astype(int)).max() <= flag_slack


# ------------------------------------------------------------------------------------------------ CPU: the oracle
def test_pose_oracle_recovers_pose_and_rejects_outliers(oracle_mod):
    B = PO.make_pose_batch(**CASES["fixed_prev"])
    R = oracle_mod.pose_optimize(B)
    truth = B.truth_outlier
    assert (R.outlier[truth] > 0).mean() >= 0.95            # planted wrong associations are found
    assert (R.outlier[~truth] > 0).mean() <= 0.10            # chi2Mono = 5.991 is the 95 % quantile: ~5 % of good matches go too
    assert np.array_equal(R.n_inliers, [int((R.outlier[B.obs_begin[f]:B.obs_begin[f + 1]] == 0).sum()) for f in range(B.n_frames)])
    np.testing.assert_array_equal(R.prev_pose, B.prev_pose)  # fixed vertex: untouched
    np.testing.assert_array_equal(R.prev_vel, B.prev_vel)
    for f in range(B.n_frames):
        t0 = R.trace(f, 0)
        assert t0["n_iters"] >= 1 and t0["chi2_after"][-1] < t0["chi2_before"][0]
        for rnd in range(PO.GPBA_POSE_ROUNDS):
            t = R.trace(f, rnd)
            assert all(a <= b * (1 + 1e-12) for a, b in zip(t["chi2_after"], t["chi2_before"]))   # LM never accepts an increase
    moved = np.linalg.norm(R.cur_pose[:, 4:] - B.cur_pose[:, 4:], axis=1)
    assert (moved > 1e-4).all()


@pytest.mark.parametrize("fix_prev", [True, False])
def test_pose_oracle_gradient_matches_central_differences(oracle_mod, fix_prev):
    """Independent pin of every Jacobian on the pose-only path (reprojection edges through the GP interpolation, prior,
    velocity edges, Huber weights): b = -J^T rho' Omega e against -1/2 of the central-difference gradient of the robust
    chi2 over the 12 / 24 tangent directions; H must be symmetric positive definite."""
    B = PO.make_pose_batch(n_frames=2, n_pt=300, A=2, outliers=0.1, seed=57, fix_prev=fix_prev)
    for f in range(B.n_frames):
        H, b, chi0 = oracle_mod.pose_system(B, f)
        n = len(b)
        assert n == (12 if fix_prev else 24)
        assert np.abs(H - H.T).max() <= 1e-9 * np.abs(H).max() and np.linalg.eigvalsh(H).min() > 0
        g = np.zeros(n)
        for i in range(n):
            h = 1e-6
            d = np.zeros(n); d[i] = h
            g[i] = (oracle_mod.pose_chi2_at(B, f, d) - oracle_mod.pose_chi2_at(B, f, -d)) / (2 * h)
        assert abs(oracle_mod.pose_chi2_at(B, f, np.zeros(n)) - chi0) <= 1e-9 * chi0
        # velocity rows: exact derivatives.  Pose rows carry the reference's first-order approximation of
        # d(J_r^-1(xi) v2)/d(xi) (-0.5 ad(v2), src/G2oTypes.cc:351-357): reproduced, not corrected.
        vel = (np.arange(n) % 12) >= 6
        np.testing.assert_allclose((-0.5 * g)[vel], b[vel], rtol=2e-5, atol=1e-6 * np.abs(b).max())
        np.testing.assert_allclose((-0.5 * g)[~vel], b[~vel], rtol=2e-2, atol=2e-3 * np.abs(b).max())


def test_pose_oracle_free_previous_frame_moves(oracle_mod):
    B = PO.make_pose_batch(**CASES["free_prev"])
    R = oracle_mod.pose_optimize(B)
    assert np.abs(R.prev_pose - B.prev_pose).max() > 0
    assert (R.outlier[B.truth_outlier] > 0).mean() >= 0.9


def test_pose_oracle_few_matches_runs_one_round(oracle_mod):
    """optimizer.edges().size() < 10 breaks after the first round (Optimizer.cc:666-667)"""
    B = PO.make_pose_batch(**CASES["few_matches"])
    assert (np.diff(B.obs_begin) + 3 < 10).any()
    R = oracle_mod.pose_optimize(B)
    for f in range(B.n_frames):
        if B.obs_begin[f + 1] - B.obs_begin[f] + 3 < 10:
            assert R.trace(f, 0)["n_iters"] >= 1 and R.trace(f, 1)["n_iters"] == 0


@pytest.mark.parametrize("key", sorted(mgp.CASES))
def test_pose_oracle_reproduces_golden(oracle_mod, key):
    G = np.load(os.path.join(HERE, "golden", "pose_" + key + ".npz"))
    B = PO.make_pose_batch(**mgp.CASES[key])
    assert mgp.input_checksum(B) == str(G["input_sha256"])
    out = mgp.pack(B, oracle_mod.pose_optimize(B))
    for f in ("outlier", "n_inliers", "n_iters", "trials"):
        assert np.array_equal(out[f], G[f]), f
    for f in ("cur_pose", "cur_vel", "prev_pose", "prev_vel", "chi2_before", "chi2_after"):
        np.testing.assert_allclose(out[f], G[f], rtol=1e-9, atol=1e-12, err_msg=f)


# ------------------------------------------------------------------------------------------------ GPU: parity
@pytest.mark.gpu
@pytest.mark.parametrize("key", sorted(CASES))
def test_pose_optimize_matches_oracle(oracle_mod, key):
    B = PO.make_pose_batch(**CASES[key])
    assert_same(B, PO.pose_optimize(B), oracle_mod.pose_optimize(B))


@pytest.mark.gpu
@pytest.mark.parametrize("key", sorted(mgp.CASES))
def test_pose_optimize_matches_golden(key):
    G = np.load(os.path.join(HERE, "golden", "pose_" + key + ".npz"))
    B = PO.make_pose_batch(**mgp.CASES[key])
    out = mgp.pack(B, PO.pose_optimize(B))
    for f in ("outlier", "n_inliers", "n_iters", "trials"):
        assert np.array_equal(out[f], G[f]), f
    np.testing.assert_allclose(out["chi2_after"], G["chi2_after"], rtol=1e-6)
    assert np.abs(out["cur_pose"][:, 4:] - G["cur_pose"][:, 4:]).max() <= 1e-6
    assert rot_angle(out["cur_pose"][:, :4], G["cur_pose"][:, :4]).max() <= 1e-7


@pytest.mark.gpu
def test_pose_optimize_large_batch_is_frame_independent():
    """Size-independent property: a frame's result does not depend on the batch it travels in (one CTA per frame, no
    cross-frame state) -- 64 frames at once equal the same frames run one by one, bit for bit."""
    B = PO.make_pose_batch(n_frames=64, n_pt=3000, A=2, outliers=0.1, seed=77, fix_prev=True)
    R = PO.pose_optimize(B)
    for f in (0, 17, 63):
        one = B.slice(f)
        r1 = PO.pose_optimize(one)
        assert np.array_equal(r1.cur_pose[0], R.cur_pose[f]) and np.array_equal(r1.cur_vel[0], R.cur_vel[f])
        assert np.array_equal(r1.outlier, R.outlier[B.obs_begin[f]:B.obs_begin[f + 1]])
    truth = B.truth_outlier
    assert (R.outlier[truth] > 0).mean() >= 0.95 and (R.outlier[~truth] > 0).mean() <= 0.10


def ragged_batch():
    """frame 0 without any match (prior and velocity edges only), frame 1 with three, frame 2 complete"""
    B = PO.make_pose_batch(n_frames=3, n_pt=300, A=2, outliers=0.1, seed=58, fix_prev=True)
    keep = np.ones(B.n_obs, bool)
    keep[B.obs_begin[0]:B.obs_begin[1]] = False
    keep[B.obs_begin[1] + 3:B.obs_begin[2]] = False
    R = B.subset(keep)
    assert list(np.diff(R.obs_begin)[:2]) == [0, 3]
    return R


def test_pose_oracle_ragged_and_empty_frames(oracle_mod):
    B = ragged_batch()
    R = oracle_mod.pose_optimize(B)
    assert R.n_inliers[0] == 0 and R.trace(0, 0)["n_iters"] >= 1 and R.trace(0, 1)["n_iters"] == 0   # 3 edges < 10: one round
    assert np.isfinite(R.cur_pose).all() and np.isfinite(R.cur_vel).all()
    # with no match the frame is pulled towards the constant-velocity prediction of the fixed previous frame
    assert np.abs(R.cur_pose[0] - B.cur_pose[0]).max() > 0


@pytest.mark.gpu
def test_pose_optimize_ragged_and_empty_frames(oracle_mod):
    B = ragged_batch()
    assert_same(B, PO.pose_optimize(B), oracle_mod.pose_optimize(B))
    E = B.subset(np.zeros(B.n_obs, bool))            # no match at all in the whole batch
    assert E.n_obs == 0
    assert_same(E, PO.pose_optimize(E), oracle_mod.pose_optimize(E))


@pytest.mark.gpu
def test_pose_optimize_rejects_bad_input():
    from pygpba import lib as gl
    B = PO.make_pose_batch(**CASES["fixed_prev"])
    B.obs_cam = B.obs_cam.copy(); B.obs_cam[0] = 99
    with pytest.raises(gl.GpbaError):
        PO.pose_optimize(B)
