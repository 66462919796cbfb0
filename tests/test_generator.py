"""Host logic: the seeded synthetic generator and the problem container."""
import numpy as np

from pygpba import synth
from pygpba.problem import OBS_CLOSE


def test_seeded_and_well_formed():
    A = synth.make_problem("c1")
    B = synth.make_problem("c1")
    for f in ("kf_pose", "kf_vel", "pt_xyz", "obs_u", "obs_v", "obs_inv_sigma2", "obs_rec", "obs_pt", "obs_flags", "rec_t"):
        assert np.array_equal(getattr(A, f), getattr(B, f)), f
    assert abs(A.n_obs - 20000) < 2500 and A.n_kf == 10 and A.n_cam == 3
    # float-typed inputs
    for f in ("cam_intr", "obs_u", "obs_v", "obs_inv_sigma2", "pt_xyz", "kf_vel"):
        a = getattr(A, f)
        assert np.array_equal(a, a.astype(np.float32).astype(np.float64)), f
    np.testing.assert_allclose(np.linalg.norm(A.kf_pose[:, :4], axis=1), 1.0, atol=1e-15)
    # async records interpolate strictly inside (t_{k-1}, t_k); sync records sit on the keyframe time
    gp = A.rec_kf1 >= 0
    assert np.all(A.rec_t[gp] > A.kf_time[A.rec_kf1[gp]]) and np.all(A.rec_t[gp] < A.kf_time[A.rec_kf2[gp]])
    assert np.all(A.rec_t[~gp] == A.kf_time[A.rec_kf2[~gp]]) and np.all(A.rec_cam[~gp] == A.n_cam - 1)
    # observations are grouped per point in keyframe order (g2o insertion order of Optimizer.cc:155-240)
    assert np.all(np.diff(A.obs_pt) >= 0)
    assert set(np.unique(A.obs_flags)) <= {0, OBS_CLOSE}
    assert A.kf_fixed[0] == 1 and A.kf_fixed[1:].sum() == 0


def test_modes_and_loop_closures():
    L = synth.make_problem("c1")
    G = synth.make_problem("loop")
    assert L.lambda_init == 1.0 and L.huber_prior == 0.0          # LocalGPBA  (Optimizer.cc:854, 903-910)
    assert G.lambda_init == 1e-5 and G.huber_prior == 21.026      # BundleAdjustment (Optimizer.cc:75, 128-130)
    # some points are re-observed a lap later -> off-band Hschur blocks
    span = np.zeros(G.n_pt, int)
    kf = G.rec_kf2[G.obs_rec]
    lo = np.full(G.n_pt, 10 ** 9); hi = np.zeros(G.n_pt, int)
    np.minimum.at(lo, G.obs_pt, kf); np.maximum.at(hi, G.obs_pt, kf)
    assert ((hi - lo) >= G.meta["lap"] - 12).mean() > 0.005


def test_subset_points_is_a_partition():
    P = synth.make_problem("tiny_global")
    m0 = np.arange(P.n_pt) % 2 == 0
    a, b = P.subset_points(m0), P.subset_points(~m0)
    assert a.n_obs + b.n_obs == P.n_obs and a.n_pt + b.n_pt == P.n_pt
    assert a.n_kf == P.n_kf and a.n_rec == P.n_rec
