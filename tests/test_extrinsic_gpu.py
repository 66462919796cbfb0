"""Extrinsic self-calibration on the device (LocalGPBA's second stage, src/Optimizer.cc:983-995,
1228-1240, 1419-1428) against the CPU oracle: block patterns bit-exact, Hpp / b / Hschur / x of the extended system, the LM
trace of the second stage, the calibrated extrinsics, and the >= 50 observations rule."""
import numpy as np
import pytest

from pygpba import synth
from test_extrinsic_oracle import perturb_tbc
from test_gpu_parity import assert_trace_equal, assert_state_close

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def G():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from pygpba import lib
    return lib


def setup(name, **kw):
    P0 = synth.make_problem(name, **kw)
    P = synth.make_problem(name, **kw)
    P.cam_Tbc = perturb_tbc(P, 0, 0.05, 1.5, seed=5)
    free = np.zeros(P.n_cam, np.uint8); free[:P.n_cam - 1] = 1
    info3 = np.tile(np.diag([40.0, 30.0, 50.0]) + 2.0, (P.n_cam, 1, 1))
    return P0, P, free, P0.cam_Tbc[:, :4].copy(), info3


def test_extended_system_matches_oracle(G, oracle_mod):
    P0, P, free, q_ini, info3 = setup("tiny")
    g = G.GpBa(P); o = oracle_mod.Oracle(P)
    g.set_extrinsics(free, q_ini, info3); o.set_extrinsics(free, q_ini, info3)
    ig, io = g.build_structure(), o.build_structure()
    for f in ("n_free_kf", "n_active_pt", "n_active_obs", "n_hpl", "n_hpp", "n_hschur"):
        assert getattr(ig, f) == getattr(io, f), f
    for a, b in zip(g.hpp_pattern() + g.hschur_pattern(), o.hpp_pattern() + o.hschur_pattern()):
        assert np.array_equal(a, b)                                          # patterns incl. the extrinsic rows: bit-exact
    cg, co = g.compute_errors(), o.compute_errors()
    assert abs(cg - co) <= 1e-11 * abs(co)
    g.build_system(); o.build_system()
    sc = np.abs(o.hpp()).max()
    np.testing.assert_allclose(g.hpp(), o.hpp(), rtol=1e-9, atol=1e-12 * sc)
    (bg, pg, Bg), (bo, po, Bo) = g.hpl(), o.hpl()
    assert np.array_equal(bg, bo) and np.array_equal(pg, po)
    np.testing.assert_allclose(Bg, Bo, rtol=1e-9, atol=1e-12 * sc)
    np.testing.assert_allclose(g.b(), o.b(), rtol=1e-9, atol=1e-10 * np.abs(o.b()).max())
    lam = P.lambda_init
    g.set_lambda(lam); o.set_lambda(lam)
    assert g.solve() and o.solve()
    (Hg, bsg), (Ho, bso) = g.hschur(), o.hschur()
    np.testing.assert_allclose(Hg, Ho, rtol=1e-8, atol=1e-11 * sc)
    np.testing.assert_allclose(bsg, bso, rtol=1e-8, atol=1e-10 * np.abs(bso).max())
    xo = o.x()
    np.testing.assert_allclose(g.x(), xo, rtol=1e-6, atol=1e-9 * np.abs(xo).max())


@pytest.mark.parametrize("name,kw", [("tiny", {}), ("c1", dict(n_pt=800))])
def test_two_stage_calibration_matches_oracle(G, oracle_mod, name, kw):
    P0, P, free, q_ini, info3 = setup(name, **kw)
    g = G.GpBa(P); o = oracle_mod.Oracle(P)
    assert_trace_equal(g.optimize(10), o.optimize(10))                       # first stage: extrinsics fixed
    assert np.array_equal(g.count_camera_observations(), o.count_camera_observations())
    o.set_extrinsics(free * (o.count_camera_observations() >= 50), q_ini, info3)
    tg, freed = g.calibrate_extrinsics(free, q_ini, info3, min_obs=50, iters=10)
    tc = o.optimize(10)
    assert np.array_equal(freed, free * (o.count_camera_observations() >= 50))
    assert_trace_equal(tg, tc)
    assert_state_close(g.state(), o.state())
    Tg, To = g.extrinsics(), o.extrinsics()
    assert np.abs(Tg[:, 4:] - To[:, 4:]).max() <= 1e-6 and np.abs(Tg[:, :4] - To[:, :4]).max() <= 1e-7
    moved = np.abs(Tg - P.cam_Tbc).max(axis=1) > 0
    assert np.array_equal(moved, freed.astype(bool))                         # only the un-fixed extrinsics moved
    np.testing.assert_allclose(g.edge_chi2(), o.edge_chi2(), rtol=1e-5, atol=1e-7)
    assert abs(g.active_robust_chi2() - o.active_robust_chi2()) <= 1e-6 * o.active_robust_chi2()


def test_threshold_prior_only_and_flags(G, oracle_mod):
    """A camera below the observation threshold stays fixed; outlier flags (isDepthPositive with the calibrated Tbc)."""
    P0, P, free, q_ini, info3 = setup("c1", n_pt=800)
    g = G.GpBa(P); o = oracle_mod.Oracle(P)
    g.optimize(10); o.optimize(10)
    n = int(o.count_camera_observations()[0]) + 1
    tg, freed = g.calibrate_extrinsics(free, q_ini, info3, min_obs=n, iters=4)     # camera 0 falls below the threshold
    assert freed[0] == 0
    o.set_extrinsics(freed, q_ini, info3)
    assert_trace_equal(tg, o.optimize(4))
    assert np.array_equal(g.outlier_flags(), o.outlier_flags())
    # reset_state brings the extrinsics back
    g.reset_state()
    assert np.array_equal(g.extrinsics(), P.cam_Tbc)


def test_stale_errors_are_handed_back(G, oracle_mod):
    """gpba_edge_errors = BaseEdge::_error of the last evaluated state (the stale-error quirk), which the adapter writes into
    the g2o edges instead of recomputing at the estimate."""
    P = synth.make_problem("c1", n_pt=600, outliers=0.1, seed=33)
    rng = np.random.default_rng(1)
    from pygpba.problem import OBS_LEVEL1
    P.obs_flags = (P.obs_flags | np.where(rng.uniform(size=P.n_obs) < 0.1, OBS_LEVEL1, 0)).astype(np.uint8)
    g = G.GpBa(P); o = oracle_mod.Oracle(P)
    g.optimize(10); o.optimize(10)
    eg, eo = g.edge_errors(), o.edge_errors()
    active = (P.obs_flags & OBS_LEVEL1) == 0
    assert np.isnan(eg[~active]).all()                                        # inactive edges: left alone
    np.testing.assert_allclose(eg[active], eo[active], rtol=1e-6, atol=1e-8)
    w = P.obs_inv_sigma2[active]
    np.testing.assert_allclose(w * (eg[active] ** 2).sum(1), g.edge_chi2()[active], rtol=1e-9, atol=1e-12)
    kp, kv, tb = g.evaluated_state()
    s = g.state()
    assert np.array_equal(kp, s[0]) == bool(np.array_equal(kv, s[1]))         # both equal (accepted) or both differ (rejected last trial)
