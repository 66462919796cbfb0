"""Golden vectors (tests/golden/*.npz, minted by tests/golden/make_golden.py from the CPU oracle).

CPU: the oracle still reproduces the committed numbers (it cannot drift silently).
GPU: the CUDA path, through the C ABI, reproduces the committed numbers -- bit-exact pattern and flags, cost within
1e-6 relative, poses within 1e-6 m / 1e-7 rad (tolerances of BASELINE.json's north_star).
The reference itself holds no golden vectors for this path and cannot be run as a whole; these fixtures
freeze the oracle, and the oracle is pinned by tests/test_ref_pin.py (the reference's own edge code, LM controller and Sim3,
compiled into oracle/_ref) and tests/test_oracle_*.py (numpy / scipy mirrors for the block-solver and Lie-group layers).
"""
import importlib.util
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "golden", "make_golden.py"))
mg = importlib.util.module_from_spec(spec)
spec.loader.exec_module(mg)

CASES = sorted(mg.CASES)


def load(key):
    return np.load(os.path.join(HERE, "golden", key + ".npz"))


def check_trace(tr, G):
    assert tr["n_iters"] == int(G["n_iters"]) and tr["result"] == int(G["result"])
    assert tr["trials"] == list(G["trials"])
    np.testing.assert_allclose(tr["chi2_before"], G["chi2_before"], rtol=1e-6)
    np.testing.assert_allclose(tr["chi2_after"], G["chi2_after"], rtol=1e-6)
    np.testing.assert_allclose(tr["lam"], G["lam"], rtol=1e-5)


def check_state(state, G):
    kp, kv, pt = state
    assert np.abs(kp[:, 4:] - G["kf_pose"][:, 4:]).max() <= 1e-6          # metres
    qa, qb = kp[:, :4], G["kf_pose"][:, :4]
    s = np.sign(np.sum(qa * qb, axis=1))[:, None]
    assert (2 * np.arcsin(np.minimum(1.0, np.linalg.norm(qa * s - qb, axis=1) / 2))).max() <= 1e-7   # radians
    assert np.abs(kv - G["kf_vel"]).max() <= 1e-5
    assert np.abs(pt - G["pt_xyz"]).max() <= 1e-5


@pytest.mark.parametrize("key", CASES)
def test_inputs_regenerate_bit_exactly(key):
    assert mg.input_checksum(mg.make_case(mg.CASES[key])) == str(load(key)["input_sha256"])


@pytest.mark.parametrize("key", CASES)
def test_oracle_reproduces_golden(oracle_mod, key):
    G = load(key)
    out = mg.run_oracle(mg.make_case(mg.CASES[key]))
    for f in ("sizes", "hpp_rows", "hpp_cols", "hs_rows", "hs_cols", "trials", "flags"):
        assert np.array_equal(out[f], G[f]), f
    for f in ("chi2_start", "chi2_before", "chi2_after", "lam", "kf_pose", "kf_vel", "pt_xyz", "edge_chi2"):
        np.testing.assert_allclose(out[f], G[f], rtol=1e-9, atol=1e-12, err_msg=f)


@pytest.mark.gpu
@pytest.mark.parametrize("key", CASES)
def test_cuda_path_reproduces_golden(key):
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from pygpba import lib as G_
    from pygpba.problem import Thresholds
    G = load(key)
    P = mg.make_case(mg.CASES[key])
    g = G_.GpBa(P)
    info = g.build_structure()
    assert [info.n_free_kf, info.n_active_pt, info.n_active_obs, info.n_hpl, info.n_hpp, info.n_hschur] == list(G["sizes"])
    for a, b in zip(g.hpp_pattern() + g.hschur_pattern(), (G["hpp_rows"], G["hpp_cols"], G["hs_rows"], G["hs_cols"])):
        assert np.array_equal(a, b)                                          # sparsity pattern: bit-exact
    assert abs(g.compute_errors() - float(G["chi2_start"])) <= 1e-10 * float(G["chi2_start"])
    g2 = G_.GpBa(P)
    check_trace(g2.optimize(10).summary(), G)
    check_state(g2.state(), G)
    th = Thresholds.local_gpba()
    c2 = G["edge_chi2"]
    band = (np.abs(c2 - th.chi2_mono) < 1e-6) | (np.abs(c2 - th.chi2_mono_close) < 1e-6)
    assert np.array_equal(g2.outlier_flags(th)[~band], G["flags"][~band])   # outlier flags: bit-exact outside the band
    np.testing.assert_allclose(g2.edge_chi2(), c2, rtol=1e-5, atol=1e-7)
