"""Essential-graph optimisation (src/Optimizer.cc:1434-1717: VertexSim3Expmap + EdgeSim3, numeric
Jacobians, BlockSolver_7_3 + LinearSolverEigen, LM with lambda_0 = 1e-16, optimize(20)).
CPU: the oracle's g2o::Sim3 restatement against scipy's matrix exponential / logarithm of 4 x 4 similarity matrices, the
numeric Jacobian against the analytic structure, loop closing on a seeded pose graph.
GPU: gpba_pose_graph_optimize / gpba_correct_points against the oracle."""
import numpy as np
import pytest
from scipy.linalg import expm, logm
from scipy.spatial.transform import Rotation as R

from pygpba import posegraph as PG


def sim3_matrix(S):
    M = np.eye(4)
    M[:3, :3] = S[7] * R.from_quat(S[:4]).as_matrix(); M[:3, 3] = S[4:7]
    return M


def generator(u):
    w, v, s = u[:3], u[3:6], u[6]
    G = np.zeros((4, 4))
    G[:3, :3] = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]]) + s * np.eye(3)
    G[:3, 3] = v
    return G


@pytest.mark.parametrize("u", [[0.3, -0.2, 0.5, 1.0, -2.0, 0.5, 0.2], [0.3, -0.2, 0.5, 1.0, -2.0, 0.5, 0.0],
                               [1e-7, 0, 0, 0.3, 0.1, -0.2, 0.1], [1e-7, 2e-7, 0, 0.3, 0.1, -0.2, 0.0], [2.5, 1.0, -0.5, 3, 2, 1, -0.4]])
def test_sim3_exp_log_match_the_matrix_exponential(oracle_mod, u):
    O = oracle_mod
    u = np.array(u, float)
    S = O.sim3_exp(u)
    np.testing.assert_allclose(sim3_matrix(S), expm(generator(u)), rtol=1e-9, atol=1e-9)   # g2o's small-angle R = I + W + W^2 is 2nd order
    back = O.sim3_log(S)
    np.testing.assert_allclose(back, u, rtol=1e-8, atol=1e-8)
    np.testing.assert_allclose(generator(back), np.real(logm(sim3_matrix(S))), rtol=1e-7, atol=1e-7)


def test_sim3_group_operations(oracle_mod):
    O = oracle_mod
    rng = np.random.default_rng(0)
    for _ in range(20):
        a, b = O.sim3_exp(rng.normal(size=7) * 0.5), O.sim3_exp(rng.normal(size=7) * 0.5)
        np.testing.assert_allclose(sim3_matrix(O.sim3_mul(a, b)), sim3_matrix(a) @ sim3_matrix(b), atol=1e-12)
        np.testing.assert_allclose(sim3_matrix(O.sim3_inv(a)), np.linalg.inv(sim3_matrix(a)), atol=1e-12)
        np.testing.assert_allclose(PG.sim3_mul(a, b), O.sim3_mul(a, b), atol=1e-12)       # the generator's numpy helpers
        np.testing.assert_allclose(PG.sim3_inv(a), O.sim3_inv(a), atol=1e-12)


@pytest.mark.parametrize("fix_scale", [True, False])
def test_oracle_closes_the_loop(oracle_mod, fix_scale):
    G = PG.make_pose_graph(n_kf=120, seed=3, fix_scale=fix_scale, scale_drift=0.0 if fix_scale else 0.002)
    out, tr = oracle_mod.pose_graph_optimize(G, 20)
    t = tr.summary()
    assert t["n_iters"] >= 2 and t["chi2_after"][t["n_iters"] - 1] < 1e-2 * t["chi2_before"][0]
    assert np.array_equal(out[0], G.sim3[0])                                             # the fixed keyframe
    if fix_scale:
        np.testing.assert_allclose(out[:, 7], G.sim3[:, 7], rtol=0, atol=1e-12)          # sigma is forced to 0
    else:
        assert np.abs(out[:, 7] - G.sim3[:, 7]).max() > 1e-6
    # consecutive keyframes keep (nearly) their measured relative pose: the error is spread over the loop
    rel = PG.sim3_mul(out[:-1], PG.sim3_inv(out[1:]))
    meas = G.edge_meas[:G.n_kf - 1]          # spanning-tree edges (i, i-1): S_(i-1) S_i^-1
    assert np.abs(rel[:, 4:7] - meas[:, 4:7]).max() < 0.05


def test_point_correction_restates_the_reference_formula(oracle_mod):
    rng = np.random.default_rng(1)
    G = PG.make_pose_graph(n_kf=40, seed=4)
    out, _ = oracle_mod.pose_graph_optimize(G, 20)
    xyz = rng.normal(size=(500, 3)) * 10; ref = rng.integers(0, 40, 500)
    got = oracle_mod.correct_points(xyz, ref, G.sim3, out)
    for i in range(0, 500, 37):
        Srw, Swr = sim3_matrix(G.sim3[ref[i]]), np.linalg.inv(sim3_matrix(out[ref[i]]))
        np.testing.assert_allclose(got[i], (Swr @ Srw @ np.append(xyz[i], 1))[:3], atol=1e-9)


# ------------------------------------------------------------------------------------------------ GPU
def shuffled(G, seed=1):
    perm = np.random.default_rng(seed).permutation(G.n_edge)
    return PG.PoseGraph(G.sim3, G.fixed, G.edge_i[perm], G.edge_j[perm], G.edge_meas[perm], G.fix_scale)


def test_the_reference_result_is_only_reproducible_to_a_band(oracle_mod):
    """EdgeSim3 is differentiated NUMERICALLY (delta = 1e-9, core/base_binary_edge.hpp:131-200): the Jacobians carry ~1e-7
    of rounding noise and the pose graph is ill-conditioned at lambda_0 = 1e-16, so everything after the error evaluation
    depends on the last bits of the arithmetic.  The reference inserts its edges in std::map / std::set order of POINTERS
    (src/Optimizer.cc:1510-1516): a permutation of the edges is a legitimate re-run of the reference, and it already moves
    the result (1e-7 m here; 1e-4 m and different iteration counts when the compiler fuses multiply-adds inside the finite
    differences, which is why the oracle is built with -ffp-contract=off).  That deviation is the floor under any parity
    claim for this row (see test_cuda_pose_graph_matches_oracle)."""
    G = PG.make_pose_graph(n_kf=60, seed=7)
    (o1, t1), (o2, t2) = oracle_mod.pose_graph_optimize(G, 20), oracle_mod.pose_graph_optimize(shuffled(G), 20)
    a, b = t1.summary(), t2.summary()
    assert abs(a["chi2_before"][0] - b["chi2_before"][0]) <= 1e-12 * a["chi2_before"][0]          # errors: exact
    assert abs(a["chi2_after"][0] - b["chi2_after"][0]) <= 1e-6 * a["chi2_after"][0]             # the first step: well determined
    assert 1e-10 < np.abs(o1[:, 4:7] - o2[:, 4:7]).max() < 1e-5                                  # the result: not bit-reproducible


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("n_kf,fix_scale", [(60, True), (150, True), (150, False)])
def test_cuda_pose_graph_matches_oracle(oracle_mod, n_kf, fix_scale):
    """Parity bar of this row: error evaluation exact (1e-10); everything downstream of the numerically differentiated
    Jacobians within 20 x the reference's own reproducibility band (oracle vs. oracle with its edges in another order),
    never looser than 2e-2 relative on the first step.  The device must reach the oracle's minimum, not its noise."""
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    G = PG.make_pose_graph(n_kf=n_kf, seed=7, fix_scale=fix_scale, scale_drift=0.0 if fix_scale else 0.002)
    og, tg = PG.optimize(G, 20)
    oc, tc = oracle_mod.pose_graph_optimize(G, 20)
    o2, t2 = oracle_mod.pose_graph_optimize(shuffled(G), 20)
    a, b, c = tg.summary(), tc.summary(), t2.summary()
    assert abs(a["chi2_before"][0] - b["chi2_before"][0]) <= 1e-10 * b["chi2_before"][0]
    assert abs(a["chi2_after"][0] - b["chi2_after"][0]) <= 2e-2 * b["chi2_after"][0]
    fa, fb, fc = a["chi2_after"][-1], b["chi2_after"][-1], c["chi2_after"][-1]
    assert abs(fa - fb) <= max(20 * abs(fb - fc), 1e-4 * fb), (fa, fb, fc)
    band_t = np.abs(oc[:, 4:7] - o2[:, 4:7]).max(); band_q = np.abs(oc[:, :4] - o2[:, :4]).max()
    assert np.abs(og[:, 4:7] - oc[:, 4:7]).max() <= max(1e-6, 20 * band_t)
    assert np.abs(og[:, :4] - oc[:, :4]).max() <= max(1e-7, 20 * band_q)
    assert np.array_equal(og[0], G.sim3[0])                                                     # the fixed keyframe
    if fix_scale:
        np.testing.assert_allclose(og[:, 7], G.sim3[:, 7], rtol=0, atol=1e-12)


@pytest.mark.gpu
def test_cuda_pose_graph_closes_a_thousand_keyframe_loop():
    """C4-sized essential graph (1k keyframes): no oracle at this size (its LinearSolverEigen restatement is a dense solve);
    the properties the optimisation must have instead."""
    G = PG.make_pose_graph(n_kf=1000, seed=7)
    out, tr = PG.optimize(G, 20)
    t = tr.summary()
    assert t["chi2_after"][-1] < 1e-4 * t["chi2_before"][0]
    assert all(x <= y * (1 + 1e-12) for x, y in zip(t["chi2_after"], t["chi2_before"]))         # LM never accepts an increase
    assert np.array_equal(out[0], G.sim3[0])
    rel = PG.sim3_mul(out[:-1], PG.sim3_inv(out[1:]))
    assert np.abs(rel[:, 4:7] - G.edge_meas[:G.n_kf - 1][:, 4:7]).max() < 0.05                  # the correction is spread over the loop


@pytest.mark.gpu
def test_cuda_point_correction_and_input_validation(oracle_mod):
    from pygpba.lib import GpbaError
    rng = np.random.default_rng(2)
    G = PG.make_pose_graph(n_kf=80, seed=9)
    out, _ = PG.optimize(G, 20)
    xyz = rng.normal(size=(20000, 3)) * 10; ref = rng.integers(0, 80, 20000)
    np.testing.assert_allclose(PG.correct_points(xyz, ref, G.sim3, out), oracle_mod.correct_points(xyz, ref, G.sim3, out), rtol=0, atol=1e-10)
    bad = PG.make_pose_graph(n_kf=20, seed=1)
    bad.edge_j[3] = 99
    with pytest.raises(GpbaError):
        PG.optimize(bad, 5)
    with pytest.raises(GpbaError):
        PG.correct_points(xyz[:10], np.full(10, 80), G.sim3, out)
    empty = PG.PoseGraph(G.sim3[:3], [1, 0, 0], [], [], np.zeros((0, 8)))
    o, tr = PG.optimize(empty, 5)                                                        # no edges: nothing to optimise
    assert np.array_equal(o, empty.sim3)
