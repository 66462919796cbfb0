"""Map mirror + graph flattening (include/gpba_map.h, csrc/gpba_map.cc).

CPU: the slot-addressed SoA mirror produces exactly the gpba_problem arrays that the object-graph restatement of
Optimizer::LocalGPBA / BundleAdjustment graph construction (oracle/map_flatten.py, src/Optimizer.cc:718-1211, :85-315)
produces, on seeded random maps driven through the same sequence of map mutations (keyframe culling, point removal,
observation erasure and re-insertion); write-back semantics of the LocalGPBA tail (:1349-1430).
GPU: a window flattened from a geometric map runs through gpba_create / gpba_optimize exactly like the hand-built problem
(oracle parity), and the result is applied back to the mirror.
"""
import os
import re

import numpy as np
import pytest

from pygpba import mapmirror as MM
from pygpba import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ARRAYS = ("kf_pose", "kf_vel", "kf_time", "kf_fixed", "pt_xyz", "rec_kf1", "rec_kf2", "rec_cam", "rec_t", "obs_u", "obs_v",
          "obs_inv_sigma2", "obs_rec", "obs_pt", "obs_flags", "prior_kf1", "prior_kf2", "velp_kf")


@pytest.fixture(scope="module")
def flat():
    import sys
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import map_flatten
    return map_flatten


class Pair:
    """drives the product mirror and the restatement with the same calls"""

    def __init__(self, flat, cam_intr, cam_Tbc, bf, qc):
        self.a = MM.MapMirror(cam_intr, cam_Tbc, bf, qc)
        self.b = flat.RefMap(cam_intr, cam_Tbc, bf, qc)

    def __getattr__(self, name):
        fa, fb = getattr(self.a, name), getattr(self.b, name)

        def both(*args, **kw):
            fb(*args, **kw)
            return fa(*args, **kw)
        return both


def random_map(flat, seed, n_kf=40, n_pt=400, A=2, stereo=True):
    rng = np.random.default_rng(seed)
    n_cam = A + 1
    intr = np.tile([500.0, 500.0, 480.0, 300.0], (n_cam, 1))
    tbc = np.tile([0, 0, 0, 1.0, 0, 0, 0], (n_cam, 1)).astype(float)
    M = Pair(flat, intr, tbc, 501.7, [0.02] * 3 + [0.002] * 3)
    ids = np.cumsum(rng.integers(1, 4, n_kf))             # keyframe ids with gaps
    for i, kid in enumerate(ids):
        q = rng.normal(size=4); q /= np.linalg.norm(q)
        t = 0.1 * i
        M.add_keyframe(kid, ids[i - 1] if i else -1, np.concatenate([q, rng.normal(size=3)]), rng.normal(size=6), t,
                       np.concatenate([t - rng.uniform(0.01, 0.09, A), [t]]))
    pids = rng.permutation(10 * n_pt)[:n_pt]
    for pid in pids:
        M.add_point(pid, rng.normal(size=3) * 10)
    # observations arrive keyframe by keyframe (tracking order), points seen over a few consecutive keyframes
    first = rng.integers(0, n_kf, n_pt)
    span = rng.integers(1, 8, n_pt)
    for i, kid in enumerate(ids):
        seen = np.nonzero((first <= i) & (i < first + span))[0]
        for j in rng.permutation(seen):
            for c in range(n_cam):
                if rng.uniform() < 0.6:
                    ur = rng.uniform(0, 900) if (stereo and c == n_cam - 1 and rng.uniform() < 0.3) else -1.0
                    M.add_observation(kid, c, pids[j], rng.uniform(0, 960), rng.uniform(0, 600), ur, float(np.float32(1.2 ** -rng.integers(0, 8))), rng.uniform() < 0.2)
    return M, ids, pids, rng


def mutate(M, ids, pids, rng, n_ops=300):
    alive_kf = list(ids)
    for _ in range(n_ops):
        op = rng.uniform()
        if op < 0.05 and len(alive_kf) > 20:
            k = alive_kf.pop(int(rng.integers(1, len(alive_kf) - 1)))
            M.set_keyframe_bad(k)
        elif op < 0.15:
            M.set_point_bad(pids[rng.integers(0, len(pids))])
        elif op < 0.6:
            p = M.b.pts[pids[rng.integers(0, len(pids))]]
            if p.obs:
                kid = list(p.obs)[rng.integers(0, len(p.obs))]
                cam = list(p.obs[kid])[rng.integers(0, len(p.obs[kid]))]
                M.erase_observation(kid, cam, p.mnId)
        else:
            p = M.b.pts[pids[rng.integers(0, len(pids))]]
            if not p.bad:
                M.add_observation(alive_kf[rng.integers(0, len(alive_kf))], rng.integers(0, M.a.n_cam), p.mnId, rng.uniform(0, 960),
                                  rng.uniform(0, 600), -1.0, 1.0, rng.uniform() < 0.5)
        if op > 0.97:
            kid = alive_kf[rng.integers(0, len(alive_kf))]
            q = rng.normal(size=4); q /= np.linalg.norm(q)
            M.set_keyframe_state(kid, np.concatenate([q, rng.normal(size=3)]), rng.normal(size=6))
            M.set_point(pids[rng.integers(0, len(pids))], rng.normal(size=3))
    return alive_kf


def assert_same_window(W, R):
    P = W.problem
    for f in ARRAYS:
        a, b = getattr(P, f), R[f]
        assert a.shape == np.asarray(b).shape and np.array_equal(a, b), f
    if R["obs_ur"] is None:
        assert P.obs_ur is None
    else:
        assert np.array_equal(P.obs_ur, R["obs_ur"])
    assert np.array_equal(P.cam_intr, R["cam_intr"]) and np.array_equal(P.cam_Tbc, R["cam_Tbc"]) and np.array_equal(P.qc, R["qc"])
    for f in ("bf", "huber_mono", "huber_stereo", "huber_prior", "lambda_init"):
        assert getattr(P, f) == R[f], f
    for f in ("kf_id", "kf_role", "pt_id", "obs_kf", "obs_cam", "obs_pt_id", "cam_obs"):
        assert np.array_equal(getattr(W, f), R[f]), f


def test_header_symbols_are_exported():
    src = open(os.path.join(ROOT, "include", "gpba_map.h")).read()
    declared = sorted(set(re.findall(r"^\s*(?:int|int32_t|void|const char\*|const gpba_problem\*)\s+(gpba_\w+)\s*\(", src, re.M)))
    assert declared == sorted(MM.SYMBOLS)
    L = MM._lib()
    for name in declared:
        assert hasattr(L, name), name


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_local_windows_match_restatement(flat, seed):
    M, ids, pids, rng = random_map(flat, seed)
    for phase in range(2):
        alive = list(ids) if phase == 0 else mutate(M, ids, pids, rng)
        assert M.a.stats()["keyframes"] == M.b.n_alive()
        for kid in (alive[-1], alive[-2], alive[len(alive) // 2], alive[3], alive[0]):
            for large in (False, True):
                cov = list(rng.permutation(ids)[:6]) + [10 ** 6]        # includes bad and unknown keyframes
                W, R = M.a.local_window(kid, large, cov), M.b.local_window(kid, large, cov)
                assert_same_window(W, R)
                assert W.iterations == 10
                roles = W.kf_role
                assert (roles == 0).sum() <= (25 if large else 10) and (roles == 1).sum() <= 1 and 1 <= (roles == 2).sum() <= 50
                assert np.all(np.diff(W.kf_id) > 0)                     # Hessian order = ascending vertex id
                W.close()


@pytest.mark.parametrize("seed", [11, 12])
def test_covisibility_graph_matches_restatement(flat, seed):
    """gpba_map_update_connections against the restatement of MultiKeyFrame::UpdateConnections / AddConnection /
    UpdateBestCovisibles / EraseConnection, interleaved with keyframe culling, point removal and observation changes; then
    local windows that take the covisible keyframe from the mirror's own list."""
    M, ids, pids, rng = random_map(flat, seed, n_kf=30, n_pt=900)
    alive = list(ids)
    for rnd in range(4):
        for kid in rng.permutation(alive)[:12]:
            M.update_connections(kid)
        for kid in alive:
            ia, wa = M.a.covisibles(kid)
            ib, wb = M.b.covisibles(kid)
            assert list(ia) == ib and list(wa) == wb, (rnd, kid)
        assert any(len(M.a.covisibles(k)[0]) > 1 for k in alive)
        for kid in (alive[-1], alive[len(alive) // 2]):
            assert_same_window(M.a.local_window(kid, False, None), M.b.local_window(kid, False, None))
        alive = mutate(M, alive, pids, rng, n_ops=150)


@pytest.mark.parametrize("n_kf", [1, 2, 3, 4])
def test_tiny_maps_match_restatement(flat, n_kf):
    """Degenerate windows: Nd = min(KeyFramesInMap - 2, 10) <= 1, a first keyframe without predecessor that becomes the
    fixed one (Optimizer.cc:776-782), no covisible keyframe, points seen once."""
    M, ids, pids, rng = random_map(flat, 20 + n_kf, n_kf=n_kf, n_pt=30)
    for kid in ids:
        for large in (False, True):
            W, R = M.a.local_window(kid, large, []), M.b.local_window(kid, large, [])
            assert_same_window(W, R)
            assert (W.kf_role == 2).sum() >= 1
    assert_same_window(M.a.global_window(ids[0]), M.b.global_window(ids[0]))


@pytest.mark.parametrize("seed", [4, 5])
def test_global_window_matches_restatement(flat, seed):
    M, ids, pids, rng = random_map(flat, seed)
    mutate(M, ids, pids, rng)
    W, R = M.a.global_window(ids[0]), M.b.global_window(ids[0])
    assert_same_window(W, R)
    P = W.problem
    assert P.kf_fixed.sum() == 1 and P.kf_fixed[0] == 1 and P.huber_prior == 21.026 and P.lambda_init == 1e-5
    assert P.n_obs == M.a.stats()["observations"] - sum(   # async observations of keyframes without a previous keyframe have no edge
        1 for p in M.b.pts.values() for kid, cams in p.obs.items() for c in cams if c < M.a.n_cam - 1 and M.b.kfs[kid].mPrevKF is None)


def test_apply_erases_flagged_observations_and_rounds_states(flat):
    M, ids, pids, rng = random_map(flat, 7)
    W = M.a.local_window(ids[-1])
    P = W.problem
    n0 = M.a.stats()["observations"]
    flags = (rng.uniform(size=P.n_obs) < 0.1).astype(np.uint8)
    kp = P.kf_pose + 1e-3 * rng.normal(size=P.kf_pose.shape)
    px = P.pt_xyz + 1e-3
    # "FAIL LOCAL-GP BA": 2 * err < err_end => nothing applied (Optimizer.cc:1354-1358)
    applied, erased = M.a.apply(W, kp, None, px, flags, err=1.0, err_end=2.5)
    assert not applied and len(erased) == 0 and M.a.stats()["observations"] == n0
    applied, erased = M.a.apply(W, kp, None, px, flags, err=3.0, err_end=2.5)
    assert applied and np.array_equal(erased, np.nonzero(flags)[0]) and M.a.stats()["observations"] == n0 - int(flags.sum())
    W2 = M.a.local_window(ids[-1])
    P2 = W2.problem
    assert np.array_equal(W2.kf_id, W.kf_id)
    free = P.kf_fixed == 0
    q = kp[:, :4].astype(np.float32).astype(np.float64)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    assert np.array_equal(P2.kf_pose[free, :4], q[free]) and np.array_equal(P2.kf_pose[free, 4:], kp[free, 4:].astype(np.float32).astype(np.float64))
    assert np.array_equal(P2.kf_pose[~free], P.kf_pose[~free]) and np.array_equal(P2.kf_vel, P.kf_vel)   # LocalGPBA writes no velocities
    keep = {(int(k), int(c), int(p)) for k, c, p, f in zip(W.obs_kf, W.obs_cam, W.obs_pt_id, flags) if not f}
    now = {(int(k), int(c), int(p)) for k, c, p in zip(W2.obs_kf, W2.obs_cam, W2.obs_pt_id)}
    gone = {(int(k), int(c), int(p)) for k, c, p, f in zip(W.obs_kf, W.obs_cam, W.obs_pt_id, flags) if f}
    assert not (now & gone) and keep <= now | keep
    # a second apply of the same window must not erase anything twice
    applied, erased = M.a.apply(W, None, None, None, flags, err=3.0, err_end=2.5)
    assert applied and len(erased) == 0


def test_errors_are_reported(flat):
    from pygpba import lib as gl
    M, ids, pids, rng = random_map(flat, 8, n_kf=8, n_pt=20)
    with pytest.raises(gl.GpbaError):
        M.a.add_keyframe(ids[0], -1, [0, 0, 0, 1, 0, 0, 0], np.zeros(6), 0.0, np.zeros(3))      # duplicate id
    with pytest.raises(gl.GpbaError):
        M.a.add_observation(ids[0], 9, pids[0], 1, 2, -1, 1, 0)                              # camera out of range
    with pytest.raises(gl.GpbaError):
        M.a.local_window(123456)


def geometric_map(name="c1", **kw):
    """a synth problem (true geometry) replayed into the mirror through the mutation hooks"""
    P = synth.make_problem(name, **kw)
    M = MM.MapMirror(P.cam_intr, P.cam_Tbc, P.bf, P.qc)
    n_cam = P.n_cam
    cam_time = np.tile(P.kf_time[:, None], (1, n_cam))
    cam_time[P.rec_kf2, P.rec_cam] = P.rec_t
    for k in range(P.n_kf):
        M.add_keyframe(10 + 2 * k, 10 + 2 * (k - 1) if k else -1, P.kf_pose[k], P.kf_vel[k], P.kf_time[k], cam_time[k])
    for j in range(P.n_pt):
        M.add_point(1000 + j, P.pt_xyz[j])
    ur = P.obs_ur if P.obs_ur is not None else -np.ones(P.n_obs)
    for i in np.argsort(P.rec_kf2[P.obs_rec], kind="stable"):
        r = P.obs_rec[i]
        M.add_observation(10 + 2 * P.rec_kf2[r], P.rec_cam[r], 1000 + P.obs_pt[i], P.obs_u[i], P.obs_v[i], ur[i], P.obs_inv_sigma2[i], P.obs_flags[i] & 1)
    return P, M


def test_bulk_observation_load_equals_one_by_one():
    P, A = geometric_map("c1", n_kf=12, n_pt=400)
    B = MM.MapMirror(P.cam_intr, P.cam_Tbc, P.bf, P.qc)
    cam_time = np.tile(P.kf_time[:, None], (1, P.n_cam)); cam_time[P.rec_kf2, P.rec_cam] = P.rec_t
    for k in range(P.n_kf):
        B.add_keyframe(10 + 2 * k, 10 + 2 * (k - 1) if k else -1, P.kf_pose[k], P.kf_vel[k], P.kf_time[k], cam_time[k])
    for j in range(P.n_pt):
        B.add_point(1000 + j, P.pt_xyz[j])
    order = np.argsort(P.rec_kf2[P.obs_rec], kind="stable")
    r = P.obs_rec[order]
    n = B.add_observations(10 + 2 * P.rec_kf2[r], P.rec_cam[r], 1000 + P.obs_pt[order], P.obs_u[order], P.obs_v[order], None,
                           P.obs_inv_sigma2[order], P.obs_flags[order] & 1)
    assert n == P.n_obs and A.stats() == B.stats()
    Wa, Wb = A.local_window(10 + 2 * 11), B.local_window(10 + 2 * 11)
    for f in ARRAYS:
        assert np.array_equal(getattr(Wa.problem, f), getattr(Wb.problem, f)), f
    from pygpba import lib as gl
    with pytest.raises(gl.GpbaError):
        B.add_observations([10, 99999], [0, 0], [1000, 1000], [1.0, 1.0], [1.0, 1.0], None, [1.0, 1.0], None)   # unknown keyframe in entry 1


def test_geometric_local_window_has_reference_shape():
    P, M = geometric_map("c1", n_kf=16, n_pt=600)
    W = M.local_window(10 + 2 * 15)
    Q = W.problem
    assert (W.kf_role == 0).sum() == 10 and Q.kf_fixed.sum() >= 1            # 10 temporal keyframes + the one before + co-observers
    assert len(Q.velp_kf) == 10 and len(Q.prior_kf1) == 9 and Q.lambda_init == 1.0 and Q.huber_prior == 0.0
    assert np.all(Q.rec_kf1[Q.rec_cam < Q.n_cam - 1] >= 0) and np.all(Q.rec_kf1[Q.rec_cam == Q.n_cam - 1] == -1)
    t1, t2 = Q.kf_time[Q.rec_kf1[Q.rec_kf1 >= 0]], Q.kf_time[Q.rec_kf2[Q.rec_kf1 >= 0]]
    assert np.all((t1 < Q.rec_t[Q.rec_kf1 >= 0]) & (Q.rec_t[Q.rec_kf1 >= 0] < t2))


def test_global_window_round_trips_a_full_size_problem():
    """Size-independent property at C2 size (294 k observations): a problem replayed into the mirror through the mutation
    hooks and flattened again is the same problem -- same keyframes, points, the same multiset of edges with the same
    measurements, every GP edge on the (previous keyframe, keyframe, camera, capture time) record it came from."""
    P, M = geometric_map("c2")
    W = M.global_window(10)
    Q = W.problem
    assert Q.n_kf == P.n_kf and Q.n_pt == P.n_pt and Q.n_obs == P.n_obs and Q.n_rec == P.n_rec
    assert np.array_equal(Q.kf_pose, P.kf_pose) and np.array_equal(Q.kf_time, P.kf_time) and np.array_equal(Q.pt_xyz, P.pt_xyz)
    assert np.array_equal(Q.kf_fixed, P.kf_fixed)

    def edge_table(X):
        k2, k1, c = X.rec_kf2[X.obs_rec], X.rec_kf1[X.obs_rec], X.rec_cam[X.obs_rec]
        t = X.rec_t[X.obs_rec]
        tab = np.stack([k2.astype(float), k1.astype(float), c.astype(float), X.obs_pt.astype(float), t, X.obs_u, X.obs_v, X.obs_inv_sigma2,
                        (X.obs_flags & 1).astype(float)], 1)
        return tab[np.lexsort((tab[:, 2], tab[:, 0], tab[:, 3]))]
    assert np.array_equal(edge_table(Q), edge_table(P))
    assert sorted(zip(Q.prior_kf1, Q.prior_kf2)) == sorted(zip(P.prior_kf1, P.prior_kf2))
    assert sorted(Q.velp_kf) == sorted(P.velp_kf)
    assert Q.huber_mono == P.huber_mono and Q.huber_prior == 21.026 and Q.lambda_init == 1e-5   # BundleAdjustment's parameters


@pytest.mark.gpu
def test_flattened_window_optimizes_like_the_oracle(oracle_mod):
    from pygpba import lib as G
    from pygpba.problem import Thresholds
    P, M = geometric_map("c1", n_kf=16, n_pt=600, outliers=0.05, seed=21)
    for k in range(6):   # the keyframes that will be fixed have been optimised before: true state
        M.set_keyframe_state(10 + 2 * k, np.concatenate([P.truth["kf_q"][k], P.truth["kf_t"][k]]), P.truth["kf_vel"][k])
    W = M.local_window(10 + 2 * 15)
    Q = W.problem
    assert Q.kf_fixed.sum() == 6 and Q.n_obs > 5000
    g = G.GpBa(Q)
    g.build_structure()
    chi0 = g.compute_errors()
    tr = g.optimize(W.iterations)
    o = oracle_mod.Oracle(Q)
    tc = o.optimize(W.iterations)
    a, b = tr.summary(), tc.summary()
    assert a["n_iters"] == b["n_iters"] and a["trials"] == b["trials"]
    np.testing.assert_allclose(a["chi2_after"], b["chi2_after"], rtol=1e-6)
    kp, kv, pt = g.state()
    kp0, kv0, pt0 = o.state()
    assert np.abs(kp[:, 4:] - kp0[:, 4:]).max() <= 1e-6 and np.abs(pt - pt0).max() <= 1e-5
    flags = g.outlier_flags(Thresholds.local_gpba())
    assert np.array_equal(flags, o.outlier_flags(Thresholds.local_gpba()))
    n0 = M.stats()["observations"]
    applied, erased = M.apply(W, kp, None, pt, flags, err=np.float32(chi0), err_end=np.float32(a["chi2_after"][-1]))
    assert applied and len(erased) == int(flags.sum()) and M.stats()["observations"] == n0 - len(erased)
    # the next window starts from the optimised (float-rounded) state and has lost the erased observations
    W2 = M.local_window(10 + 2 * 15)
    assert W2.problem.n_obs == Q.n_obs - len(erased)
    g2 = G.GpBa(W2.problem)
    g2.build_structure()
    assert g2.compute_errors() < 0.1 * chi0     # the wrong associations are gone


def test_extrinsic_write_back_follows_the_observation_threshold():
    """LocalGPBA's tail for the calibrated extrinsics (src/Optimizer.cc:1419-1428): cameras with cam_obs >= extrin_thresh
    get MultiKeyFrame::mTbc[c] = estimate.cast<float>(), the others and the synchronous camera keep theirs; later windows are
    flattened with the new extrinsics."""
    P, M = geometric_map("c1", n_kf=12, n_pt=300, outliers=0.0, seed=23)
    W = M.local_window(10 + 2 * 11)
    before = M.extrinsics()
    assert np.array_equal(before, P.cam_Tbc)
    new = before.copy()
    new[:, 4:] += 0.0123456789
    q = new[:, :4] + np.array([0.01, -0.02, 0.005, 0.0]); new[:, :4] = q / np.linalg.norm(q, axis=1, keepdims=True)
    thresh = int(np.sort(W.cam_obs[:-1])[-1])                   # only the best-observed asynchronous camera passes
    n = M.apply_extrinsics(W, new, min_obs=thresh)
    after = M.extrinsics()
    passed = W.cam_obs[:-1] >= thresh
    assert n == int(passed.sum()) >= 1 and W.cam_obs[-1] == 0
    for c in range(P.n_cam):
        if c < P.n_cam - 1 and passed[c]:
            f = new[c].astype(np.float32).astype(np.float64)
            f[:4] /= np.linalg.norm(f[:4])
            assert np.array_equal(after[c], f) and not np.array_equal(after[c], before[c])
        else:
            assert np.array_equal(after[c], before[c])
    assert np.array_equal(M.local_window(10 + 2 * 11).problem.cam_Tbc, after)
