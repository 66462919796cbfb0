"""Mints tests/golden/ref_g2o_<case>.npz from THE REFERENCE'S OWN OPTIMISATION PATH.

oracle/_ref/libamc_ref_g2o.so is g2o's core (hyper graph, optimizable graph, sparse optimizer, BlockSolverX, sparse block
matrices, the quadratic forms of base_*_edge.hpp, robust kernels, Levenberg-Marquardt, LinearSolverDense) and AMC-SLAM's
src/G2oTypes.cc, GaussianProcess.cc, Pose3utils.cc, all compiled UNMODIFIED from /root/reference against the stand-in headers
of oracle/ref_shim/ (oracle/Makefile target _ref; oracle/ref_g2o_run.cc builds the graph the way Optimizer.cc does and calls
SparseOptimizer::optimize).  This script runs it on seeded problems and stores what it returns; tests/test_whole_path_reference.py holds
the oracle (CPU) and the CUDA path (-m gpu) to these numbers.  Needs /root/reference: build container only; the .npz travels.

    python tests/golden/make_golden_ref_g2o.py [case ...]

The reference run always uses LinearSolverDense (the sparse one, LinearSolverEigen, needs Eigen's SimplicialLDLT); the reduced
system has one solution, so this only moves rounding.
"""
import importlib.util
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "amc-slam_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "make_golden.py"))
mg = importlib.util.module_from_spec(spec)
spec.loader.exec_module(mg)

ITERS = 10
N_PT_SAMPLE, N_OBS_SAMPLE = 512, 2048


def make_case(key):
    """The four seeded problems of make_golden.py (same inputs as the oracle's own fixtures), BASELINE config C1 as it is,
    stereo edges (EdgeStereo and EdgeStereoGP), a start far from the optimum (rejected trials), inactive edges and edges
    without kernel."""
    from pygpba import synth
    if key in mg.CASES:
        return mg.make_case(mg.CASES[key])
    if key == "c1_full":
        return synth.make_problem("c1")
    if key == "loop_full":      # global BA after a loop closure at its preset size: 60 keyframes, 3 000 points, Huber on the priors
        return synth.make_problem("loop")
    if key == "loop_200":       # the C4 family cut to 200 keyframes driven twice round a 100-keyframe loop, 20 000 points
        return synth.make_problem("loop", n_kf=200, n_pt=20000, lap=100, seed=14)
    if key == "loop_500":       # half of C4's trajectory: 500 keyframes round a 250-keyframe loop, 50 000 points, 5 988 pose unknowns
        return synth.make_problem("loop", n_kf=500, n_pt=50000, lap=250, seed=15)
    if key == "stereo":
        return synth.add_stereo(synth.make_problem("c1", n_pt=500, seed=39), 0.5, gp_fraction=0.4)
    if key == "far_start":
        P = synth.make_problem("tiny", seed=51)
        rng = np.random.default_rng(51)
        P.pt_xyz = np.ascontiguousarray(P.pt_xyz + rng.normal(size=P.pt_xyz.shape) * 3.0)
        P.lambda_init = 1e-8
        return P
    if key == "levels":
        from pygpba.problem import OBS_LEVEL1, OBS_NO_KERNEL
        P = synth.make_problem("c1", n_pt=400, outliers=0.2, seed=53)
        rng = np.random.default_rng(53)
        fl = np.array(P.obs_flags, np.uint8)
        fl[rng.random(P.n_obs) < 0.15] |= OBS_LEVEL1
        fl[rng.random(P.n_obs) < 0.30] |= OBS_NO_KERNEL
        P.obs_flags = fl
        return P
    raise KeyError(key)


CASES = list(mg.CASES) + ["c1_full", "loop_full", "stereo", "far_start", "levels"]
SLOW_CASES = ["loop_200", "loop_500"]      # minutes through oracle/_ref; minted on request, tested when the file is there
# (loop_500: not minted -- over 70 minutes through the stand-in LDLT of 5 988 unknowns; stopped)


def samples(P):
    rng = np.random.default_rng(12345)
    ip = np.sort(rng.choice(P.n_pt, min(P.n_pt, N_PT_SAMPLE), replace=False))
    io = np.sort(rng.choice(P.n_obs, min(P.n_obs, N_OBS_SAMPLE), replace=False))
    return ip, io


def run_reference(P):
    import ref_py as R
    out = R.g2o_optimize(P, ITERS)
    ip, io = samples(P)
    return dict(n=np.int32(out["n"]), trials=np.array(out["trials"], np.int32), chi2_start=np.float64(out["chi2_start"]),
                chi2_stored=np.array(out["chi2_stored"]), lam=np.array(out["lam"]),
                last_trial_chi2=np.float64(out["last_trial_chi2"]), kf_pose=out["kf_pose"], kf_vel=out["kf_vel"],
                pt_xyz=out["pt_xyz"][ip], edge_chi2=out["edge_chi2"][io], sizes=out["sizes"], flags_packed=np.packbits(out["flags"]),
                edge_chi2_near=near_threshold(out["edge_chi2"]))


def near_threshold(chi2):
    """indices and values of the edges whose chi2 lies within 1e-3 of a threshold (the north star excludes 1e-6)"""
    from pygpba.problem import Thresholds
    th = Thresholds.local_gpba()
    idx = np.nonzero((np.abs(chi2 - th.chi2_mono) < 1e-3) | (np.abs(chi2 - th.chi2_mono_close) < 1e-3) | (np.abs(chi2 - th.chi2_stereo) < 1e-3))[0]
    return np.stack([idx.astype(np.float64), chi2[idx]]) if len(idx) else np.zeros((2, 0))


# BASELINE config C3's schedule (30 % outliers, Huber, four chi2 rejection rounds) on a C3 map cut to 2 000 points (29.5k
# observations; the stand-in matrices make the full 490k-observation C3 a half-hour run)
ROUNDS_CASE = dict(name="c3", n_pt=2000, seed=3)


def make_rounds_case():
    from pygpba import synth
    a = dict(ROUNDS_CASE)
    return synth.make_problem(a.pop("name"), **a)


def run_reference_rounds():
    import ref_py as R
    P = make_rounds_case()
    r = R.g2o_rejection_rounds(P, 4, ITERS)
    ip, io = samples(P)
    out = dict(flags_packed=np.packbits(r["flags"]), n_flagged=np.int64(r["flags"].sum()), kf_pose=r["kf_pose"], kf_vel=r["kf_vel"],
               pt_xyz=r["pt_xyz"][ip], edge_chi2=r["edge_chi2"][io], edge_chi2_near=near_threshold(r["edge_chi2"]),
               chi2_start=np.array(r["chi2_start"]), input_sha256=np.array(mg.input_checksum(P)))
    for i, t in enumerate(r["traces"]):
        out["round%d_trials" % i] = np.array(t["trials"], np.int32); out["round%d_chi2_stored" % i] = np.array(t["chi2_after"])
    return out


def run_reference_c2():
    """BASELINE config C2 (4 async cameras, 30 keyframes, 20k points, ~300k observations), the config the local-BA metric is
    quoted on; the same problem and sample indices as the oracle's fixture tests/golden/baseline_c2.npz.  Minutes."""
    import ref_py as R
    sys.path.insert(0, HERE)
    import make_golden_baseline as mb
    P = mb.make_case("c2")
    r = R.g2o_optimize(P, mb.CASES["c2"]["iters"])
    si_pt, si_obs = mb.sample_idx(P.n_pt, 101), mb.sample_idx(P.n_obs, 102)
    return dict(n=np.int32(r["n"]), trials=np.array(r["trials"], np.int32), chi2_start=np.float64(r["chi2_start"]),
                chi2_stored=np.array(r["chi2_stored"]), lam=np.array(r["lam"]), kf_pose=r["kf_pose"], kf_vel=r["kf_vel"],
                pt_xyz=r["pt_xyz"][si_pt], edge_chi2=r["edge_chi2"][si_obs], flags_packed=np.packbits(r["flags"]), sizes=r["sizes"],
                input_sha256=np.array(mg.input_checksum(P)))


# essential graph (Optimizer::OptimizeEssentialGraph): closed loops, with the scale fixed and free
POSE_GRAPHS = {"pg60": dict(n_kf=60, seed=1), "pg120_scale": dict(n_kf=120, seed=2, fix_scale=False, scale_drift=0.01)}
PG_ITERS = 20


def make_pose_graph(key):
    from pygpba import posegraph as pg
    return pg.make_pose_graph(**POSE_GRAPHS[key])


def run_reference_pose_graph(G):
    import ref_py as R
    sim3, tr = R.g2o_pose_graph(G, PG_ITERS)
    s = tr.summary()
    return dict(sim3=sim3, n=np.int32(s["n_iters"]), trials=np.array(s["trials"], np.int32), chi2_start=np.float64(tr.chi2_before[0]),
                chi2_stored=np.array(s["chi2_after"]), lam=np.array(s["lam"]))


# extrinsic self-calibration (LocalGPBA's two stages): the inputs of tests/test_extrinsic_gpu.py::setup
EXT_CASES = {"ext_tiny": ("tiny", {}), "ext_c1": ("c1", dict(n_pt=800))}
EXT_ITERS = 10


def make_ext_case(key):
    from pygpba import synth
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_extrinsic_oracle import perturb_tbc
    name, kw = EXT_CASES[key]
    P0 = synth.make_problem(name, **kw)
    P = synth.make_problem(name, **kw)
    P.cam_Tbc = perturb_tbc(P, 0, 0.05, 1.5, seed=5)
    free = np.zeros(P.n_cam, np.uint8); free[:P.n_cam - 1] = 1
    info3 = np.tile(np.diag([40.0, 30.0, 50.0]) + 2.0, (P.n_cam, 1, 1))
    return P, free, P0.cam_Tbc[:, :4].copy(), info3


def camera_observations(P):
    """cam_obs of LocalGPBA (:1131): GP observations per asynchronous camera"""
    n = np.zeros(P.n_cam, np.int64)
    gp = P.rec_kf1[P.obs_rec] >= 0
    np.add.at(n, P.rec_cam[P.obs_rec][gp], 1)
    return n


def run_reference_ext(key):
    import ref_py as R
    P, free, q_ini, info3 = make_ext_case(key)
    freed = (free * (camera_observations(P) >= 50)).astype(np.uint8)       # the >= 50 observations rule (:1224-1235)
    r = R.g2o_local_gpba_ext(P, freed, q_ini, info3, EXT_ITERS, EXT_ITERS)
    out = dict(freed=freed, kf_pose=r["kf_pose"], kf_vel=r["kf_vel"], pt_xyz=r["pt_xyz"], Tbc=r["Tbc"], chi2_start=np.float64(r["chi2_start"]),
               chi2_start2=np.float64(r["chi2_start2"]))
    for st in ("stage1", "stage2"):
        out[st + "_trials"] = np.array(r[st]["trials"], np.int32)
        out[st + "_chi2_stored"] = np.array(r[st]["chi2_after"]); out[st + "_lam"] = np.array(r[st]["lam"])
    return out


def run_reference_c3():
    """BASELINE config C3 as it is stated (4 cameras, 50 keyframes, ~490k observations, 30 % outliers, Huber + four chi2
    rejection rounds): same problem and sample indices as the oracle's fixture tests/golden/baseline_c3.npz.  Half an hour."""
    import ref_py as R
    sys.path.insert(0, HERE)
    import make_golden_baseline as mb
    P = mb.make_case("c3")
    r = R.g2o_rejection_rounds(P, mb.CASES["c3"]["rounds"], mb.CASES["c3"]["iters"])
    si_pt, si_obs = mb.sample_idx(P.n_pt, 101), mb.sample_idx(P.n_obs, 102)
    out = dict(flags_packed=np.packbits(r["flags"]), n_flagged=np.int64(r["flags"].sum()), kf_pose=r["kf_pose"], kf_vel=r["kf_vel"],
               pt_xyz=r["pt_xyz"][si_pt], edge_chi2=r["edge_chi2"][si_obs], chi2_start=np.array(r["chi2_start"]),
               input_sha256=np.array(mg.input_checksum(P)))
    for i, t in enumerate(r["traces"]):
        out["round%d_trials" % i] = np.array(t["trials"], np.int32); out["round%d_chi2_stored" % i] = np.array(t["chi2_after"])
        out["round%d_lam" % i] = np.array(t["lam"])
    return out


def _load(name):
    sp = importlib.util.spec_from_file_location(name, os.path.join(HERE, name + ".py"))
    m = importlib.util.module_from_spec(sp)
    sp.loader.exec_module(m)
    return m


mgp, mgv = _load("make_golden_pose"), _load("make_golden_vel")   # same seeded batches as the oracle's own fixtures


def run_reference_pose(key):
    """Optimizer::PoseGPOptimizationFromeLastFrame: real graph, solver and edges; packed like make_golden_pose.pack, except
    that chi2_after holds the chi2 of the STORED errors (see Recorder in oracle/ref_g2o_run.cc)."""
    import ref_py as R
    from pygpba.pose import make_pose_batch
    B = make_pose_batch(**mgp.CASES[key])
    return mgp.pack(B, R.g2o_pose_optimize(B))


def run_reference_vel(key):
    import ref_py as R
    from pygpba.velransac import make_vel_batch
    B = make_vel_batch(**mgv.CASES[key])
    return mgv.pack(B, R.g2o_vel_ransac(B))


if __name__ == "__main__":
    import ref_py as R
    assert R.build(), "needs /root/reference"
    if sys.argv[1:] == ["c3"]:
        r = run_reference_c3()
        np.savez_compressed(os.path.join(HERE, "ref_g2o_c3.npz"), **r)
        print("c3 flagged", int(r["n_flagged"]), "trials", [[int(t) for t in r["round%d_trials" % i]] for i in range(4)])
        sys.exit(0)
    if sys.argv[1:] == ["c2"]:
        r = run_reference_c2()
        np.savez_compressed(os.path.join(HERE, "ref_g2o_c2.npz"), **r)
        print("c2 iterations", int(r["n"]), "trials", [int(t) for t in r["trials"]], "chi2", float(r["chi2_start"]), "->", float(r["chi2_stored"][-1]))
        sys.exit(0)
    if not sys.argv[1:]:
        r = run_reference_rounds()
        np.savez_compressed(os.path.join(HERE, "ref_g2o_rounds_c3.npz"), **r)
        print("rounds flagged", int(r["n_flagged"]), "trials", [[int(t) for t in r["round%d_trials" % i]] for i in range(4)])
        for key in mgp.CASES:
            r = run_reference_pose(key)
            np.savez_compressed(os.path.join(HERE, "ref_g2o_pose_" + key + ".npz"), **r)
            print("pose", key, "inliers", r["n_inliers"], "iterations", r["n_iters"])
        for key in EXT_CASES:
            r = run_reference_ext(key)
            np.savez_compressed(os.path.join(HERE, "ref_g2o_" + key + ".npz"), **r)
            print("ext", key, "freed", r["freed"], "stage 2 trials", [int(t) for t in r["stage2_trials"]], "chi2", float(r["chi2_start2"]), "->",
                  float(r["stage2_chi2_stored"][-1]))
        for key in mgv.CASES:
            r = run_reference_vel(key)
            np.savez_compressed(os.path.join(HERE, "ref_g2o_vel_" + key + ".npz"), **r)
            print("vel", key, "best", int(r["best"]), "inliers of the winner", int(r["inliers"][int(r["best"])]))
        out = {}
        for key in POSE_GRAPHS:
            r = run_reference_pose_graph(make_pose_graph(key))
            out.update({key + "_" + k: v for k, v in r.items()})
            print(key, "iterations", int(r["n"]), "trials", [int(t) for t in r["trials"]], "chi2", float(r["chi2_start"]), "->", float(r["chi2_stored"][-1]))
        np.savez_compressed(os.path.join(HERE, "ref_g2o_posegraph.npz"), **out)
    for key in (sys.argv[1:] or CASES):
        P = make_case(key)
        out = run_reference(P)
        out["input_sha256"] = np.array(mg.input_checksum(P))
        np.savez_compressed(os.path.join(HERE, "ref_g2o_" + key + ".npz"), **out)
        print(key, "n_obs", P.n_obs, "iterations", int(out["n"]), "trials", [int(t) for t in out["trials"]], "chi2", float(out["chi2_start"]),
              "->", float(out["chi2_stored"][-1]))
