"""Mints tests/golden/ref_edges.npz from the REFERENCE'S OWN edge code.

oracle/_ref/libamc_ref_edges.so is the reference's src/Pose3utils.cc, src/GaussianProcess.cc and src/G2oTypes.cc compiled
unmodified (oracle/Makefile target _ref, stand-in headers oracle/ref_shim/, entry points oracle/ref_pin.cc).  This script
runs it on seeded inputs and stores inputs and outputs; tests/test_ref_pin.py replays the inputs through the oracle
restatement.  Needs /root/reference, so it runs in the build container only; the .npz travels.

    python tests/golden/make_golden_ref.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

N_CASES = 40
SEED = 20261019


def unit_pose(rng, scale_t, scale_r):
    """A pose built without any library exp: random unit quaternion by axis-angle, translation."""
    axis = rng.normal(size=3); axis /= np.linalg.norm(axis)
    ang = rng.normal() * scale_r
    q = np.append(axis * np.sin(ang / 2), np.cos(ang / 2))
    return np.concatenate([q, rng.normal(size=3) * scale_t])


def quat_mul(a, b):
    ax, ay, az, aw = a; bx, by, bz, bw = b
    return np.array([aw * bx + ax * bw + ay * bz - az * by, aw * by + ay * bw + az * bx - ax * bz,
                     aw * bz + az * bw + ax * by - ay * bx, aw * bw - ax * bx - ay * by - az * bz])


def quat_rot(q, p):
    v = q[:3]; uv = 2 * np.cross(v, p)
    return p + q[3] * uv + np.cross(v, uv)


def compose(A, B):
    q = quat_mul(A[:4], B[:4]); q /= np.linalg.norm(q)
    return np.concatenate([q, A[4:] + quat_rot(A[:4], B[4:])])


def make_inputs(seed=SEED, n=N_CASES):
    """Inputs of every probe, one row per case.  The relative motion between the two keyframes cycles through ordinary
    (0.2), small (1e-3), below the 1e-5 series threshold of LeftJacobianPose3Q (1e-7), pure translation, and large (2.5 rad);
    the query time cycles through interior points and both ends of the interval."""
    rng = np.random.default_rng(seed)
    I = {k: [] for k in ("qc", "T1", "T2", "v1", "v2", "t1", "t2", "t", "Tbc", "intr", "bf", "Xw", "obs_mono", "obs_stereo", "xi",
                         "q_ini", "upd12", "Tlast", "dt_cam", "vel", "Xw_vel", "obs_vel", "w3")}
    rel_scales = [0.2, 1e-3, 1e-7, 0.0, 2.5]
    for c in range(n):
        rs = rel_scales[c % len(rel_scales)]
        T1 = unit_pose(rng, 3.0, 1.0)
        rel = unit_pose(rng, 0.3, rs)
        T2 = compose(T1, rel)
        v1 = np.concatenate([rng.normal(size=3), rng.normal(size=3) * 0.3]); v2 = v1 + rng.normal(size=6) * 0.1
        t1 = rng.uniform(0, 10); t2 = t1 + rng.uniform(0.05, 0.5)
        t = [t1 + rng.uniform(0.1, 0.9) * (t2 - t1), t1, t2][(c // len(rel_scales)) % 3]
        Tbc = unit_pose(rng, 0.5, 0.6)
        Tbc[4:] = np.float32(Tbc[4:])
        intr = np.array([rng.uniform(400, 600), rng.uniform(400, 600), rng.uniform(300, 340), rng.uniform(220, 260)])
        bf = rng.uniform(20, 60)
        # a landmark a few metres in front of the camera near the second keyframe; float-representable because the
        # pose-only edges take Eigen::Vector3f (include/G2oTypes.h:190,223,250)
        Xc = np.array([rng.normal(), rng.normal(), rng.uniform(3, 12)])
        Xw = np.float32(compose(compose(T2, Tbc), np.concatenate([[0, 0, 0, 1], Xc]))[4:]).astype(np.float64)
        I["qc"].append(rng.uniform(0.3, 3.0, size=6)); I["T1"].append(T1); I["T2"].append(T2); I["v1"].append(v1); I["v2"].append(v2)
        I["t1"].append(t1); I["t2"].append(t2); I["t"].append(t); I["Tbc"].append(Tbc); I["intr"].append(intr); I["bf"].append(bf)
        I["Xw"].append(Xw)
        I["obs_mono"].append(np.array([rng.uniform(0, 640), rng.uniform(0, 480), -1.0]))
        I["obs_stereo"].append(np.array([rng.uniform(0, 640), rng.uniform(0, 480), rng.uniform(0, 600)]))
        xi_scale = [1.0, 1e-7, 1e-3, 3.0][c % 4]
        I["xi"].append(np.concatenate([rng.normal(size=3), rng.normal(size=3) * xi_scale]))
        I["q_ini"].append(unit_pose(rng, 0, 0.6)[:4])
        I["upd12"].append(rng.normal(size=12) * [0.1, 1e-8, 1.0][c % 3])
        Tlast = unit_pose(rng, 3.0, 1.0)
        vel = np.concatenate([rng.normal(size=3), rng.normal(size=3) * 0.3]) * [1.0, 1e-6][c % 2]
        I["Tlast"].append(Tlast); I["dt_cam"].append(rng.uniform(-0.1, 0.1)); I["vel"].append(vel)
        Xc2 = np.array([rng.normal(), rng.normal(), rng.uniform(3, 12)])
        I["Xw_vel"].append(compose(compose(Tlast, Tbc), np.concatenate([[0, 0, 0, 1], Xc2]))[4:])
        I["obs_vel"].append(np.array([rng.uniform(0, 640), rng.uniform(0, 480)]))
        I["w3"].append(rng.normal(size=3) * [1.0, 1e-7, 2.5][c % 3])
    return {k: np.array(v) for k, v in I.items()}


def run_reference(I):
    import ref_py as R
    n = len(I["t"])
    out = {}

    def put(k, c, v):
        v = np.asarray(v, dtype=np.float64)
        out.setdefault(k, np.zeros((n,) + v.shape))[c] = v
    for c in range(n):
        g = {k: I[k][c] for k in I}
        for w in range(5):
            put("jac_pose3_%d" % w, c, R.jac_pose3(g["xi"], w))
        for w in range(3):
            put("jac_small_%d" % w, c, R.jac_small(g["xi"], w))
        put("circle_dot", c, R.circle_dot(g["Xw"]))
        put("so3_rj", c, R.so3_helper(g["w3"], 0)); put("so3_rj_inv", c, R.so3_helper(g["w3"], 1))
        Tq, A, P, dT, xi12, same = R.query_pose(g["qc"], g["T1"], g["T2"], g["v1"], g["v2"], g["t1"], g["t2"], g["t"])
        assert same, "the two QueryPose overloads of the reference disagree"
        put("query_T", c, Tq); put("query_At1", c, A); put("query_Pt1", c, P); put("query_dT", c, dT); put("query_xi12", c, xi12)
        Qi, QiInv, Phi = R.gp_matrices(g["qc"], g["t2"] - g["t1"], g["t1"], g["t2"])
        put("gp_Qi", c, Qi); put("gp_QiInv", c, QiInv); put("gp_Phi", c, Phi)
        args = (g["T1"], g["v1"], g["t1"], g["T2"], g["v2"], g["t2"], g["t"], g["Tbc"], g["intr"], g["bf"], g["Xw"])
        for gp in (1, 0):
            for name, obs in (("mono", g["obs_mono"]), ("stereo", g["obs_stereo"])):
                e, J1, J2, Jp, depth = R.edge_eval(g["qc"], gp, *args, obs)
                k = "edge_%s_%s_" % ("gp" if gp else "sync", name)
                put(k + "err", c, e); put(k + "J1", c, J1); put(k + "J2", c, J2); put(k + "Jp", c, Jp); put(k + "depth", c, depth)
                if not (gp and name == "stereo"):   # the reference has no stereo GP pose-only edge
                    e, J1, J2, depth = R.pose_edge_eval(g["qc"], gp, *args, obs)
                    k = "pose_%s_%s_" % ("gp" if gp else "sync", name)
                    put(k + "err", c, e); put(k + "J1", c, J1); put(k + "J2", c, J2)
        e, J1, J2, Jp, Je, depth = R.edge_ext_eval(g["qc"], *args, g["obs_mono"][:2])
        put("ext_err", c, e); put("ext_J1", c, J1); put("ext_J2", c, J2); put("ext_Jp", c, Jp); put("ext_Jext", c, Je)
        put("ext_depth", c, depth)
        e, Ji, Jj = R.prior_eval(g["T1"], g["v1"], g["t1"], g["T2"], g["v2"], g["t2"])
        put("prior_err", c, e); put("prior_Ji", c, Ji); put("prior_Jj", c, Jj)
        e, J = R.ext_prior_eval(g["q_ini"], g["Tbc"])
        put("extprior_err", c, e); put("extprior_J", c, J)
        e, J = R.velocity_edge_eval(g["T1"], g["v1"])
        put("velocity_err", c, e); put("velocity_J", c, J)
        e, J = R.vel_edge_eval(g["Tlast"], g["Tbc"], g["intr"], g["dt_cam"], g["vel"], g["Xw_vel"], g["obs_vel"])
        put("veledge_err", c, e); put("veledge_J", c, J)
        To, vo = R.posevel_update(g["T1"], g["v1"], g["upd12"])
        put("update_T", c, To); put("update_v", c, vo)
        put("update_Tbc", c, R.extrinsic_update(g["Tbc"], g["upd12"][:6]))
    return out


# (problem, lambda_init (0 = computeLambdaInit), noise on the start [m], maxTrialsAfterFailure, seed, outlier fraction):
# plain runs, the tau * max-diagonal start, runs with up to 8 rejected trials in one iteration, termination by the trial limit
# and by the three-bad-iterations rule
LM_CASES = [("tiny", 1.0, 0.0, 10, 0, None), ("tiny", 0.0, 0.0, 10, 0, None), ("tiny_global", 1e-5, 0.0, 10, 0, None),
            ("tiny_global", 0.0, 0.3, 10, 0, None), ("c1", 1.0, 0.0, 10, 0, None), ("loop", 1e-5, 0.0, 10, 0, None),
            ("tiny", 1e-8, 3.0, 10, 0, None), ("tiny", 1.0, 3.0, 10, 0, None), ("tiny", 1e-10, 3.0, 2, 0, None),
            ("tiny_global", 1e-10, 5.0, 2, 0, None), ("c1", 1e-8, 1.0, 10, 0, 0.3), ("tiny", 1e-10, 10.0, 3, 3, None)]
LM_ITERS = 10


def lm_case_oracles(case, n=2):
    """n oracle instances of one LM case at the same perturbed start (the perturbation is part of the case)."""
    sys.path.insert(0, os.path.join(ROOT, "amc-slam_b200"))
    import oracle_py as O
    from pygpba import synth
    name, lam, noise, max_trials, seed, outliers = case
    P = synth.make_problem(name, **({} if outliers is None else dict(outliers=outliers)))
    P.lambda_init = lam
    rng = np.random.default_rng(seed)
    out = [O.Oracle(P) for _ in range(n)]
    kp, kv, pt = out[0].state()
    pt2 = pt + rng.normal(size=pt.shape) * noise
    kp2 = kp.copy(); kp2[:, 4:] += rng.normal(size=(len(kp), 3)) * noise * 0.2
    for o in out:
        o.L.oracle_reset_state(o.h, O._p(kp2), O._p(kv), O._p(pt2))
    return out


def run_reference_lm():
    import ref_py as R
    res = {}
    for i, case in enumerate(LM_CASES):
        (o,) = lm_case_oracles(case, 1)
        tr, log = R.lm_optimize(o, LM_ITERS, case[1], case[3])
        s = tr.summary()
        res["lm%d_trials" % i] = np.array(s["trials"], np.int32)
        res["lm%d_result" % i] = np.int32(s["result"])
        for k in ("chi2_before", "chi2_after", "lam"):
            res["lm%d_%s" % (i, k)] = np.array(s[k])
        res["lm%d_last_trial_chi2" % i] = np.float64(s["last_trial_chi2"])
        res["lm%d_chi2_log" % i] = log
        res["lm%d_pose" % i] = o.state()[0]
    return res


# Huber deltas of the path (Optimizer.cc: sqrt(5.991) mono, sqrt(7.815) stereo, sqrt(21.026) priors in global BA) and chi2
# values on both sides of delta^2, including the doubles next to the FLOAT-rounded threshold (robust_kernel_impl.h:84)
HUBER_DELTAS = [np.sqrt(5.991), np.sqrt(7.815), np.sqrt(21.026), 1.0, 2.5]


def huber_inputs():
    rows = []
    for d in HUBER_DELTAS:
        f = float(np.float32(d * d))
        for e in (0.0, 0.5 * f, np.nextafter(f, 0), f, np.nextafter(f, np.inf), d * d, 1.5 * f, 10 * f, 1e4 * f):
            rows.append((d, e))
    return np.array(rows)


def sim3_inputs(seed=SEED + 1, n=60):
    """Tangents [omega, upsilon, sigma] on both sides of the 1e-5 thresholds of g2o::Sim3's exp / log (sim3.h:88, 161)."""
    rng = np.random.default_rng(seed)
    u = np.array([np.concatenate([rng.normal(size=3) * [1.0, 1e-7, 1e-3, 2.5][i % 4], rng.normal(size=3) * 3,
                                  [rng.normal() * [0.3, 1e-7, 0.0][i % 3]]]) for i in range(n)])
    u2 = np.array([np.concatenate([rng.normal(size=3), rng.normal(size=3) * 3, [rng.normal() * 0.2]]) for i in range(n)])
    return u, u2


def run_reference_sim3(u, u2):
    import ref_py as R
    out = {k: [] for k in ("exp", "log", "mul", "inv", "edge", "update_free", "update_fixed")}
    for a, b in zip(u, u2):
        S, S2 = R.sim3_exp(a), R.sim3_exp(b)
        meas = R.sim3_mul(R.sim3_exp(0.05 * b), R.sim3_mul(S2, R.sim3_inv(S)))   # a relative-pose measurement, slightly off
        out["exp"].append(S); out["log"].append(R.sim3_log(S)); out["mul"].append(R.sim3_mul(S, S2)); out["inv"].append(R.sim3_inv(S))
        out["edge"].append(np.concatenate([meas, R.sim3_edge_error(meas, S, S2)]))
        out["update_free"].append(R.sim3_update(S, 0.1 * b, 0)); out["update_fixed"].append(R.sim3_update(S, 0.1 * b, 1))
    return {"sim3_" + k: np.array(v) for k, v in out.items()}


def main():
    import ref_py as R
    assert R.build(force=True), "needs /root/reference"
    lm = run_reference_lm()
    path = os.path.join(HERE, "ref_lm_traces.npz")
    np.savez_compressed(path, **lm)
    print("wrote %s: %d LM cases, %.0f kB" % (path, len(LM_CASES), os.path.getsize(path) / 1e3))
    I = make_inputs()
    out = run_reference(I)
    path = os.path.join(HERE, "ref_edges.npz")
    hub = huber_inputs()
    out["huber"] = np.array([R.huber(d, e) for d, e in hub])
    out.update(run_reference_sim3(*sim3_inputs()))
    np.savez_compressed(path, seed=SEED, huber_in=hub, **{"in_" + k: v for k, v in I.items()},
                        **{"ref_" + k: v for k, v in out.items()})
    print("wrote %s: %d cases, %d reference arrays, %.0f kB" % (path, len(I["t"]), len(out), os.path.getsize(path) / 1e3))


if __name__ == "__main__":
    main()
