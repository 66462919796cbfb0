"""Pins the oracle restatement against the REFERENCE'S OWN CODE for the per-edge layer.

tests/golden/ref_edges.npz holds inputs and the outputs of the reference's src/Pose3utils.cc, src/GaussianProcess.cc and
src/G2oTypes.cc, compiled unmodified into oracle/_ref/libamc_ref_edges.so (oracle/Makefile target _ref; stand-in headers for
the absent Eigen / Sophus / g2o base classes in oracle/ref_shim/; generator tests/golden/make_golden_ref.py).  These tests
replay the inputs through the oracle (oracle/gp_edges.h, lie.h, pose_only.h, vel_ransac.h) and require agreement at 1e-11
relative to max(1, |value|) -- both sides evaluate the same closed forms in double, the difference is summation order.
Where oracle/_ref is present (the build container) a second test repeats the comparison live on fresh random inputs.

What this pins: GP interpolation (QueryPose, Qi / QiInv / Transition), the SE(3) Jacobian helpers, error and every Jacobian
block of EdgeMonoGP, EdgeStereoGP, EdgeMono, EdgeStereo, EdgeMonoGPExtrinsic (incl. J_ext), EdgeMonoGPOnlyPose,
EdgeMonoOnlyPose, EdgeStereoOnlyPose, EdgeGaussianPrior, EdgeExtrinsicPrior, EdgeVelocity, EdgeVelReproj, and the vertex
updates.  What it does not: Sophus exp/log/quaternion algebra and Eigen's inverse are stand-ins on the reference side
(pinned separately against scipy, tests/test_oracle_math.py), and the solver layer (g2o block solver, LM) is not compiled.
"""
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden", "ref_edges.npz")
TOL = 1e-11


def close(a, b, what, tol=TOL):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    err = np.abs(a - b).max() / max(1.0, np.abs(b).max())
    assert err <= tol, "%s: %.3e" % (what, err)


def same_pose(a, b, what, tol=TOL):
    """[q, t] with q and -q the same rotation."""
    a = np.array(a, dtype=np.float64)
    if np.dot(a[:4], b[:4]) < 0:
        a[:4] = -a[:4]
    close(a, b, what, tol)


def check_case(O, g, ref, tag):
    """g: inputs of one case; ref(name) -> the reference's output of that name for the case."""
    for w in range(5):
        close(O.jac_pose3(g["xi"], w), ref("jac_pose3_%d" % w), "%s jac_pose3 %d" % (tag, w))
    Tq, A, P = O.query_pose(g["qc"], g["T1"], g["T2"], g["v1"], g["v2"], g["t1"], g["t2"], g["t"])
    same_pose(Tq, ref("query_T"), tag + " QueryPose")
    close(A, ref("query_At1"), tag + " At1"); close(P, ref("query_Pt1"), tag + " Pt1")
    args = (g["T1"], g["v1"], g["t1"], g["T2"], g["v2"], g["t2"], g["t"], g["Tbc"], g["intr"], g["bf"], g["Xw"])
    for gp in (1, 0):
        for name, obs in (("mono", g["obs_mono"]), ("stereo", g["obs_stereo"])):
            e, J1, J2, Jp = O.edge_eval(g["qc"], gp, *args, obs)
            for k in ("edge", "pose"):
                if k == "pose" and gp and name == "stereo":
                    continue
                key = "%s_%s_%s_" % (k, "gp" if gp else "sync", name)
                close(e, ref(key + "err"), tag + " " + key + "err")
                close(J2, ref(key + "J2"), tag + " " + key + "J2")
                if gp:
                    close(J1, ref(key + "J1"), tag + " " + key + "J1")
                if k == "edge":
                    close(Jp, ref(key + "Jp"), tag + " " + key + "Jp")
    # isDepthPositive (G2oTypes.h:366-375, 436-442, 306-316): both keyframe poses for the GP edges, not the interpolated one
    z = [O.se3_act(O.se3_inv(O.se3_mul(T, g["Tbc"])), g["Xw"])[2] for T in (g["T1"], g["T2"])]
    assert ref("edge_gp_mono_depth") == ref("ext_depth") == float(z[0] > 0 and z[1] > 0), tag + " isDepthPositive (GP)"
    assert ref("edge_sync_mono_depth") == float(z[1] > 0), tag + " isDepthPositive"
    e, J1, J2, Jp = O.edge_eval(g["qc"], 1, *args, g["obs_mono"])
    close(e, ref("ext_err"), tag + " ext err"); close(J1, ref("ext_J1"), tag + " ext J1"); close(J2, ref("ext_J2"), tag + " ext J2")
    close(Jp, ref("ext_Jp"), tag + " ext Jp")
    close(O.edge_jext(g["qc"], 1, *args, g["obs_mono"]), ref("ext_Jext"), tag + " J_ext")
    e, Ji, Jj = O.prior_eval(g["T1"], g["v1"], g["t1"], g["T2"], g["v2"], g["t2"])
    close(e, ref("prior_err"), tag + " prior err"); close(Ji, ref("prior_Ji"), tag + " prior Ji"); close(Jj, ref("prior_Jj"), tag + " prior Jj")
    e, J = O.ext_prior_eval(g["q_ini"], g["Tbc"])
    close(e, ref("extprior_err"), tag + " extrinsic prior err")
    close(J, ref("extprior_J")[:, 3:], tag + " extrinsic prior J (rotation)")
    assert not ref("extprior_J")[:, :3].any(), tag + ": the translation block of EdgeExtrinsicPrior's Jacobian is zero"
    close(O.right_jacobian_so3_orb(g["w3"]), ref("so3_rj"), tag + " RightJacobianSO3")
    e, J = O.vel_edge_eval(g["Tlast"], g["Tbc"], g["intr"], g["dt_cam"], g["vel"], g["Xw_vel"], g["obs_vel"])
    close(e, ref("veledge_err"), tag + " EdgeVelReproj err"); close(J, ref("veledge_J"), tag + " EdgeVelReproj J")
    To, vo = O.posevel_update(g["T1"], g["v1"], g["upd12"])
    same_pose(To, ref("update_T"), tag + " PoseVelocity::Update pose"); close(vo, ref("update_v"), tag + " PoseVelocity::Update velocity")
    # EdgeVelocity (G2oTypes.h:496-519): error = Vel[2], Jacobian = unit row on the 9th tangent slot -- what the oracle
    # hard-codes as H(8, 8) += 1/Qc(2,2), b(8) -= Vel[2]/Qc(2,2) (pose_only.h build_system, gpba_oracle.cc velocity priors)
    assert ref("velocity_err")[0] == g["v1"][2]
    J = np.zeros((1, 12)); J[0, 8] = 1.0
    assert np.array_equal(ref("velocity_J"), J)


def test_golden_inputs_come_from_the_committed_script():
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden_ref", os.path.join(HERE, "golden", "make_golden_ref.py"))
    mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(mod)
    Z = np.load(GOLDEN)
    I = mod.make_inputs(int(Z["seed"]), len(Z["in_t"]))
    for k, v in I.items():
        assert np.array_equal(Z["in_" + k], v), k


def test_oracle_matches_reference_code_on_golden_vectors(oracle_mod):
    Z = np.load(GOLDEN)
    n = len(Z["in_t"])
    assert n >= 40
    ins = [k[3:] for k in Z.files if k.startswith("in_")]
    for c in range(n):
        g = {k: Z["in_" + k][c] for k in ins}
        check_case(oracle_mod, g, lambda name: Z["ref_" + name][c], "case %d" % c)


def test_huber_kernel_matches_reference_code(oracle_mod):
    """RobustKernelHuber from the reference's g2o, float-typed dsqr included: equal, not close."""
    Z = np.load(GOLDEN)
    mod = _golden_mod()
    assert np.array_equal(Z["huber_in"], mod.huber_inputs())
    inl = 0
    for (d, e), rho in zip(Z["huber_in"], Z["ref_huber"]):
        assert np.array_equal(oracle_mod.huber(d, e), rho), (d, e)
        inl += rho[1] == 1.0
    assert 0 < inl < len(Z["ref_huber"])


def test_sim3_matches_reference_code(oracle_mod):
    """g2o::Sim3 exp / log / product / inverse, EdgeSim3's error and VertexSim3Expmap's update from the reference's own
    Thirdparty/g2o/g2o/types/sim3.h against the oracle of the essential-graph optimisation (oracle/pose_graph.h)."""
    O = oracle_mod
    Z = np.load(GOLDEN)
    u, u2 = _golden_mod().sim3_inputs()
    assert len(u) == len(Z["ref_sim3_exp"])
    th = np.linalg.norm(u[:, :3], axis=1)
    assert (th < 1e-5).any() and (th > 1e-5).any() and (np.abs(u[:, 6]) < 1e-5).any() and (np.abs(u[:, 6]) > 1e-5).any()
    for i, (a, b) in enumerate(zip(u, u2)):
        S, S2 = O.sim3_exp(a), O.sim3_exp(b)
        close(S, Z["ref_sim3_exp"][i], "Sim3 exp %d" % i, 1e-14)
        close(O.sim3_log(Z["ref_sim3_exp"][i]), Z["ref_sim3_log"][i], "Sim3 log %d" % i)
        close(O.sim3_mul(S, S2), Z["ref_sim3_mul"][i], "Sim3 product %d" % i, 1e-14)
        close(O.sim3_inv(S), Z["ref_sim3_inv"][i], "Sim3 inverse %d" % i, 1e-14)
        meas, err = Z["ref_sim3_edge"][i][:8], Z["ref_sim3_edge"][i][8:]
        close(O.sim3_log(O.sim3_mul(O.sim3_mul(meas, S), O.sim3_inv(S2))), err, "EdgeSim3 error %d" % i)
        for fixed, key in ((0, "ref_sim3_update_free"), (1, "ref_sim3_update_fixed")):
            d = 0.1 * b
            if fixed:
                d = d.copy(); d[6] = 0.0
            close(O.sim3_mul(O.sim3_exp(d), S), Z[key][i], "VertexSim3Expmap update %d" % i, 1e-14)


def test_golden_covers_the_branches():
    """The fixture exercises both sides of the thresholds the closed forms switch on."""
    Z = np.load(GOLDEN)
    th = np.linalg.norm(Z["in_xi"][:, 3:], axis=1)
    assert (th < 1e-5).any() and (th > 1e-5).any()            # LeftJacobianPose3Q series / closed form (Pose3utils.cc:12)
    rel = np.linalg.norm(Z["ref_query_xi12"][:, 3:], axis=1)
    assert (rel == 0).any() and (rel > 2.0).any()             # LeftJacobianRot3(Inv) identity branch; large rotations
    assert (Z["in_t"] == Z["in_t1"]).any() and (Z["in_t"] == Z["in_t2"]).any()   # query at both ends of the interval
    w = np.linalg.norm(Z["in_w3"], axis=1)
    assert (w < 1e-5).any() and (w > 1e-5).any()              # RightJacobianSO3 (G2oTypes.cc:573-590)
    assert (Z["ref_edge_sync_mono_depth"] == 1).all()         # landmarks lie in front of the second keyframe's camera ...
    d = Z["ref_edge_gp_mono_depth"]
    assert (d == 1).any() and (d == 0).any()                  # ... and, after a large rotation, behind the first one's


def test_oracle_matches_reference_code_live(oracle_mod):
    """Fresh random inputs through the compiled reference sources; only where /root/reference (or a built oracle/_ref) is."""
    import ref_py as R
    if not R.available():
        pytest.skip("oracle/_ref is not built and /root/reference is absent (GPU box)")
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden_ref", os.path.join(HERE, "golden", "make_golden_ref.py"))
    mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(mod)
    I = mod.make_inputs(seed=7, n=120)
    out = mod.run_reference(I)
    for c in range(len(I["t"])):
        g = {k: I[k][c] for k in I}
        check_case(oracle_mod, g, lambda name: out[name][c], "live %d" % c)


def test_standin_lie_layer_agrees_with_the_oracle(oracle_mod):
    """The Sophus stand-in under oracle/ref_shim is not a pin; this only states how close the two restatements are."""
    import ref_py as R
    if not R.available():
        pytest.skip("oracle/_ref is not built and /root/reference is absent (GPU box)")
    rng = np.random.default_rng(3)
    for i in range(60):
        xi = np.concatenate([rng.normal(size=3) * 2, rng.normal(size=3) * [1.0, 1e-8, 1e-3][i % 3]])
        T = oracle_mod.se3_exp(xi)
        close(R.standin_se3_exp(xi), T, "exp", 1e-14)
        close(R.standin_se3_log(T), oracle_mod.se3_log(T), "log", 1e-13)


# ---- the Levenberg-Marquardt controller ----------------------------------------------------------------
# oracle/_ref/libg2o_ref_lm.so is the reference's Thirdparty/g2o/g2o/core/optimization_algorithm_levenberg.cpp (with
# optimization_algorithm_with_hessian.cpp, optimization_algorithm.cpp, solver.cpp, stuff/property.cpp ...) compiled
# unmodified; oracle/ref_lm_pin.cc gives it a g2o::Solver and a SparseOptimizer made of the oracle's level-1 steps.  The
# linear algebra is then the same code on both sides, so the traces must be EQUAL, not close.

def _golden_mod():
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden_ref", os.path.join(HERE, "golden", "make_golden_ref.py"))
    mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(mod)
    return mod


def test_oracle_lm_controller_equals_reference_controller_live(oracle_mod):
    import ref_py as R
    if not R.available():
        pytest.skip("oracle/_ref is not built and /root/reference is absent (GPU box)")
    from pygpba.problem import LmParams
    import ctypes as C
    mod = _golden_mod()
    seen_rejections = seen_trial_limit = seen_bad_stop = 0
    for case in mod.LM_CASES:
        a, b = mod.lm_case_oracles(case, 2)
        prm = LmParams(); a.L.oracle_default_lm_params(C.byref(prm)); prm.max_trials_after_failure = case[3]
        sa = a.optimize(mod.LM_ITERS, prm).summary()
        tb, log = R.lm_optimize(b, mod.LM_ITERS, case[1], case[3])
        sb = tb.summary()
        assert sa == sb, (case, sa, sb)                      # iterations, trials, chi2 before / after, lambda, result: bit for bit
        for x, y in zip(a.state(), b.state()):
            assert np.array_equal(x, y), case
        assert len(log) == sa["n_iters"] + sa["total_trials"]   # one chi2 per linearisation and one per trial
        seen_rejections += max(sa["trials"]) > 2
        seen_trial_limit += sa["result"] == 2 and sa["trials"][-1] == case[3]
        seen_bad_stop += sa["result"] == 2 and sa["trials"][-1] < case[3]
    assert seen_rejections >= 2 and seen_trial_limit >= 2 and seen_bad_stop >= 2   # the cases reach every exit of solve()


def test_oracle_lm_controller_matches_reference_traces_golden(oracle_mod):
    """The same comparison against the committed traces (runs everywhere).  Equal iteration / trial counts and result; the
    chi2 and lambda values within 1e-9 relative: the stored numbers went through this container's build of the oracle."""
    from pygpba.problem import LmParams
    import ctypes as C
    mod = _golden_mod()
    Z = np.load(os.path.join(HERE, "golden", "ref_lm_traces.npz"))
    for i, case in enumerate(mod.LM_CASES):
        (a,) = mod.lm_case_oracles(case, 1)
        prm = LmParams(); a.L.oracle_default_lm_params(C.byref(prm)); prm.max_trials_after_failure = case[3]
        s = a.optimize(mod.LM_ITERS, prm).summary()
        assert s["trials"] == list(Z["lm%d_trials" % i]) and s["result"] == int(Z["lm%d_result" % i]), case
        for k in ("chi2_before", "chi2_after", "lam"):
            np.testing.assert_allclose(s[k], Z["lm%d_%s" % (i, k)], rtol=1e-9, err_msg=str(case))
        np.testing.assert_allclose(s["last_trial_chi2"], float(Z["lm%d_last_trial_chi2" % i]), rtol=1e-9)
        np.testing.assert_allclose(a.state()[0], Z["lm%d_pose" % i], atol=1e-9)
