"""TEST INFRASTRUCTURE: ctypes access to oracle/_ref/libamc_ref_edges.so -- the reference's own src/Pose3utils.cc,
src/GaussianProcess.cc and src/G2oTypes.cc compiled unmodified against the stand-in headers in oracle/ref_shim/
(oracle/Makefile target _ref, entry points in oracle/ref_pin.cc).  Only tests/ and tests/golden/make_golden_ref.py use it.
The library exists only where /root/reference does (this container, not the GPU box); `available()` says which.
Conventions as oracle_py: poses [qx qy qz qw tx ty tz], matrices row-major, tangent [translation, rotation]."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "libamc_ref_edges.so")
_LM_SO = os.path.join(_HERE, "_ref", "libg2o_ref_lm.so")
_G2O_SO = os.path.join(_HERE, "_ref", "libamc_ref_g2o.so")
REFERENCE = "/root/reference"
_LIB = None


def build(force=False):
    """Compile oracle/_ref when the reference sources are present; returns the path or None."""
    if not os.path.isdir(os.path.join(REFERENCE, "src")):
        return _SO if os.path.exists(_SO) else None
    if force:
        import shutil
        shutil.rmtree(os.path.join(_HERE, "_ref", "obj"), ignore_errors=True)
        for so in (_SO, _LM_SO, _G2O_SO, os.path.join(_HERE, "_ref", "libadapter_check.so")):
            if os.path.exists(so):
                os.remove(so)
    jobs = str(min(16, os.cpu_count() or 1))
    subprocess.check_call(["make", "-C", _HERE, "-s", "-j", jobs, "_ref_core"])
    # the adapter check links libgpba.so; where that has not been built yet the three libraries above are still usable
    if os.path.exists(os.path.join(os.path.dirname(_HERE), "amc-slam_b200", "libgpba.so")):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-j", jobs, "_ref/libadapter_check.so"])
    return _SO


def available():
    return os.path.exists(_SO) or os.path.isdir(os.path.join(REFERENCE, "src"))


def lib():
    global _LIB
    if _LIB is None:
        so = build()
        if so is None:
            raise RuntimeError("oracle/_ref is not built and /root/reference is absent")
        _LIB = C.CDLL(so)
        for n in ("ref_query_pose", "ref_edge_eval", "ref_edge_ext_eval", "ref_pose_edge_eval"):
            getattr(_LIB, n).restype = C.c_int
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _f(x):
    return C.c_double(float(x))


def jac_pose3(xi, which):
    o = np.zeros((6, 6)); lib().ref_jac_pose3(_p(_d(xi)), int(which), _p(o)); return o


def jac_small(xi, which):
    """0 LeftJacobianPose3Q, 1 LeftJacobianRot3(xi[3:]), 2 LeftJacobianRot3Inv(xi[3:])."""
    o = np.zeros((3, 3)); lib().ref_jac_small(_p(_d(xi)), int(which), _p(o)); return o


def circle_dot(p3):
    o = np.zeros((4, 6)); lib().ref_circle_dot(_p(_d(p3)), _p(o)); return o


def so3_helper(w, which):
    """0 RightJacobianSO3, 1 InverseRightJacobianSO3, 2 ExpSO3, 3 Skew (src/G2oTypes.cc:514-599)."""
    o = np.zeros((3, 3)); lib().ref_so3_helper(_p(_d(w)), int(which), _p(o)); return o


def log_so3(R):
    o = np.zeros(3); lib().ref_log_so3(_p(_d(R)), _p(o)); return o


def query_pose(qc, T1, T2, v1, v2, t1, t2, t):
    o = np.zeros(7); A = np.zeros((6, 12)); P = np.zeros((6, 12)); dT = np.zeros(7); xi = np.zeros(6)
    same = lib().ref_query_pose(_p(_d(qc)), _p(_d(T1)), _p(_d(T2)), _p(_d(v1)), _p(_d(v2)), _f(t1), _f(t2), _f(t), _p(o), _p(A),
                                _p(P), _p(dT), _p(xi))
    return o, A, P, dT, xi, bool(same)


def gp_matrices(qc, dt, t1, t2):
    a = np.zeros((12, 12)); b = np.zeros((12, 12)); c = np.zeros((12, 12))
    lib().ref_gp_matrices(_p(_d(qc)), _f(dt), _f(t1), _f(t2), _p(a), _p(b), _p(c))
    return a, b, c


def edge_eval(qc, gp, T1, v1, t1, T2, v2, t2, t, Tbc, intr, bf, Xw, obs3):
    dim = 3 if obs3[2] >= 0 else 2
    err = np.zeros(3); J1 = np.zeros((dim, 12)); J2 = np.zeros((dim, 12)); Jp = np.zeros((dim, 3))
    depth = lib().ref_edge_eval(_p(_d(qc)), int(gp), _p(_d(T1)), _p(_d(v1)), _f(t1), _p(_d(T2)), _p(_d(v2)), _f(t2), _f(t),
                                _p(_d(Tbc)), _p(_d(intr)), _f(bf), _p(_d(Xw)), _p(_d(obs3)), _p(err), _p(J1), _p(J2), _p(Jp))
    return err[:dim], J1, J2, Jp, depth


def edge_ext_eval(qc, T1, v1, t1, T2, v2, t2, t, Tbc, intr, bf, Xw, obs2):
    err = np.zeros(2); J1 = np.zeros((2, 12)); J2 = np.zeros((2, 12)); Jp = np.zeros((2, 3)); Je = np.zeros((2, 6))
    depth = lib().ref_edge_ext_eval(_p(_d(qc)), _p(_d(T1)), _p(_d(v1)), _f(t1), _p(_d(T2)), _p(_d(v2)), _f(t2), _f(t),
                                    _p(_d(Tbc)), _p(_d(intr)), _f(bf), _p(_d(Xw)), _p(_d(obs2)), _p(err), _p(J1), _p(J2), _p(Jp),
                                    _p(Je))
    return err, J1, J2, Jp, Je, depth


def pose_edge_eval(qc, gp, T1, v1, t1, T2, v2, t2, t, Tbc, intr, bf, Xw, obs3):
    dim = 3 if obs3[2] >= 0 else 2
    err = np.zeros(3); J1 = np.zeros((dim, 12)); J2 = np.zeros((dim, 12))
    depth = lib().ref_pose_edge_eval(_p(_d(qc)), int(gp), _p(_d(T1)), _p(_d(v1)), _f(t1), _p(_d(T2)), _p(_d(v2)), _f(t2), _f(t),
                                     _p(_d(Tbc)), _p(_d(intr)), _f(bf), _p(_d(Xw)), _p(_d(obs3)), _p(err), _p(J1), _p(J2))
    return err[:dim], J1, J2, depth


def prior_eval(T1, v1, t1, T2, v2, t2):
    e = np.zeros(12); Ji = np.zeros((12, 12)); Jj = np.zeros((12, 12))
    lib().ref_prior_eval(_p(_d(T1)), _p(_d(v1)), _f(t1), _p(_d(T2)), _p(_d(v2)), _f(t2), _p(e), _p(Ji), _p(Jj))
    return e, Ji, Jj


def ext_prior_eval(q_ini, Tbc):
    e = np.zeros(3); J = np.zeros((3, 6)); lib().ref_ext_prior_eval(_p(_d(q_ini)), _p(_d(Tbc)), _p(e), _p(J)); return e, J


def velocity_edge_eval(T7, v):
    e = np.zeros(1); J = np.zeros((1, 12)); lib().ref_velocity_edge_eval(_p(_d(T7)), _p(_d(v)), _p(e), _p(J)); return e, J


def vel_edge_eval(Tlast, Tbc, intr, dt, vel, Xw, obs2):
    e = np.zeros(2); J = np.zeros((2, 6))
    lib().ref_vel_edge_eval(_p(_d(Tlast)), _p(_d(Tbc)), _p(_d(intr)), _f(dt), _p(_d(vel)), _p(_d(Xw)), _p(_d(obs2)), _p(e), _p(J))
    return e, J


def posevel_update(T7, v, upd12):
    To = np.zeros(7); vo = np.zeros(6); lib().ref_posevel_update(_p(_d(T7)), _p(_d(v)), _p(_d(upd12)), _p(To), _p(vo)); return To, vo


def extrinsic_update(Tbc, upd6):
    To = np.zeros(7); lib().ref_extrinsic_update(_p(_d(Tbc)), _p(_d(upd6)), _p(To)); return To


def standin_se3_exp(xi):
    o = np.zeros(7); lib().ref_standin_se3_exp(_p(_d(xi)), _p(o)); return o


def standin_se3_log(T7):
    o = np.zeros(6); lib().ref_standin_se3_log(_p(_d(T7)), _p(o)); return o


# ---- the reference's Levenberg-Marquardt controller on the oracle's level-1 steps (oracle/ref_lm_pin.cc) ---------------
_LM = None


def lm_lib():
    global _LM
    if _LM is None:
        if build() is None or not os.path.exists(_LM_SO):
            raise RuntimeError("oracle/_ref is not built and /root/reference is absent")
        import oracle_py
        oracle_py.lib()                      # libgpba_oracle.so first: libg2o_ref_lm.so resolves the oracle_* symbols from it
        _LM = C.CDLL(_LM_SO)
        _LM.ref_lm_optimize.restype = C.c_int
    return _LM


def lm_optimize(oracle, iters, lambda_init, max_trials=0):
    """g2o's OptimizationAlgorithmLevenberg::solve (compiled from the reference) inside SparseOptimizer::optimize's loop,
    on an oracle_py.Oracle instance.  Returns (trace, every chi2 the controller read)."""
    from pygpba.problem import LmTrace
    tr = LmTrace(); log = np.zeros(4096); n = C.c_int()
    lm_lib().ref_lm_optimize(oracle.h, int(iters), _f(lambda_init), int(max_trials), None, C.byref(tr), _p(log), len(log), C.byref(n))
    return tr, log[:n.value].copy()


def huber(delta, e):
    """RobustKernelHuber::robustify compiled from the reference's g2o (robust_kernel_impl.cpp:65-91)."""
    r = np.zeros(3); lib().ref_huber(_f(delta), _f(e), _p(r)); return r


# ---- g2o::Sim3 from the reference's Thirdparty/g2o/g2o/types/sim3.h (essential graph) ----------------------------------
def sim3_exp(u7):
    o = np.zeros(8); lib().ref_sim3_exp(_p(_d(u7)), _p(o)); return o


def sim3_log(S8):
    o = np.zeros(7); lib().ref_sim3_log(_p(_d(S8)), _p(o)); return o


def sim3_mul(a, b):
    o = np.zeros(8); lib().ref_sim3_mul(_p(_d(a)), _p(_d(b)), _p(o)); return o


def sim3_inv(a):
    o = np.zeros(8); lib().ref_sim3_inv(_p(_d(a)), _p(o)); return o


def sim3_edge_error(meas, Si, Sj):
    o = np.zeros(7); lib().ref_sim3_edge_error(_p(_d(meas)), _p(_d(Si)), _p(_d(Sj)), _p(o)); return o


def sim3_update(S, u7, fix_scale):
    o = np.zeros(8); lib().ref_sim3_update(_p(_d(S)), _p(_d(u7)), int(fix_scale), _p(o)); return o


# ---- the reference's whole optimisation path: real g2o + real AMC-SLAM edges (oracle/ref_g2o_run.cc) ---------------------
_G2O = None


def g2o_lib():
    global _G2O
    if _G2O is None:
        if build() is None or not os.path.exists(_G2O_SO):
            raise RuntimeError("oracle/_ref is not built and /root/reference is absent")
        _G2O = C.CDLL(_G2O_SO)
        _G2O.ref_g2o_optimize.restype = C.c_int
    return _G2O


def g2o_optimize(prob, iters=10, max_trials=0):
    """Builds the reference's g2o graph from a pygpba Problem and runs the real SparseOptimizer::optimize (BlockSolverX,
    LinearSolverDense, Levenberg-Marquardt).  Returns a dict: n (optimize's return value), trace summary, kf_pose, kf_vel,
    pt_xyz, edge_chi2 (stored errors), sizes [active vertices, active edges, pose dimension, landmark dimension], flags
    (LocalGPBA's inlier check on the final graph, float-typed thresholds)."""
    from pygpba.problem import LmTrace, Thresholds
    c = prob.to_c()
    kp = np.zeros((prob.n_kf, 7)); kv = np.zeros((prob.n_kf, 6)); pt = np.zeros((prob.n_pt, 3)); chi = np.zeros(prob.n_obs)
    tr = LmTrace(); sz = np.zeros(4, np.int64); th = Thresholds.local_gpba(); fl = np.zeros(prob.n_obs, np.uint8)
    n = g2o_lib().ref_g2o_optimize(C.byref(c), int(iters), int(max_trials), _p(kp), _p(kv), _p(pt), _p(chi), C.byref(tr), _p(sz),
                                   C.byref(th), _p(fl))
    s = tr.summary()
    return dict(n=n, trials=s["trials"], chi2_start=tr.chi2_before[0], chi2_stored=s["chi2_after"], lam=s["lam"],
                last_trial_chi2=s["last_trial_chi2"], kf_pose=kp, kf_vel=kv, pt_xyz=pt, edge_chi2=chi, sizes=sz, flags=fl)


def g2o_rejection_rounds(prob, n_rounds=4, iters=10):
    """BASELINE config C3's schedule (rounds of optimize + re-flagging, kernels off after the third) with the reference's
    real solver and edges.  Returns flags, trace summaries per round (chi2_after = chi2 of the stored errors), state, edge chi2."""
    from pygpba.problem import LmTrace, Thresholds
    c = prob.to_c()
    kp = np.zeros((prob.n_kf, 7)); kv = np.zeros((prob.n_kf, 6)); pt = np.zeros((prob.n_pt, 3)); chi = np.zeros(prob.n_obs)
    th = Thresholds.local_gpba(); fl = np.zeros(prob.n_obs, np.uint8); traces = (LmTrace * n_rounds)()
    L = g2o_lib()
    L.ref_g2o_rejection_rounds.restype = C.c_int
    L.ref_g2o_rejection_rounds(C.byref(c), int(n_rounds), int(iters), C.byref(th), _p(kp), _p(kv), _p(pt), _p(chi), _p(fl), traces)
    return dict(flags=fl, traces=[t.summary() for t in traces], chi2_start=[t.chi2_before[0] for t in traces], kf_pose=kp, kf_vel=kv,
                pt_xyz=pt, edge_chi2=chi)


def g2o_pose_graph(G, iters=20):
    """The essential-graph optimisation with the real VertexSim3Expmap / EdgeSim3 / BlockSolver_7_3 / LM of the reference
    (oracle/ref_g2o_run.cc) on a pygpba.posegraph.PoseGraph.  Returns (sim3 [n_kf][8], trace)."""
    from pygpba.problem import LmTrace
    c = G.to_c()
    out = np.zeros((G.n_kf, 8)); tr = LmTrace()
    g2o_lib().ref_g2o_pose_graph.restype = C.c_int
    g2o_lib().ref_g2o_pose_graph(C.byref(c), int(iters), _p(out), C.byref(tr))
    return out, tr


def g2o_pose_optimize(B):
    """Optimizer::PoseGPOptimizationFromeLastFrame with the reference's real graph and solver for every frame of a
    pygpba.pose.PoseBatch (oracle/ref_g2o_run.cc).  Returns a pygpba.pose.PoseResult like oracle_py.pose_optimize."""
    from pygpba.pose import PoseResult, GPBA_POSE_ROUNDS
    from pygpba.problem import LmTrace
    c = B.to_c()
    R = PoseResult(B)
    L = g2o_lib()
    ob = np.asarray(B.obs_begin)
    for f in range(B.n_frames):
        tr = (LmTrace * GPBA_POSE_ROUNDS)()
        n_in = C.c_int32()
        out = np.zeros(int(ob[f + 1] - ob[f]), np.uint8)
        L.ref_g2o_pose_optimize(C.byref(c), f, _p(R.cur_pose[f]), _p(R.cur_vel[f]), _p(R.prev_pose[f]), _p(R.prev_vel[f]), _p(out),
                                C.byref(n_in), tr)
        R.outlier[ob[f]:ob[f + 1]] = out
        R.n_inliers[f] = n_in.value
        for r in range(GPBA_POSE_ROUNDS):
            C.memmove(C.byref(R.traces[f * GPBA_POSE_ROUNDS + r]), C.byref(tr[r]), C.sizeof(LmTrace))
    return R


def g2o_vel_ransac(B):
    """Every hypothesis of a pygpba.velransac.VelBatch through the reference's real Optimizer::OptimizeVel graph and solver
    (oracle/ref_g2o_run.cc); the winner rule of Tracking::MCRansac (first hypothesis with the most inliers, Tracking.cc:1973)
    is applied here.  Returns a pygpba.velransac.VelResult like oracle_py.vel_ransac."""
    from pygpba.velransac import VelResult
    c = B.to_c()
    R = VelResult(B)
    L = g2o_lib()
    L.ref_g2o_optimize_vel.restype = C.c_int
    best, best_inl = -1, 0
    for h in range(B.n_hyp):
        R.inliers[h] = L.ref_g2o_optimize_vel(C.byref(c), h, _p(R.vel[h]), _p(R.mask[h]), C.byref(R.traces[h]))
        if R.inliers[h] > best_inl:
            best, best_inl = h, int(R.inliers[h])
    R.best.value = best
    return R


def g2o_local_gpba_ext(prob, ext_free, prior_q, prior_info, it1=10, it2=10):
    """LocalGPBA's two stages with extrinsic self-calibration through the reference's real VertexExtrinsic /
    EdgeMonoGPExtrinsic / EdgeExtrinsicPrior (oracle/ref_g2o_run.cc).  ext_free [n_cam]: cameras released in stage 2."""
    from pygpba.problem import LmTrace
    c = prob.to_c()
    kp = np.zeros((prob.n_kf, 7)); kv = np.zeros((prob.n_kf, 6)); pt = np.zeros((prob.n_pt, 3)); T = np.zeros((prob.n_cam, 7))
    t1, t2 = LmTrace(), LmTrace()
    f = np.ascontiguousarray(ext_free, np.uint8); q = _d(prior_q); w = _d(prior_info)
    L = g2o_lib()
    L.ref_g2o_local_gpba_ext.restype = C.c_int
    L.ref_g2o_local_gpba_ext(C.byref(c), _p(f), _p(q), _p(w), int(it1), int(it2), _p(kp), _p(kv), _p(pt), _p(T), C.byref(t1), C.byref(t2))
    return dict(kf_pose=kp, kf_vel=kv, pt_xyz=pt, Tbc=T, stage1=t1.summary(), stage2=t2.summary(), chi2_start=t1.chi2_before[0],
                chi2_start2=t2.chi2_before[0])


# ---- the reference-side binding adapter/g2o_gpba_solver.h against the real g2o headers (oracle/ref_adapter_check.cc) -------
_ADAPTER_SO = os.path.join(_HERE, "_ref", "libadapter_check.so")
_ADAPTER = None


def adapter_available():
    build()
    return os.path.exists(_ADAPTER_SO)


def adapter_lib():
    global _ADAPTER
    if _ADAPTER is None:
        if build() is None or not os.path.exists(_ADAPTER_SO):
            raise RuntimeError("oracle/_ref/libadapter_check.so is not built (needs /root/reference and libgpba.so)")
        import oracle_py
        oracle_py.lib()
        _ADAPTER = C.CDLL(_ADAPTER_SO)
        _ADAPTER.ref_adapter_roundtrip.restype = C.c_int
        _ADAPTER.ref_adapter_no_device.restype = C.c_int
    return _ADAPTER


def adapter_roundtrip(prob, iters=10):
    """problem -> the reference's real g2o graph -> the adapter's FlatGraph::build -> the oracle's optimize on what came out."""
    from pygpba.problem import LmTrace
    c = prob.to_c()
    kp = np.zeros((prob.n_kf, 7)); kv = np.zeros((prob.n_kf, 6)); pt = np.zeros((prob.n_pt, 3)); chi = np.zeros(prob.n_obs)
    counts = np.zeros(6, np.int64); tr = LmTrace()
    rc = adapter_lib().ref_adapter_roundtrip(C.byref(c), int(iters), _p(counts), _p(kp), _p(kv), _p(pt), _p(chi), C.byref(tr))
    return dict(rc=rc, counts=counts, kf_pose=kp, kf_vel=kv, pt_xyz=pt, edge_chi2=chi, trace=tr.summary())


def adapter_no_device(prob):
    c = prob.to_c()
    kp = np.zeros((prob.n_kf, 7))
    return adapter_lib().ref_adapter_no_device(C.byref(c), _p(kp)), kp


def adapter_optimize(prob, iters=10, device=0):
    """gpba::GpBaLevenberg inside the reference's real SparseOptimizer on a CUDA device (needs a GPU)."""
    from pygpba.problem import LmTrace
    c = prob.to_c()
    kp = np.zeros((prob.n_kf, 7)); kv = np.zeros((prob.n_kf, 6)); pt = np.zeros((prob.n_pt, 3)); chi = np.zeros(prob.n_obs)
    tr = LmTrace()
    L = adapter_lib()
    L.ref_adapter_optimize.restype = C.c_int
    n = L.ref_adapter_optimize(C.byref(c), int(iters), int(device), _p(kp), _p(kv), _p(pt), _p(chi), C.byref(tr))
    return dict(n=n, kf_pose=kp, kf_vel=kv, pt_xyz=pt, edge_chi2=chi, trace=tr.summary())


# ---- the binding's host logic on a TEST DOUBLE of the C ABI (oracle/abi_double.cc; no GPU needed, never shipped) ----------
_DOUBLE_SO = os.path.join(_HERE, "_ref", "libadapter_on_double.so")
_DOUBLE = None


def adapter_on_double(prob, iters=10, seam="A"):
    """seam "A": gpba::GpBaLevenberg, seam "B": gpba::GpBaBlockSolver under the stock g2o LM -- inside the reference's real
    SparseOptimizer, with the C ABI answered by the CPU oracle (test double).  Returns like g2o_optimize."""
    global _DOUBLE
    from pygpba.problem import LmTrace
    if _DOUBLE is None:
        if build() is None or not os.path.exists(_DOUBLE_SO):
            raise RuntimeError("oracle/_ref is not built and /root/reference is absent")
        import oracle_py
        oracle_py.lib()
        _DOUBLE = C.CDLL(_DOUBLE_SO)
        _DOUBLE.ref_adapter_optimize.restype = C.c_int
        _DOUBLE.ref_adapter_block_solver.restype = C.c_int
    c = prob.to_c()
    kp = np.zeros((prob.n_kf, 7)); kv = np.zeros((prob.n_kf, 6)); pt = np.zeros((prob.n_pt, 3)); chi = np.zeros(prob.n_obs)
    tr = LmTrace()
    fn = _DOUBLE.ref_adapter_optimize if seam == "A" else _DOUBLE.ref_adapter_block_solver
    n = fn(C.byref(c), int(iters), 0, _p(kp), _p(kv), _p(pt), _p(chi), C.byref(tr))
    return dict(n=n, kf_pose=kp, kf_vel=kv, pt_xyz=pt, edge_chi2=chi, trace=tr.summary())


def adapter_ext_on_double(prob, ext_free, prior_q, prior_info, it1=10, it2=10):
    """LocalGPBA's two stages through gpba::GpBaLevenberg on the graph with extrinsic vertices, C ABI = the test double."""
    from pygpba.problem import LmTrace
    adapter_on_double  # noqa: B018  (the loader below is shared)
    global _DOUBLE
    if _DOUBLE is None:
        import oracle_py
        build(); oracle_py.lib()
        _DOUBLE = C.CDLL(_DOUBLE_SO)
        _DOUBLE.ref_adapter_optimize.restype = C.c_int
        _DOUBLE.ref_adapter_block_solver.restype = C.c_int
    c = prob.to_c()
    kp = np.zeros((prob.n_kf, 7)); kv = np.zeros((prob.n_kf, 6)); pt = np.zeros((prob.n_pt, 3)); T = np.zeros((prob.n_cam, 7))
    t1, t2 = LmTrace(), LmTrace()
    f = np.ascontiguousarray(ext_free, np.uint8); q = _d(prior_q); w = _d(prior_info)
    _DOUBLE.ref_adapter_ext.restype = C.c_int
    n = _DOUBLE.ref_adapter_ext(C.byref(c), _p(f), _p(q), _p(w), int(it1), int(it2), 0, _p(kp), _p(kv), _p(pt), _p(T), C.byref(t1), C.byref(t2))
    return dict(n=n, kf_pose=kp, kf_vel=kv, pt_xyz=pt, Tbc=T, stage1=t1.summary(), stage2=t2.summary())
