"""ctypes binding of oracle/libgpba_oracle.so -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) may import this.
PARITY: pinned per layer against the reference's own sources compiled into oracle/_ref (ref_py.py).
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(_HERE), "amc-slam_b200"))
from pygpba.problem import CProblem, LmParams, LmTrace, StructureInfo, Thresholds  # noqa: E402

_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "libgpba_oracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("gpba_oracle.cc", "gp_edges.h", "lie.h", "pose_only.h", "vel_ransac.h", "pose_graph.h")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = build()
        L = C.CDLL(so)
        L.oracle_create.restype = C.c_void_p
        L.oracle_create.argtypes = [C.POINTER(CProblem)]
        for name in dir(L):
            pass
        _LIB = L
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class Oracle:
    def __init__(self, prob, threads=1):
        self.prob = prob
        self._c = prob.to_c()
        self.L = lib()
        self.h = C.c_void_p(self.L.oracle_create(C.byref(self._c)))
        self.L.oracle_set_threads(self.h, int(threads))
        self.info = None

    def close(self):
        if self.h:
            self.L.oracle_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_extrinsics(self, free, prior_q=None, prior_info=None):
        """free [n_cam] uint8; prior_q [n_cam][4] (R_ini, xyzw) + prior_info [n_cam][3][3], or None for no EdgeExtrinsicPrior"""
        f = np.ascontiguousarray(free, np.uint8)
        q = None if prior_q is None else np.ascontiguousarray(prior_q, np.float64)
        w = None if prior_info is None else np.ascontiguousarray(prior_info, np.float64)
        self.L.oracle_set_extrinsics(self.h, _p(f), None if q is None else _p(q), None if w is None else _p(w))

    def extrinsics(self):
        a = np.zeros((self.prob.n_cam, 7)); self.L.oracle_get_extrinsics(self.h, _p(a)); return a

    def count_camera_observations(self):
        a = np.zeros(self.prob.n_cam, np.int64); self.L.oracle_count_camera_observations(self.h, _p(a)); return a

    def reset_state(self, kf_pose, kf_vel, pt_xyz):
        self.L.oracle_reset_state(self.h, _p(np.ascontiguousarray(kf_pose, np.float64)), _p(np.ascontiguousarray(kf_vel, np.float64)),
                                  _p(np.ascontiguousarray(pt_xyz, np.float64)))

    def batch_stats(self, reset=False):
        """seconds per stage under the G2OBatchStatistics names (g2o/core/batch_stats.h:39-78)"""
        a = (C.c_double * 5)()
        self.L.oracle_batch_stats(self.h, a, int(reset))
        return dict(zip(("timeResiduals", "timeQuadraticForm", "timeSchurComplement", "timeLinearSolver", "timeUpdate"), [float(v) for v in a]))

    def build_structure(self):
        info = StructureInfo()
        self.L.oracle_build_structure(self.h, C.byref(info))
        self.info = info
        return info

    def _pattern(self, fn, n):
        r = np.zeros(n, np.int32); c = np.zeros(n, np.int32)
        fn(self.h, _p(r), _p(c))
        return r, c

    def hpp_pattern(self):
        return self._pattern(self.L.oracle_get_hpp_pattern, self.info.n_hpp)

    def hschur_pattern(self):
        return self._pattern(self.L.oracle_get_hschur_pattern, self.info.n_hschur)

    def compute_errors(self):
        chi = C.c_double()
        self.L.oracle_compute_errors(self.h, C.byref(chi))
        return chi.value

    def build_system(self):
        self.L.oracle_build_system(self.h)

    def set_lambda(self, lam, backup=True):
        self.L.oracle_set_lambda(self.h, C.c_double(lam), int(backup))

    def restore_diagonal(self):
        self.L.oracle_restore_diagonal(self.h)

    def solve(self):
        ok = C.c_int()
        self.L.oracle_solve(self.h, C.byref(ok))
        return bool(ok.value)

    def vector_size(self):
        n = C.c_int64()
        self.L.oracle_vector_size(self.h, C.byref(n))
        return n.value

    def x(self):
        a = np.zeros(self.vector_size()); self.L.oracle_get_x(self.h, _p(a)); return a

    def b(self):
        a = np.zeros(self.vector_size()); self.L.oracle_get_b(self.h, _p(a)); return a

    def hpp(self):
        a = np.zeros((self.info.n_hpp, 12, 12)); self.L.oracle_get_hpp(self.h, _p(a)); return a

    def hschur(self):
        a = np.zeros((self.info.n_hschur, 12, 12)); bs = np.zeros(self.info.n_free_kf * 12)
        self.L.oracle_get_hschur(self.h, _p(a), _p(bs))
        return a, bs

    def hll(self):
        a = np.zeros((self.info.n_active_pt, 3, 3)); self.L.oracle_get_hll(self.h, _p(a)); return a

    def hpl(self):
        beg = np.zeros(self.info.n_active_pt + 1, np.int64); pose = np.zeros(self.info.n_hpl, np.int32)
        blk = np.zeros((self.info.n_hpl, 12, 3))
        self.L.oracle_get_hpl(self.h, _p(beg), _p(pose), _p(blk))
        return beg, pose, blk

    def oplus(self, x=None):
        self.L.oracle_oplus(self.h, None if x is None else _p(np.ascontiguousarray(x, np.float64)))

    def push(self):
        self.L.oracle_push(self.h)

    def pop(self):
        self.L.oracle_pop(self.h)

    def discard_top(self):
        self.L.oracle_discard_top(self.h)

    def optimize(self, iters=10, params=None):
        tr = LmTrace()
        self.L.oracle_optimize(self.h, int(iters), None, C.byref(params) if params is not None else None, C.byref(tr))
        return tr

    def state(self):
        P = self.prob
        kp = np.zeros((P.n_kf, 7)); kv = np.zeros((P.n_kf, 6)); pt = np.zeros((P.n_pt, 3))
        self.L.oracle_download_state(self.h, _p(kp), _p(kv), _p(pt))
        return kp, kv, pt

    def edge_chi2(self):
        a = np.zeros(self.prob.n_obs); self.L.oracle_edge_chi2(self.h, _p(a)); return a

    def edge_errors(self):
        a = np.zeros((self.prob.n_obs, 3)); self.L.oracle_edge_errors(self.h, _p(a)); return a

    def active_robust_chi2(self):
        c = C.c_double(); self.L.oracle_active_robust_chi2(self.h, C.byref(c)); return c.value

    def outlier_flags(self, th=None):
        th = th or Thresholds.local_gpba()
        f = np.zeros(self.prob.n_obs, np.uint8)
        self.L.oracle_outlier_flags(self.h, C.byref(th), _p(f))
        return f

    def set_levels(self, level):
        self.L.oracle_set_levels(self.h, _p(np.ascontiguousarray(level, np.uint8)))

    def set_robust_kernel(self, enabled):
        self.L.oracle_set_robust_kernel(self.h, int(enabled))

    def compute_errors_inactive(self):
        self.L.oracle_compute_errors_inactive(self.h)

    def rejection_rounds(self, n_rounds=4, iters=10, th=None, params=None):
        th = th or Thresholds.local_gpba()
        f = np.zeros(self.prob.n_obs, np.uint8)
        traces = (LmTrace * n_rounds)()
        self.L.oracle_rejection_rounds(self.h, n_rounds, iters, C.byref(th),
                                       C.byref(params) if params is not None else None, _p(f), traces)
        return f, list(traces)


# ---- math probes -------------------------------------------------------------------------------
def _d(a):
    return np.ascontiguousarray(a, np.float64)


def se3_exp(xi):
    o = np.zeros(7); lib().oracle_se3_exp(_p(_d(xi)), _p(o)); return o


def se3_log(T7):
    o = np.zeros(6); lib().oracle_se3_log(_p(_d(T7)), _p(o)); return o


def se3_mul(a, b):
    o = np.zeros(7); lib().oracle_se3_mul(_p(_d(a)), _p(_d(b)), _p(o)); return o


def se3_inv(a):
    o = np.zeros(7); lib().oracle_se3_inv(_p(_d(a)), _p(o)); return o


def se3_adj(a):
    o = np.zeros((6, 6)); lib().oracle_se3_adj(_p(_d(a)), _p(o)); return o


def se3_matrix(a):
    o = np.zeros((3, 4)); lib().oracle_se3_matrix(_p(_d(a)), _p(o))
    return np.vstack([o, [0, 0, 0, 1]])


def se3_act(a, p):
    o = np.zeros(3); lib().oracle_se3_act(_p(_d(a)), _p(_d(p)), _p(o)); return o


def jac_pose3(xi, which):
    """which: 0 Jl, 1 Jr, 2 Jl^-1, 3 Jr^-1, 4 se3Adj (curly hat)."""
    o = np.zeros((6, 6)); lib().oracle_jac_pose3(_p(_d(xi)), int(which), _p(o)); return o


def query_pose(qc, T1, T2, v1, v2, t1, t2, t):
    o = np.zeros(7); A = np.zeros((6, 12)); P = np.zeros((6, 12))
    lib().oracle_query_pose(_p(_d(qc)), _p(_d(T1)), _p(_d(T2)), _p(_d(v1)), _p(_d(v2)), C.c_double(t1), C.c_double(t2),
                            C.c_double(t), _p(o), _p(A), _p(P))
    return o, A, P


def edge_eval(qc, gp, T1, v1, t1, T2, v2, t2, t, Tbc, intr, bf, Xw, obs3, jac=True):
    dim = 3 if obs3[2] >= 0 else 2
    err = np.zeros(3); J1 = np.zeros((dim, 12)); J2 = np.zeros((dim, 12)); Jp = np.zeros((dim, 3))
    lib().oracle_edge_eval(_p(_d(qc)), int(gp), _p(_d(T1)), _p(_d(v1)), C.c_double(t1), _p(_d(T2)), _p(_d(v2)),
                           C.c_double(t2), C.c_double(t), _p(_d(Tbc)), _p(_d(intr)), C.c_double(bf), _p(_d(Xw)),
                           _p(_d(obs3)), _p(err), _p(J1) if jac else None, _p(J2) if jac else None,
                           _p(Jp) if jac else None)
    return err[:dim], J1, J2, Jp


def edge_jext(qc, gp, T1, v1, t1, T2, v2, t2, t, Tbc, intr, bf, Xw, obs3):
    dim = 3 if obs3[2] >= 0 else 2
    J = np.zeros((dim, 6))
    lib().oracle_edge_jext(_p(_d(qc)), int(gp), _p(_d(T1)), _p(_d(v1)), C.c_double(t1), _p(_d(T2)), _p(_d(v2)), C.c_double(t2),
                           C.c_double(t), _p(_d(Tbc)), _p(_d(intr)), C.c_double(bf), _p(_d(Xw)), _p(_d(obs3)), _p(J))
    return J


def ext_prior_eval(q_ini, Tbc):
    e = np.zeros(3); J = np.zeros((3, 3))
    lib().oracle_ext_prior_eval(_p(_d(q_ini)), _p(_d(Tbc)), _p(e), _p(J))
    return e, J


def prior_eval(T1, v1, t1, T2, v2, t2):
    e = np.zeros(12); Ji = np.zeros((12, 12)); Jj = np.zeros((12, 12))
    lib().oracle_prior_eval(_p(_d(T1)), _p(_d(v1)), C.c_double(t1), _p(_d(T2)), _p(_d(v2)), C.c_double(t2), _p(e),
                            _p(Ji), _p(Jj))
    return e, Ji, Jj


def vel_edge_eval(Tlast, Tbc, intr, dt, vel, Xw, obs2):
    """EdgeVelReproj on one match: error (2), Jacobian (2 x 6)."""
    e = np.zeros(2); J = np.zeros((2, 6))
    lib().oracle_vel_edge_eval(_p(_d(Tlast)), _p(_d(Tbc)), _p(_d(intr)), C.c_double(dt), _p(_d(vel)), _p(_d(Xw)), _p(_d(obs2)),
                               _p(e), _p(J))
    return e, J


def posevel_update(T7, v, upd12):
    To = np.zeros(7); vo = np.zeros(6)
    lib().oracle_posevel_update(_p(_d(T7)), _p(_d(v)), _p(_d(upd12)), _p(To), _p(vo))
    return To, vo


def right_jacobian_so3_orb(w):
    o = np.zeros((3, 3)); lib().oracle_right_jacobian_so3_orb(_p(_d(w)), _p(o)); return o


def huber(delta, e):
    r = np.zeros(3); lib().oracle_huber(C.c_double(delta), C.c_double(e), _p(r)); return r


def ldlt_dense(A, b):
    n = len(b); x = np.zeros(n)
    ok = lib().oracle_ldlt_dense(n, _p(_d(A)), _p(_d(b)), _p(x))
    return bool(ok), x


def pose_optimize(B):
    """CPU restatement of Optimizer::PoseGPOptimizationFromeLastFrame on a pygpba.pose.PoseBatch (oracle/pose_only.h)."""
    from pygpba.pose import PoseResult
    c = B.to_c()
    R = PoseResult(B)
    lib().oracle_pose_optimize(C.byref(c), *R.args())
    return R


def vel_ransac(B):
    """CPU restatement of Tracking::MCRansac's hypotheses (Optimizer::OptimizeVel each) on a pygpba.velransac.VelBatch
    (oracle/vel_ransac.h)."""
    from pygpba.velransac import VelResult
    c = B.to_c()
    R = VelResult(B)
    lib().oracle_vel_ransac(C.byref(c), *R.args())
    return R


def pose_graph_optimize(G, iters=20, params=None):
    """CPU restatement of the optimisation inside Optimizer::OptimizeEssentialGraph on a pygpba.posegraph.PoseGraph"""
    c = G.to_c()
    out = np.zeros((G.n_kf, 8)); tr = LmTrace()
    lib().oracle_pose_graph_optimize(C.byref(c), int(iters), C.byref(params) if params is not None else None, _p(out), C.byref(tr))
    return out, tr


def correct_points(xyz, ref_kf, sim3_before, sim3_after):
    x = _d(xyz); r = np.ascontiguousarray(ref_kf, np.int32); out = np.zeros_like(x)
    lib().oracle_correct_points(C.c_int64(len(x)), _p(x), _p(r), _p(_d(sim3_before)), _p(_d(sim3_after)), _p(out))
    return out


def sim3_exp(u):
    o = np.zeros(8); lib().oracle_sim3_exp(_p(_d(u)), _p(o)); return o


def sim3_log(S):
    o = np.zeros(7); lib().oracle_sim3_log(_p(_d(S)), _p(o)); return o


def sim3_mul(a, b):
    o = np.zeros(8); lib().oracle_sim3_mul(_p(_d(a)), _p(_d(b)), _p(o)); return o


def sim3_inv(a):
    o = np.zeros(8); lib().oracle_sim3_inv(_p(_d(a)), _p(o)); return o


def pose_system(B, f):
    """(H, b, chi2) of frame f of a PoseBatch at its initial estimate (oracle/pose_only.h build_system)"""
    c = B.to_c()
    H = np.zeros(24 * 24); b = np.zeros(24); chi2 = C.c_double(0)
    L = lib()
    n = L.oracle_pose_system(C.byref(c), int(f), H.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p), C.byref(chi2))
    return H[:n * n].reshape(n, n).copy(), b[:n].copy(), chi2.value


def pose_chi2_at(B, f, delta):
    c = B.to_c()
    d = np.ascontiguousarray(delta, np.float64)
    L = lib()
    L.oracle_pose_chi2_at.restype = C.c_double
    return L.oracle_pose_chi2_at(C.byref(c), int(f), d.ctypes.data_as(C.c_void_p))
