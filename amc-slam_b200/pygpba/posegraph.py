"""Essential-graph optimisation (include/gpba.h: gpba_pose_graph_optimize / gpba_correct_points): ctypes harness and a seeded
synthetic pose graph shaped like the one Optimizer::OptimizeEssentialGraph builds (src/Optimizer.cc:1434-1717): a trajectory
that closes a loop, spanning-tree edges between consecutive keyframes measured BEFORE the loop correction, covisibility
edges, and loop edges measured between the corrected poses of the loop's two ends."""
import ctypes as C

import numpy as np

from .problem import LmParams, LmTrace
from .synth import quat_mul, quat_rot, se3_exp, se3_mul


class CPoseGraph(C.Structure):
    _fields_ = [("n_kf", C.c_int32), ("sim3", C.POINTER(C.c_double)), ("fixed", C.POINTER(C.c_uint8)), ("fix_scale", C.c_int32),
                ("n_edge", C.c_int64), ("edge_i", C.POINTER(C.c_int32)), ("edge_j", C.POINTER(C.c_int32)),
                ("edge_meas", C.POINTER(C.c_double)), ("lambda_init", C.c_double)]


def sim3_inv(S):
    q = S[..., :4] * np.array([-1, -1, -1, 1.0])
    s = S[..., 7:8]
    t = quat_rot(q, -S[..., 4:7] / s)
    return np.concatenate([q, t, 1.0 / s], -1)


def sim3_mul(a, b):
    q = quat_mul(a[..., :4], b[..., :4])
    t = a[..., 7:8] * quat_rot(a[..., :4], b[..., 4:7]) + a[..., 4:7]
    return np.concatenate([q, t, a[..., 7:8] * b[..., 7:8]], -1)


class PoseGraph:
    def __init__(self, sim3, fixed, edge_i, edge_j, edge_meas, fix_scale=True, lambda_init=1e-16):
        self.sim3 = np.ascontiguousarray(sim3, np.float64).reshape(-1, 8)
        self.fixed = np.ascontiguousarray(fixed, np.uint8)
        self.edge_i = np.ascontiguousarray(edge_i, np.int32)
        self.edge_j = np.ascontiguousarray(edge_j, np.int32)
        self.edge_meas = np.ascontiguousarray(edge_meas, np.float64).reshape(-1, 8)
        self.fix_scale = bool(fix_scale)
        self.lambda_init = float(lambda_init)

    n_kf = property(lambda s: len(s.sim3))
    n_edge = property(lambda s: len(s.edge_i))

    def to_c(self):
        pd, pi, pb = C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(C.c_uint8)
        return CPoseGraph(self.n_kf, self.sim3.ctypes.data_as(pd), self.fixed.ctypes.data_as(pb), int(self.fix_scale), self.n_edge,
                          self.edge_i.ctypes.data_as(pi), self.edge_j.ctypes.data_as(pi), self.edge_meas.ctypes.data_as(pd), self.lambda_init)


def make_pose_graph(n_kf=200, seed=0, fix_scale=True, drift=0.02, covis=3, scale_drift=0.0):
    """A closed loop of n_kf keyframes.  World-to-camera Sim3 S_iw of the drifted odometry; the last `n_loop` keyframes are
    already corrected (CorrectedSim3: moved onto the loop's start), every other edge is measured between uncorrected poses."""
    rng = np.random.default_rng(seed)
    k = np.arange(n_kf)
    w0 = 2 * np.pi / (n_kf * 0.1)
    twist = np.tile([4.0, 0, 0, 0, 0, w0], (n_kf, 1))
    q = np.zeros((n_kf, 4)); q[0, 3] = 1; t = np.zeros((n_kf, 3))
    qd, td = se3_exp((twist + rng.normal(size=(n_kf, 6)) * drift * np.array([1, 1, 1, 0.1, 0.1, 0.1])) * 0.1)
    for i in range(1, n_kf):
        q[i], t[i] = se3_mul(q[i - 1], t[i - 1], qd[i - 1], td[i - 1])
    s = np.exp(np.cumsum(rng.normal(size=n_kf) * scale_drift)) if scale_drift > 0 else np.ones(n_kf)
    Twc = np.concatenate([q, t, s[:, None]], 1)           # camera-to-world with the odometry's scale
    Scw = sim3_inv(Twc)                                    # vertices hold S_iw (Optimizer.cc:1476-1480)
    ei, ej, meas = [], [], []

    def add(i, j, Si, Sj):                                  # S_ji = S_jw S_wi (:1527-1528)
        ei.append(i); ej.append(j); meas.append(sim3_mul(Sj, sim3_inv(Si)))
    for i in range(1, n_kf):                                # spanning tree: parent = previous keyframe (:1556-1579)
        add(i, i - 1, Scw[i], Scw[i - 1])
    for i in range(n_kf):                                   # covisibility edges to older keyframes (:1608-1640)
        for d in range(2, 2 + covis):
            if i - d >= 0:
                add(i, i - d, Scw[i], Scw[i - d])
    # the loop: the last keyframes see the first ones again; their corrected poses sit where the loop start says they are
    n_loop = max(3, n_kf // 40)
    corrected = Scw.copy()
    true_rel = sim3_inv(np.concatenate([np.array([0, 0, 0, 1.0]), np.zeros(3), [1.0]]))   # identity: the loop closes exactly
    anchor = sim3_mul(true_rel, Scw[0])
    shift = sim3_mul(anchor, sim3_inv(Scw[n_kf - 1]))       # moves the current keyframe onto the loop keyframe
    for i in range(n_kf - n_loop, n_kf):
        corrected[i] = sim3_mul(Scw[i], sim3_mul(sim3_inv(Scw[n_kf - 1]), sim3_mul(sim3_inv(shift), sim3_mul(shift, sim3_mul(shift, Scw[n_kf - 1])))))
        corrected[i] = sim3_mul(sim3_mul(Scw[i], sim3_inv(Scw[n_kf - 1])), anchor)
    for i in range(n_kf - n_loop, n_kf):                    # loop edges between CORRECTED poses (:1507-1543)
        for j in range(0, n_loop):
            add(i, j, corrected[i], corrected[j])
    fixed = np.zeros(n_kf, np.uint8); fixed[0] = 1          # the map's initial keyframe (:1484)
    return PoseGraph(corrected, fixed, ei, ej, np.array(meas), fix_scale=fix_scale)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def optimize(G, iters=20, params=None, device=-1):
    from .lib import lib, GpbaError
    L = lib()
    c = G.to_c()
    out = np.zeros((G.n_kf, 8)); tr = LmTrace()
    rc = L.gpba_pose_graph_optimize(C.byref(c), C.c_int(device), C.c_int(iters), C.byref(params) if params is not None else None, _p(out), C.byref(tr))
    if rc != 0:
        raise GpbaError(f"gpba_pose_graph_optimize failed ({rc}): {L.gpba_last_error().decode()}")
    return out, tr


def correct_points(xyz, ref_kf, sim3_before, sim3_after, device=-1):
    from .lib import lib, GpbaError
    L = lib()
    x = np.ascontiguousarray(xyz, np.float64); r = np.ascontiguousarray(ref_kf, np.int32)
    a = np.ascontiguousarray(sim3_before, np.float64); b = np.ascontiguousarray(sim3_after, np.float64)
    out = np.zeros_like(x)
    rc = L.gpba_correct_points(C.c_int(device), C.c_int64(len(x)), _p(x), _p(r), C.c_int32(len(a)), _p(a), _p(b), _p(out))
    if rc != 0:
        raise GpbaError(f"gpba_correct_points failed ({rc}): {L.gpba_last_error().decode()}")
    return out
