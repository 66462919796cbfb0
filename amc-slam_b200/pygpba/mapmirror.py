"""ctypes binding of include/gpba_map.h: the persistent SoA mirror of the map and the flattening of LocalGPBA /
BundleAdjustment graphs into gpba_problem arrays (host side)."""
import ctypes as C

import numpy as np

from . import lib as gl
from .problem import CProblem, Problem

SYMBOLS = [
    "gpba_map_last_error", "gpba_map_create", "gpba_map_destroy", "gpba_map_add_keyframe", "gpba_map_set_keyframe_state",
    "gpba_map_set_keyframe_bad", "gpba_map_add_point", "gpba_map_set_point", "gpba_map_set_point_bad",
    "gpba_map_add_observation", "gpba_map_add_observations", "gpba_map_erase_observation", "gpba_map_update_connections", "gpba_map_covisibles", "gpba_map_stats", "gpba_map_local_window",
    "gpba_map_global_window", "gpba_window_destroy", "gpba_window_problem", "gpba_window_iterations", "gpba_window_ids",
    "gpba_window_cam_obs", "gpba_window_apply", "gpba_map_apply_extrinsics", "gpba_map_extrinsics",
]


class CMapConfig(C.Structure):
    _fields_ = [("n_cam", C.c_int32), ("cam_intr", C.POINTER(C.c_double)), ("cam_Tbc", C.POINTER(C.c_double)),
                ("bf", C.c_double), ("qc", C.c_double * 6)]


_ready = False


def _lib():
    global _ready
    L = gl.lib()
    if not _ready:
        L.gpba_map_last_error.restype = C.c_char_p
        L.gpba_window_problem.restype = C.POINTER(CProblem)
        L.gpba_window_problem.argtypes = [C.c_void_p]
        L.gpba_window_iterations.argtypes = [C.c_void_p]
        L.gpba_window_destroy.argtypes = [C.c_void_p]
        L.gpba_map_destroy.argtypes = [C.c_void_p]
        L.gpba_map_add_keyframe.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p]
        L.gpba_map_set_keyframe_state.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
        L.gpba_map_set_keyframe_bad.argtypes = [C.c_void_p, C.c_int64]
        L.gpba_map_add_point.argtypes = [C.c_void_p, C.c_int64, C.c_void_p]
        L.gpba_map_set_point.argtypes = [C.c_void_p, C.c_int64, C.c_void_p]
        L.gpba_map_set_point_bad.argtypes = [C.c_void_p, C.c_int64]
        L.gpba_map_add_observation.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_int64, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int32]
        L.gpba_map_add_observations.argtypes = [C.c_void_p, C.c_int64] + [C.c_void_p] * 9
        L.gpba_map_erase_observation.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_int64]
        L.gpba_map_stats.argtypes = [C.c_void_p, C.c_void_p]
        L.gpba_map_update_connections.argtypes = [C.c_void_p, C.c_int64]
        L.gpba_map_covisibles.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]
        L.gpba_map_local_window.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_int32, C.POINTER(C.c_void_p)]
        L.gpba_map_global_window.argtypes = [C.c_void_p, C.c_int64, C.POINTER(C.c_void_p)]
        L.gpba_window_ids.argtypes = [C.c_void_p] + [C.c_void_p] * 6
        L.gpba_window_cam_obs.argtypes = [C.c_void_p, C.c_void_p]
        L.gpba_window_apply.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_float,
                                        C.c_void_p, C.c_void_p, C.c_void_p]
        _ready = True
    return L


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f64(a, n):
    a = np.ascontiguousarray(a, np.float64).reshape(-1)
    assert a.size == n
    return a


class Window:
    """A flattened BA graph (gpba_window). `.problem` is a pygpba Problem holding COPIES of the arrays."""

    def __init__(self, L, h, n_cam):
        self.L, self.h = L, h
        c = L.gpba_window_problem(h).contents
        def arr(ptr, n, dt):
            if n == 0 or not ptr:
                return np.zeros(0, dt)
            return np.ctypeslib.as_array(ptr, shape=(n,)).astype(dt, copy=True)
        n_kf, n_pt, n_rec, n_obs = c.n_kf, c.n_pt, c.n_rec, c.n_obs
        self.problem = Problem(
            cam_intr=arr(c.cam_intr, 4 * c.n_cam, np.float64), cam_Tbc=arr(c.cam_Tbc, 7 * c.n_cam, np.float64), bf=c.bf,
            kf_pose=arr(c.kf_pose, 7 * n_kf, np.float64), kf_vel=arr(c.kf_vel, 6 * n_kf, np.float64), kf_time=arr(c.kf_time, n_kf, np.float64),
            kf_fixed=arr(c.kf_fixed, n_kf, np.uint8), pt_xyz=arr(c.pt_xyz, 3 * n_pt, np.float64),
            rec_kf1=arr(c.rec_kf1, n_rec, np.int32), rec_kf2=arr(c.rec_kf2, n_rec, np.int32), rec_cam=arr(c.rec_cam, n_rec, np.int32),
            rec_t=arr(c.rec_t, n_rec, np.float64), obs_u=arr(c.obs_u, n_obs, np.float64), obs_v=arr(c.obs_v, n_obs, np.float64),
            obs_ur=arr(c.obs_ur, n_obs, np.float64) if c.obs_ur else None, obs_inv_sigma2=arr(c.obs_inv_sigma2, n_obs, np.float64),
            obs_rec=arr(c.obs_rec, n_obs, np.int32), obs_pt=arr(c.obs_pt, n_obs, np.int32), obs_flags=arr(c.obs_flags, n_obs, np.uint8),
            prior_kf1=arr(c.prior_kf1, c.n_prior, np.int32), prior_kf2=arr(c.prior_kf2, c.n_prior, np.int32),
            velp_kf=arr(c.velp_kf, c.n_velp, np.int32), qc=np.array(list(c.qc)), huber_mono=c.huber_mono, huber_stereo=c.huber_stereo,
            huber_prior=c.huber_prior, lambda_init=c.lambda_init, linear_solver=c.linear_solver)
        self.iterations = int(L.gpba_window_iterations(h))
        self.kf_id = np.zeros(n_kf, np.int64); self.kf_role = np.zeros(n_kf, np.int32); self.pt_id = np.zeros(n_pt, np.int64)
        self.obs_kf = np.zeros(n_obs, np.int64); self.obs_cam = np.zeros(n_obs, np.int32); self.obs_pt_id = np.zeros(n_obs, np.int64)
        L.gpba_window_ids(h, _p(self.kf_id), _p(self.kf_role), _p(self.pt_id), _p(self.obs_kf), _p(self.obs_cam), _p(self.obs_pt_id))
        self.cam_obs = np.zeros(n_cam, np.int32)
        L.gpba_window_cam_obs(h, _p(self.cam_obs))

    def c_problem(self):
        """the window's own gpba_problem (no copy): what a C++ caller hands to gpba_create"""
        return self.L.gpba_window_problem(self.h)

    def close(self):
        if self.h:
            self.L.gpba_window_destroy(self.h)
            self.h = None

    __del__ = close


class MapMirror:
    def __init__(self, cam_intr, cam_Tbc, bf, qc):
        self.L = _lib()
        self._intr = np.ascontiguousarray(cam_intr, np.float64).reshape(-1, 4)
        self._tbc = np.ascontiguousarray(cam_Tbc, np.float64).reshape(-1, 7)
        self.n_cam = len(self._intr)
        cfg = CMapConfig()
        cfg.n_cam = self.n_cam
        cfg.cam_intr = self._intr.ctypes.data_as(C.POINTER(C.c_double)); cfg.cam_Tbc = self._tbc.ctypes.data_as(C.POINTER(C.c_double))
        cfg.bf = float(bf)
        for i in range(6):
            cfg.qc[i] = float(qc[i])
        h = C.c_void_p()
        self._ck(self.L.gpba_map_create(C.byref(cfg), C.byref(h)))
        self.h = h

    def _ck(self, rc):
        if rc != 0:
            raise gl.GpbaError(f"gpba_map call failed ({rc}): {self.L.gpba_map_last_error().decode()}")

    def add_keyframe(self, id, prev_id, pose, vel, time, cam_time):
        self._ck(self.L.gpba_map_add_keyframe(self.h, int(id), int(prev_id), _p(_f64(pose, 7)), _p(_f64(vel, 6)), float(time), _p(_f64(cam_time, self.n_cam))))

    def set_keyframe_state(self, id, pose, vel=None):
        self._ck(self.L.gpba_map_set_keyframe_state(self.h, int(id), _p(_f64(pose, 7)), None if vel is None else _p(_f64(vel, 6))))

    def set_keyframe_bad(self, id):
        self._ck(self.L.gpba_map_set_keyframe_bad(self.h, int(id)))

    def add_point(self, id, xyz):
        self._ck(self.L.gpba_map_add_point(self.h, int(id), _p(_f64(xyz, 3))))

    def set_point(self, id, xyz):
        self._ck(self.L.gpba_map_set_point(self.h, int(id), _p(_f64(xyz, 3))))

    def set_point_bad(self, id):
        self._ck(self.L.gpba_map_set_point_bad(self.h, int(id)))

    def add_observation(self, kf, cam, pt, u, v, ur, inv_sigma2, close):
        self._ck(self.L.gpba_map_add_observation(self.h, int(kf), int(cam), int(pt), float(u), float(v), float(ur), float(inv_sigma2), int(close)))

    def add_observations(self, kf, cam, pt, u, v, ur, inv_sigma2, close):
        """bulk form of add_observation (array order); ur / close may be None"""
        kf = np.ascontiguousarray(kf, np.int64); cam = np.ascontiguousarray(cam, np.int32); pt = np.ascontiguousarray(pt, np.int64)
        u = np.ascontiguousarray(u, np.float64); v = np.ascontiguousarray(v, np.float64); w = np.ascontiguousarray(inv_sigma2, np.float64)
        ur = None if ur is None else np.ascontiguousarray(ur, np.float64)
        cl = None if close is None else np.ascontiguousarray(close, np.uint8)
        assert len(cam) == len(pt) == len(u) == len(v) == len(w) == len(kf)
        done = C.c_int64(0)
        self._ck(self.L.gpba_map_add_observations(self.h, len(kf), _p(kf), _p(cam), _p(pt), _p(u), _p(v), _p(ur), _p(w), _p(cl), C.byref(done)))
        return int(done.value)

    def erase_observation(self, kf, cam, pt):
        self._ck(self.L.gpba_map_erase_observation(self.h, int(kf), int(cam), int(pt)))

    def update_connections(self, kf):
        self._ck(self.L.gpba_map_update_connections(self.h, int(kf)))

    def covisibles(self, kf):
        """(ids, weights) of MultiKeyFrame::GetVectorCovisibleKeyFrames as the mirror maintains it"""
        n = C.c_int32(0)
        self._ck(self.L.gpba_map_covisibles(self.h, int(kf), None, None, 0, C.byref(n)))
        ids = np.zeros(max(n.value, 1), np.int64); w = np.zeros(max(n.value, 1), np.int32)
        self._ck(self.L.gpba_map_covisibles(self.h, int(kf), _p(ids), _p(w), n.value, C.byref(n)))
        return ids[:n.value].copy(), w[:n.value].copy()

    def stats(self):
        a = np.zeros(3, np.int64)
        self._ck(self.L.gpba_map_stats(self.h, _p(a)))
        return dict(keyframes=int(a[0]), points=int(a[1]), observations=int(a[2]))

    def local_window(self, kf_id, large=False, covisible=()):
        """covisible=None: use the mirror's own covisibility list (update_connections)"""
        h = C.c_void_p()
        if covisible is None:
            self._ck(self.L.gpba_map_local_window(self.h, int(kf_id), int(bool(large)), None, -1, C.byref(h)))
            return Window(self.L, h, self.n_cam)
        cov = np.ascontiguousarray(covisible, np.int64)
        self._ck(self.L.gpba_map_local_window(self.h, int(kf_id), int(bool(large)), _p(cov), len(cov), C.byref(h)))
        return Window(self.L, h, self.n_cam)

    def global_window(self, init_kf_id):
        h = C.c_void_p()
        self._ck(self.L.gpba_map_global_window(self.h, int(init_kf_id), C.byref(h)))
        return Window(self.L, h, self.n_cam)

    def apply(self, win, kf_pose=None, kf_vel=None, pt_xyz=None, flags=None, err=0.0, err_end=0.0):
        """returns (applied, indices of the erased observations in the window's arrays)"""
        applied = C.c_int32(0); n_er = C.c_int64(0)
        erased = np.zeros(max(win.problem.n_obs, 1), np.int64)
        f = None if flags is None else np.ascontiguousarray(flags, np.uint8)
        kp = None if kf_pose is None else np.ascontiguousarray(kf_pose, np.float64)
        kv = None if kf_vel is None else np.ascontiguousarray(kf_vel, np.float64)
        px = None if pt_xyz is None else np.ascontiguousarray(pt_xyz, np.float64)
        self._ck(self.L.gpba_window_apply(self.h, win.h, _p(kp), _p(kv), _p(px), _p(f), float(err), float(err_end),
                                          C.byref(applied), C.byref(n_er), _p(erased)))
        return bool(applied.value), erased[:n_er.value].copy()

    def apply_extrinsics(self, win, cam_Tbc, min_obs=50):
        """MultiKeyFrame::mTbc[c] = estimate.cast<float>() for cameras with cam_obs >= min_obs (Optimizer.cc:1419-1428)"""
        n = C.c_int32(0)
        t = np.ascontiguousarray(cam_Tbc, np.float64)
        self._ck(self.L.gpba_map_apply_extrinsics(C.c_void_p(self.h) if not isinstance(self.h, C.c_void_p) else self.h, win.h, _p(t), C.c_int32(int(min_obs)), C.byref(n)))
        return n.value

    def extrinsics(self):
        a = np.zeros((self.n_cam, 7))
        self._ck(self.L.gpba_map_extrinsics(C.c_void_p(self.h) if not isinstance(self.h, C.c_void_p) else self.h, _p(a)))
        return a

    def close(self):
        if self.h:
            self.L.gpba_map_destroy(self.h)
            self.h = None

    __del__ = close
