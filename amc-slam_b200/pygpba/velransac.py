"""Velocity RANSAC batches (`gpba_vel_batch` of include/gpba.h): numpy owner + ctypes view, a seeded generator and the
libgpba binding.  Mirrors Tracking::MCRansac / Optimizer::OptimizeVel (src/Tracking.cc:1939-2002, src/Optimizer.cc:2364-2447)."""
import ctypes as C

import numpy as np

from .problem import LmTrace

_pd = C.POINTER(C.c_double)
_pi = C.POINTER(C.c_int32)


class CVelBatch(C.Structure):
    _fields_ = [
        ("n_cam", C.c_int32), ("cam_intr", _pd), ("cam_Tbc", _pd), ("cam_dt", _pd), ("last_pose", C.c_double * 7),
        ("vel_init", C.c_double * 6), ("n_match", C.c_int32), ("obs_u", _pd), ("obs_v", _pd), ("obs_inv_sigma2", _pd),
        ("obs_xw", _pd), ("obs_cam", _pi), ("n_hyp", C.c_int32), ("set_size", C.c_int32), ("samples", _pi),
        ("huber_delta", C.c_double), ("threshold", C.c_double), ("iterations", C.c_int32),
    ]


class VelBatch:
    def __init__(self, **kw):
        f64 = lambda a: np.ascontiguousarray(a, np.float64)
        self.cam_intr = f64(kw["cam_intr"]).reshape(-1, 4); self.cam_Tbc = f64(kw["cam_Tbc"]).reshape(-1, 7); self.cam_dt = f64(kw["cam_dt"])
        self.last_pose = f64(kw["last_pose"]); self.vel_init = f64(kw["vel_init"])
        self.obs_u = f64(kw["obs_u"]); self.obs_v = f64(kw["obs_v"]); self.obs_inv_sigma2 = f64(kw["obs_inv_sigma2"])
        self.obs_xw = f64(kw["obs_xw"]).reshape(-1, 3); self.obs_cam = np.ascontiguousarray(kw["obs_cam"], np.int32)
        self.samples = np.ascontiguousarray(kw["samples"], np.int32).reshape(-1, int(kw.get("set_size", 3)))
        self.huber_delta = float(kw.get("huber_delta", 5.991)); self.threshold = float(kw.get("threshold", 2.0))
        self.iterations = int(kw.get("iterations", 40))
        self.truth_vel = kw.get("truth_vel"); self.truth_outlier = kw.get("truth_outlier")

    n_cam = property(lambda s: len(s.cam_intr))
    n_match = property(lambda s: len(s.obs_u))
    n_hyp = property(lambda s: s.samples.shape[0])
    set_size = property(lambda s: s.samples.shape[1])

    def to_c(self):
        p = lambda a, t: a.ctypes.data_as(t) if a.size else C.cast(None, t)
        c = CVelBatch()
        c.n_cam, c.cam_intr, c.cam_Tbc, c.cam_dt = self.n_cam, p(self.cam_intr, _pd), p(self.cam_Tbc, _pd), p(self.cam_dt, _pd)
        for i in range(7):
            c.last_pose[i] = float(self.last_pose[i])
        for i in range(6):
            c.vel_init[i] = float(self.vel_init[i])
        c.n_match, c.obs_u, c.obs_v, c.obs_inv_sigma2 = self.n_match, p(self.obs_u, _pd), p(self.obs_v, _pd), p(self.obs_inv_sigma2, _pd)
        c.obs_xw, c.obs_cam = p(self.obs_xw, _pd), p(self.obs_cam, _pi)
        c.n_hyp, c.set_size, c.samples = self.n_hyp, self.set_size, p(self.samples, _pi)
        c.huber_delta, c.threshold, c.iterations = self.huber_delta, self.threshold, self.iterations
        return c


def _quat_mul(a, b):
    ax, ay, az, aw = a; bx, by, bz, bw = b
    return np.array([aw * bx + ax * bw + ay * bz - az * by, aw * by - ax * bz + ay * bw + az * bx,
                     aw * bz + ax * by - ay * bx + az * bw, aw * bw - ax * bx - ay * by - az * bz])


def _quat_rot(q, p):
    R = _quat_R(q)
    return p @ R.T


def _quat_R(q):
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def _se3_exp(xi):
    """Sophus SE3::exp, tangent [upsilon; omega] (series form, fine for a generator)"""
    u, w = xi[:3], xi[3:]
    th = np.linalg.norm(w)
    W = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
    if th < 1e-9:
        R, Vm = np.eye(3) + W, np.eye(3) + 0.5 * W
    else:
        R = np.eye(3) + np.sin(th) / th * W + (1 - np.cos(th)) / th ** 2 * W @ W
        Vm = np.eye(3) + (1 - np.cos(th)) / th ** 2 * W + (th - np.sin(th)) / th ** 3 * W @ W
    return R, Vm @ u


def make_vel_batch(n_match=600, n_hyp=23, A=2, outliers=0.2, seed=71, set_size=3, vel_noise=0.3):
    """One frame pair: the last frame's pose, a constant body twist, matches of the current frame's cameras (captured at
    t_last + dt_cam) against exact world points, `outliers` of them replaced by uniform pixels; n_hyp sample sets."""
    from . import synth
    P = synth.make_problem("tiny", A=A)
    rng = np.random.default_rng(seed)
    n_cam = P.n_cam
    q = rng.normal(size=4); q /= np.linalg.norm(q)
    last_pose = np.concatenate([q, rng.normal(size=3) * 5])
    v_true = np.array([4.0, 0.1, -0.05, 0.01, -0.02, 0.15]) + rng.normal(size=6) * 0.05
    cam_dt = np.sort(rng.uniform(0.02, 0.1, n_cam))
    Rl, tl = _quat_R(last_pose[:4]), last_pose[4:]
    cam = rng.integers(0, n_cam, n_match).astype(np.int32)
    u = np.zeros(n_match); v = np.zeros(n_match); xw = np.zeros((n_match, 3)); lvl = rng.integers(0, 8, n_match)
    for c in range(n_cam):
        sel = np.nonzero(cam == c)[0]
        fx, fy, cx, cy = P.cam_intr[c]
        Re, te = _se3_exp(v_true * cam_dt[c])
        Rbc, tbc = _quat_R(P.cam_Tbc[c, :4]), P.cam_Tbc[c, 4:]
        Rwc = Rl @ Re @ Rbc
        twc = Rl @ (Re @ tbc + te) + tl
        z = rng.uniform(3, 30, len(sel))
        uu = rng.uniform(20, synth.IMG_W - 20, len(sel)); vv = rng.uniform(20, synth.IMG_H - 20, len(sel))
        Xc = np.stack([(uu - cx) / fx * z, (vv - cy) / fy * z, z], 1)
        xw[sel] = Xc @ Rwc.T + twc
        sig = 1.2 ** lvl[sel]
        u[sel] = np.float32(uu + rng.normal(size=len(sel)) * 0.3 * sig); v[sel] = np.float32(vv + rng.normal(size=len(sel)) * 0.3 * sig)
    out = rng.uniform(size=n_match) < outliers
    u[out] = np.float32(rng.uniform(0, synth.IMG_W, out.sum())); v[out] = np.float32(rng.uniform(0, synth.IMG_H, out.sum()))
    samples = np.stack([rng.choice(n_match, set_size, replace=False) for _ in range(n_hyp)]) if n_hyp else np.zeros((0, set_size), np.int32)
    return VelBatch(cam_intr=P.cam_intr, cam_Tbc=P.cam_Tbc, cam_dt=cam_dt, last_pose=last_pose,
                    vel_init=(v_true + rng.normal(size=6) * vel_noise).astype(np.float32), obs_u=u, obs_v=v,
                    obs_inv_sigma2=(1.2 ** (-2.0 * lvl)).astype(np.float32), obs_xw=xw.astype(np.float32), obs_cam=cam, samples=samples,
                    set_size=set_size, truth_vel=v_true, truth_outlier=out)


class VelResult:
    def __init__(self, B):
        self.vel = np.zeros((B.n_hyp, 6)); self.inliers = np.zeros(B.n_hyp, np.int32)
        self.mask = np.zeros((B.n_hyp, B.n_match), np.uint8); self.best = C.c_int32(-1)
        self.traces = (LmTrace * max(B.n_hyp, 1))()

    def args(self):
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        return (p(self.vel), p(self.inliers), p(self.mask), C.cast(C.byref(self.best), C.c_void_p), C.cast(self.traces, C.c_void_p))

    def trace(self, h):
        return self.traces[h].summary()


def vel_ransac(B, device=-1):
    """libgpba: all hypotheses of one Tracking::MCRansac call in one launch (one CTA per hypothesis)."""
    from . import lib as gl
    L = gl.lib()
    c = B.to_c()
    R = VelResult(B)
    rc = L.gpba_vel_ransac(C.byref(c), int(device), *R.args())
    if rc != 0:
        raise gl.GpbaError(f"gpba_vel_ransac failed ({rc}): {L.gpba_last_error().decode()}")
    return R
