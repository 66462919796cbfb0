"""Pose-only GP optimisation batches (`gpba_pose_batch` of include/gpba.h): numpy owner + ctypes view, a seeded
generator, and the libgpba binding.  Mirrors Optimizer::PoseGPOptimizationFromeLastFrame (src/Optimizer.cc:369-686)."""
import ctypes as C

import numpy as np

from .problem import LmTrace, OBS_CLOSE, OBS_LEVEL1

GPBA_POSE_ROUNDS = 4
_pd = C.POINTER(C.c_double)
_pi = C.POINTER(C.c_int32)
_pl = C.POINTER(C.c_int64)
_pb = C.POINTER(C.c_uint8)


class CPoseBatch(C.Structure):
    _fields_ = [
        ("n_cam", C.c_int32), ("cam_intr", _pd), ("cam_Tbc", _pd), ("bf", C.c_double), ("qc", C.c_double * 6),
        ("n_frames", C.c_int32),
        ("prev_pose", _pd), ("prev_vel", _pd), ("prev_time", _pd), ("prev_fixed", _pb),
        ("cur_pose", _pd), ("cur_vel", _pd), ("cur_time", _pd), ("cam_time", _pd),
        ("obs_begin", _pl), ("obs_u", _pd), ("obs_v", _pd), ("obs_ur", _pd), ("obs_inv_sigma2", _pd), ("obs_xw", _pd),
        ("obs_cam", _pi), ("obs_flags", _pb),
        ("huber_mono", C.c_double), ("huber_stereo", C.c_double),
    ]


class PoseBatch:
    FIELDS = ("cam_intr", "cam_Tbc", "prev_pose", "prev_vel", "prev_time", "prev_fixed", "cur_pose", "cur_vel", "cur_time",
              "cam_time", "obs_begin", "obs_u", "obs_v", "obs_ur", "obs_inv_sigma2", "obs_xw", "obs_cam", "obs_flags")

    def __init__(self, **kw):
        f64 = lambda a: np.ascontiguousarray(a, np.float64)
        self.cam_intr = f64(kw["cam_intr"]).reshape(-1, 4)
        self.cam_Tbc = f64(kw["cam_Tbc"]).reshape(-1, 7)
        self.bf = float(kw["bf"]); self.qc = f64(kw["qc"])
        self.prev_pose = f64(kw["prev_pose"]).reshape(-1, 7); self.prev_vel = f64(kw["prev_vel"]).reshape(-1, 6)
        self.prev_time = f64(kw["prev_time"]); self.prev_fixed = np.ascontiguousarray(kw["prev_fixed"], np.uint8)
        self.cur_pose = f64(kw["cur_pose"]).reshape(-1, 7); self.cur_vel = f64(kw["cur_vel"]).reshape(-1, 6)
        self.cur_time = f64(kw["cur_time"]); self.cam_time = f64(kw["cam_time"]).reshape(len(self.cur_time), -1)
        self.obs_begin = np.ascontiguousarray(kw["obs_begin"], np.int64)
        self.obs_u = f64(kw["obs_u"]); self.obs_v = f64(kw["obs_v"])
        self.obs_ur = None if kw.get("obs_ur") is None else f64(kw["obs_ur"])
        self.obs_inv_sigma2 = f64(kw["obs_inv_sigma2"]); self.obs_xw = f64(kw["obs_xw"]).reshape(-1, 3)
        self.obs_cam = np.ascontiguousarray(kw["obs_cam"], np.int32)
        self.obs_flags = np.ascontiguousarray(kw["obs_flags"], np.uint8)
        self.huber_mono = float(kw["huber_mono"]); self.huber_stereo = float(kw["huber_stereo"])
        self.truth_outlier = kw.get("truth_outlier")

    n_frames = property(lambda s: len(s.cur_time))
    n_cam = property(lambda s: len(s.cam_intr))
    n_obs = property(lambda s: len(s.obs_u))

    def slice(self, f):
        """the batch holding frame f alone"""
        a, b = int(self.obs_begin[f]), int(self.obs_begin[f + 1])
        return PoseBatch(cam_intr=self.cam_intr, cam_Tbc=self.cam_Tbc, bf=self.bf, qc=self.qc,
                         prev_pose=self.prev_pose[f:f + 1], prev_vel=self.prev_vel[f:f + 1], prev_time=self.prev_time[f:f + 1],
                         prev_fixed=self.prev_fixed[f:f + 1], cur_pose=self.cur_pose[f:f + 1], cur_vel=self.cur_vel[f:f + 1],
                         cur_time=self.cur_time[f:f + 1], cam_time=self.cam_time[f:f + 1], obs_begin=[0, b - a],
                         obs_u=self.obs_u[a:b], obs_v=self.obs_v[a:b], obs_ur=None if self.obs_ur is None else self.obs_ur[a:b],
                         obs_inv_sigma2=self.obs_inv_sigma2[a:b], obs_xw=self.obs_xw[a:b], obs_cam=self.obs_cam[a:b],
                         obs_flags=self.obs_flags[a:b], huber_mono=self.huber_mono, huber_stereo=self.huber_stereo,
                         truth_outlier=None if self.truth_outlier is None else self.truth_outlier[a:b])

    def subset(self, keep):
        """the batch with only the matches selected by the boolean mask `keep` (frames may end up empty)"""
        keep = np.asarray(keep, bool)
        cnt = np.add.reduceat(np.concatenate([keep, [False]]).astype(np.int64), self.obs_begin[:-1].clip(max=len(keep)))
        cnt[np.diff(self.obs_begin) == 0] = 0
        return PoseBatch(cam_intr=self.cam_intr, cam_Tbc=self.cam_Tbc, bf=self.bf, qc=self.qc, prev_pose=self.prev_pose,
                         prev_vel=self.prev_vel, prev_time=self.prev_time, prev_fixed=self.prev_fixed, cur_pose=self.cur_pose,
                         cur_vel=self.cur_vel, cur_time=self.cur_time, cam_time=self.cam_time,
                         obs_begin=np.concatenate([[0], np.cumsum(cnt)]), obs_u=self.obs_u[keep], obs_v=self.obs_v[keep],
                         obs_ur=None if self.obs_ur is None else self.obs_ur[keep], obs_inv_sigma2=self.obs_inv_sigma2[keep],
                         obs_xw=self.obs_xw[keep], obs_cam=self.obs_cam[keep], obs_flags=self.obs_flags[keep],
                         huber_mono=self.huber_mono, huber_stereo=self.huber_stereo,
                         truth_outlier=None if self.truth_outlier is None else self.truth_outlier[keep])

    def to_c(self):
        def p(a, t):
            return a.ctypes.data_as(t) if a is not None and a.size else C.cast(None, t)
        c = CPoseBatch()
        c.n_cam, c.cam_intr, c.cam_Tbc, c.bf = self.n_cam, p(self.cam_intr, _pd), p(self.cam_Tbc, _pd), self.bf
        for i in range(6):
            c.qc[i] = float(self.qc[i])
        c.n_frames = self.n_frames
        c.prev_pose, c.prev_vel, c.prev_time, c.prev_fixed = p(self.prev_pose, _pd), p(self.prev_vel, _pd), p(self.prev_time, _pd), p(self.prev_fixed, _pb)
        c.cur_pose, c.cur_vel, c.cur_time, c.cam_time = p(self.cur_pose, _pd), p(self.cur_vel, _pd), p(self.cur_time, _pd), p(self.cam_time, _pd)
        c.obs_begin, c.obs_u, c.obs_v, c.obs_ur = p(self.obs_begin, _pl), p(self.obs_u, _pd), p(self.obs_v, _pd), p(self.obs_ur, _pd)
        c.obs_inv_sigma2, c.obs_xw, c.obs_cam, c.obs_flags = p(self.obs_inv_sigma2, _pd), p(self.obs_xw, _pd), p(self.obs_cam, _pi), p(self.obs_flags, _pb)
        c.huber_mono, c.huber_stereo = self.huber_mono, self.huber_stereo
        return c

    def input_bytes(self):
        return int(sum(getattr(self, f).nbytes for f in self.FIELDS if getattr(self, f) is not None))


def shard_frames(B, rank, world):
    """Multi-GPU use of the pose-only path: frames are independent, so a batch shards without any data-path collective
    ("replicas" in DESIGN terms).  Contiguous frame ranges balanced by match count; returns (sub-batch, frame indices)."""
    cnt = np.concatenate([[0], np.cumsum(np.diff(B.obs_begin))]).astype(np.int64)
    total = int(cnt[-1])
    cuts = [0] + [int(np.searchsorted(cnt, total * r // world, side="left")) for r in range(1, world)] + [B.n_frames]
    cuts = np.maximum.accumulate(np.clip(cuts, 0, B.n_frames))
    lo, hi = int(cuts[rank]), int(cuts[rank + 1])
    a, b = int(B.obs_begin[lo]), int(B.obs_begin[hi])
    sub = PoseBatch(cam_intr=B.cam_intr, cam_Tbc=B.cam_Tbc, bf=B.bf, qc=B.qc, prev_pose=B.prev_pose[lo:hi], prev_vel=B.prev_vel[lo:hi],
                    prev_time=B.prev_time[lo:hi], prev_fixed=B.prev_fixed[lo:hi], cur_pose=B.cur_pose[lo:hi], cur_vel=B.cur_vel[lo:hi],
                    cur_time=B.cur_time[lo:hi], cam_time=B.cam_time[lo:hi], obs_begin=B.obs_begin[lo:hi + 1] - a, obs_u=B.obs_u[a:b],
                    obs_v=B.obs_v[a:b], obs_ur=None if B.obs_ur is None else B.obs_ur[a:b], obs_inv_sigma2=B.obs_inv_sigma2[a:b],
                    obs_xw=B.obs_xw[a:b], obs_cam=B.obs_cam[a:b], obs_flags=B.obs_flags[a:b], huber_mono=B.huber_mono,
                    huber_stereo=B.huber_stereo, truth_outlier=None if B.truth_outlier is None else B.truth_outlier[a:b])
    return sub, np.arange(lo, hi)


def make_pose_batch(n_frames=4, n_pt=600, A=2, outliers=0.15, seed=51, fix_prev=True, stereo_fraction=0.0, obs_per_pt=None):
    """Frames cut out of a seeded synthetic map (synth.make_problem): frame f = keyframe f+1 of the map with keyframe f as
    the previous frame; its matches are the map's observations of that keyframe against the (noisy, then frozen) map
    points; a fraction of the matches is replaced by uniform pixels (wrong associations)."""
    from . import synth
    n_kf = n_frames + 1
    P = synth.make_problem("c1", A=A, n_kf=n_kf, n_pt=n_pt, obs_per_pt=obs_per_pt or min(3 * (A + 1) * 2, 3 * n_kf), seed=seed, outliers=0.0)
    rng = np.random.default_rng(seed + 1000)
    n_cam = P.n_cam
    # tracked map points are well triangulated: true position + 1 cm, float-rounded like MapPoint::GetWorldPos
    xw_all = (P.truth["pt"] + rng.normal(size=P.truth["pt"].shape) * 0.01).astype(np.float32).astype(np.float64)
    rec_kf2, rec_cam = P.rec_kf2, P.rec_cam
    cam_time = np.zeros((n_kf, n_cam))
    cam_time[rec_kf2, rec_cam] = P.rec_t
    begin, cols = [0], {k: [] for k in ("u", "v", "ur", "w", "xw", "cam", "fl", "out")}
    for f in range(n_frames):
        k = f + 1
        sel = np.nonzero(rec_kf2[P.obs_rec] == k)[0]
        u, v = P.obs_u[sel].copy(), P.obs_v[sel].copy()
        out = rng.uniform(size=len(sel)) < outliers
        u[out] = np.float32(rng.uniform(0, synth.IMG_W, out.sum())); v[out] = np.float32(rng.uniform(0, synth.IMG_H, out.sum()))
        cam = rec_cam[P.obs_rec[sel]]
        ur = -np.ones(len(sel))
        if stereo_fraction > 0:
            pick = (cam == n_cam - 1) & (rng.uniform(size=len(sel)) < stereo_fraction)
            ur[pick] = np.float32(u[pick] - P.bf / rng.uniform(8.0, 40.0, pick.sum()))
        cols["u"].append(u); cols["v"].append(v); cols["ur"].append(ur); cols["w"].append(P.obs_inv_sigma2[sel])
        cols["xw"].append(xw_all[P.obs_pt[sel]]); cols["cam"].append(cam); cols["fl"].append(P.obs_flags[sel] & OBS_CLOSE); cols["out"].append(out)
        begin.append(begin[-1] + len(sel))
    cat = lambda k: np.concatenate(cols[k])
    # the previous frame has already been optimised: true state, float-rounded like Frame::GetPose / GetVelocity
    q = P.truth["kf_q"].astype(np.float32).astype(np.float64)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    prev_pose = np.concatenate([q, P.truth["kf_t"].astype(np.float32).astype(np.float64)], 1)[:n_frames]
    prev_vel = P.truth["kf_vel"].astype(np.float32).astype(np.float64)[:n_frames]
    return PoseBatch(cam_intr=P.cam_intr, cam_Tbc=P.cam_Tbc, bf=P.bf, qc=P.qc,
                     prev_pose=prev_pose, prev_vel=prev_vel, prev_time=P.kf_time[:n_frames],
                     prev_fixed=np.full(n_frames, 1 if fix_prev else 0, np.uint8),
                     cur_pose=P.kf_pose[1:], cur_vel=P.kf_vel[1:], cur_time=P.kf_time[1:], cam_time=cam_time[1:],
                     obs_begin=begin, obs_u=cat("u"), obs_v=cat("v"), obs_ur=cat("ur") if stereo_fraction > 0 else None,
                     obs_inv_sigma2=cat("w"), obs_xw=cat("xw"), obs_cam=cat("cam"), obs_flags=cat("fl"),
                     huber_mono=P.huber_mono, huber_stereo=P.huber_stereo, truth_outlier=cat("out"))


class PoseResult:
    def __init__(self, B):
        self.cur_pose = np.zeros((B.n_frames, 7)); self.cur_vel = np.zeros((B.n_frames, 6))
        self.prev_pose = np.zeros((B.n_frames, 7)); self.prev_vel = np.zeros((B.n_frames, 6))
        self.outlier = np.zeros(B.n_obs, np.uint8); self.n_inliers = np.zeros(B.n_frames, np.int32)
        self.traces = (LmTrace * (B.n_frames * GPBA_POSE_ROUNDS))()

    def args(self):
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        return (p(self.cur_pose), p(self.cur_vel), p(self.prev_pose), p(self.prev_vel), p(self.outlier), p(self.n_inliers), self.traces)

    def trace(self, f, rnd):
        return self.traces[f * GPBA_POSE_ROUNDS + rnd].summary()


def pose_optimize(B, device=-1):
    """libgpba: one CTA per frame runs the whole 4 x 10 LM schedule on the device."""
    from . import lib as gl
    L = gl.lib()
    c = B.to_c()
    R = PoseResult(B)
    rc = L.gpba_pose_optimize(C.byref(c), int(device), *R.args())
    if rc != 0:
        raise gl.GpbaError(f"gpba_pose_optimize failed ({rc}): {L.gpba_last_error().decode()}")
    return R
