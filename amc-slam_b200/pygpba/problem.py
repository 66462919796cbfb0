"""Flattened GP-BA problem (the `gpba_problem` struct of include/gpba.h) as numpy SoA + ctypes view.

This is the wire format the g2o-side adapter fills from the graph that Optimizer::BundleAdjustment
(src/Optimizer.cc:61-367) / Optimizer::LocalGPBA (:713-1432) build; see include/gpba.h for the
field-by-field citations.
"""
import ctypes as C
import numpy as np

GPBA_MAX_ITERS = 64
GPBA_N_STAGES = 11
OBS_CLOSE, OBS_LEVEL1, OBS_NO_KERNEL = 1, 2, 4
SOLVER_DENSE_CHOL, SOLVER_PCG, SOLVER_SPARSE_CHOL = 0, 1, 2

_pd = C.POINTER(C.c_double)
_pi = C.POINTER(C.c_int32)
_pb = C.POINTER(C.c_uint8)


class CProblem(C.Structure):
    _fields_ = [
        ("n_cam", C.c_int32), ("cam_intr", _pd), ("cam_Tbc", _pd), ("bf", C.c_double),
        ("n_kf", C.c_int32), ("kf_pose", _pd), ("kf_vel", _pd), ("kf_time", _pd), ("kf_fixed", _pb),
        ("n_pt", C.c_int32), ("pt_xyz", _pd),
        ("n_rec", C.c_int32), ("rec_kf1", _pi), ("rec_kf2", _pi), ("rec_cam", _pi), ("rec_t", _pd),
        ("n_obs", C.c_int64), ("obs_u", _pd), ("obs_v", _pd), ("obs_ur", _pd), ("obs_inv_sigma2", _pd),
        ("obs_rec", _pi), ("obs_pt", _pi), ("obs_flags", _pb),
        ("n_prior", C.c_int32), ("prior_kf1", _pi), ("prior_kf2", _pi),
        ("n_velp", C.c_int32), ("velp_kf", _pi),
        ("qc", C.c_double * 6),
        ("huber_mono", C.c_double), ("huber_stereo", C.c_double), ("huber_prior", C.c_double),
        ("lambda_init", C.c_double), ("linear_solver", C.c_int32),
    ]


class LmTrace(C.Structure):
    _fields_ = [
        ("n_iters", C.c_int32), ("result", C.c_int32),
        ("levenberg_iterations", C.c_int32 * GPBA_MAX_ITERS),
        ("chi2_before", C.c_double * GPBA_MAX_ITERS),
        ("chi2_after", C.c_double * GPBA_MAX_ITERS),
        ("lambda_", C.c_double * GPBA_MAX_ITERS),
        ("total_trials", C.c_int32), ("cg_iterations", C.c_int32),
        ("last_trial_chi2", C.c_double),
    ]

    def summary(self):
        n = self.n_iters
        return dict(n_iters=n, result=self.result, trials=list(self.levenberg_iterations[:n]),
                    chi2_before=list(self.chi2_before[:n]), chi2_after=list(self.chi2_after[:n]),
                    lam=list(self.lambda_[:n]), total_trials=self.total_trials, cg_iterations=self.cg_iterations,
                    last_trial_chi2=self.last_trial_chi2)


class LmParams(C.Structure):
    _fields_ = [("max_trials_after_failure", C.c_int32), ("tau", C.c_double), ("good_step_lower", C.c_double),
                ("good_step_upper", C.c_double), ("pcg_tolerance", C.c_double), ("pcg_max_iterations", C.c_int32)]


class Thresholds(C.Structure):
    _fields_ = [("chi2_mono", C.c_double), ("chi2_mono_close", C.c_double), ("chi2_stereo", C.c_double)]

    @staticmethod
    def local_gpba():
        """Float-typed thresholds of LocalGPBA (Optimizer.cc:975-978, 1273, 1289)."""
        m = np.float32(5.991)
        return Thresholds(float(m), float(np.float32(1.5) * m), float(np.float32(7.815)))


CREATE_ASYNC_UPLOAD = 0x1


class CreateOptions(C.Structure):
    _fields_ = [("device", C.c_int32), ("rank", C.c_int32), ("nranks", C.c_int32), ("nccl_id", C.c_char_p), ("flags", C.c_uint32)]


class Extrinsics(C.Structure):
    """gpba_extrinsics (include/gpba.h)"""
    _fields_ = [("free_mask", C.POINTER(C.c_uint8)), ("prior_R", C.POINTER(C.c_double)), ("prior_info", C.POINTER(C.c_double))]


class StructureInfo(C.Structure):
    _fields_ = [("n_free_kf", C.c_int32), ("n_active_pt", C.c_int32), ("n_active_obs", C.c_int64),
                ("n_hpl", C.c_int64), ("n_hpp", C.c_int32), ("n_hschur", C.c_int32)]


def _arr(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


class Problem:
    """numpy owner of a gpba_problem. Field names = struct field names."""

    def __init__(self, **kw):
        self.cam_intr = _arr(kw["cam_intr"], np.float64).reshape(-1, 4)
        self.cam_Tbc = _arr(kw["cam_Tbc"], np.float64).reshape(-1, 7)
        self.bf = float(kw.get("bf", 0.0))
        self.kf_pose = _arr(kw["kf_pose"], np.float64).reshape(-1, 7)
        self.kf_vel = _arr(kw["kf_vel"], np.float64).reshape(-1, 6)
        self.kf_time = _arr(kw["kf_time"], np.float64)
        self.kf_fixed = _arr(kw["kf_fixed"], np.uint8)
        self.pt_xyz = _arr(kw["pt_xyz"], np.float64).reshape(-1, 3)
        self.rec_kf1 = _arr(kw["rec_kf1"], np.int32)
        self.rec_kf2 = _arr(kw["rec_kf2"], np.int32)
        self.rec_cam = _arr(kw["rec_cam"], np.int32)
        self.rec_t = _arr(kw["rec_t"], np.float64)
        self.obs_u = _arr(kw["obs_u"], np.float64)
        self.obs_v = _arr(kw["obs_v"], np.float64)
        ur = kw.get("obs_ur")
        self.obs_ur = None if ur is None else _arr(ur, np.float64)
        self.obs_inv_sigma2 = _arr(kw["obs_inv_sigma2"], np.float64)
        self.obs_rec = _arr(kw["obs_rec"], np.int32)
        self.obs_pt = _arr(kw["obs_pt"], np.int32)
        fl = kw.get("obs_flags")
        self.obs_flags = np.zeros(len(self.obs_u), np.uint8) if fl is None else _arr(fl, np.uint8)
        self.prior_kf1 = _arr(kw.get("prior_kf1", []), np.int32)
        self.prior_kf2 = _arr(kw.get("prior_kf2", []), np.int32)
        self.velp_kf = _arr(kw.get("velp_kf", []), np.int32)
        self.qc = _arr(kw["qc"], np.float64)
        self.huber_mono = float(kw.get("huber_mono", 0.0))
        self.huber_stereo = float(kw.get("huber_stereo", 0.0))
        self.huber_prior = float(kw.get("huber_prior", 0.0))
        self.lambda_init = float(kw.get("lambda_init", 0.0))
        self.linear_solver = int(kw.get("linear_solver", SOLVER_DENSE_CHOL))
        self.meta = dict(kw.get("meta", {}))
        self.truth = kw.get("truth")

    n_cam = property(lambda s: len(s.cam_intr))
    n_kf = property(lambda s: len(s.kf_pose))
    n_pt = property(lambda s: len(s.pt_xyz))
    n_rec = property(lambda s: len(s.rec_kf1))
    n_obs = property(lambda s: len(s.obs_u))

    def to_c(self):
        """ctypes struct borrowing this object's buffers (keep `self` alive while it is in use)."""
        def p(a, t):
            return a.ctypes.data_as(t) if a is not None and a.size else C.cast(None, t)
        c = CProblem()
        c.n_cam, c.cam_intr, c.cam_Tbc, c.bf = self.n_cam, p(self.cam_intr, _pd), p(self.cam_Tbc, _pd), self.bf
        c.n_kf, c.kf_pose, c.kf_vel = self.n_kf, p(self.kf_pose, _pd), p(self.kf_vel, _pd)
        c.kf_time, c.kf_fixed = p(self.kf_time, _pd), p(self.kf_fixed, _pb)
        c.n_pt, c.pt_xyz = self.n_pt, p(self.pt_xyz, _pd)
        c.n_rec, c.rec_kf1, c.rec_kf2 = self.n_rec, p(self.rec_kf1, _pi), p(self.rec_kf2, _pi)
        c.rec_cam, c.rec_t = p(self.rec_cam, _pi), p(self.rec_t, _pd)
        c.n_obs, c.obs_u, c.obs_v = self.n_obs, p(self.obs_u, _pd), p(self.obs_v, _pd)
        c.obs_ur = p(self.obs_ur, _pd)
        c.obs_inv_sigma2, c.obs_rec, c.obs_pt = p(self.obs_inv_sigma2, _pd), p(self.obs_rec, _pi), p(self.obs_pt, _pi)
        c.obs_flags = p(self.obs_flags, _pb)
        c.n_prior, c.prior_kf1, c.prior_kf2 = len(self.prior_kf1), p(self.prior_kf1, _pi), p(self.prior_kf2, _pi)
        c.n_velp, c.velp_kf = len(self.velp_kf), p(self.velp_kf, _pi)
        for i in range(6):
            c.qc[i] = float(self.qc[i])
        c.huber_mono, c.huber_stereo, c.huber_prior = self.huber_mono, self.huber_stereo, self.huber_prior
        c.lambda_init, c.linear_solver = self.lambda_init, self.linear_solver
        return c

    def input_bytes(self):
        """Host bytes uploaded by gpba_create (h2d per optimize call)."""
        arrs = [self.cam_intr, self.cam_Tbc, self.kf_pose, self.kf_vel, self.kf_time, self.kf_fixed, self.pt_xyz,
                self.rec_kf1, self.rec_kf2, self.rec_cam, self.rec_t, self.obs_u, self.obs_v, self.obs_inv_sigma2,
                self.obs_rec, self.obs_pt, self.obs_flags, self.prior_kf1, self.prior_kf2, self.velp_kf]
        if self.obs_ur is not None:
            arrs.append(self.obs_ur)
        return int(sum(a.nbytes for a in arrs))

    def subset_points(self, keep_pt_mask):
        """Problem restricted to a subset of points (all keyframes / records kept): the multi-GPU shard."""
        keep_pt_mask = np.asarray(keep_pt_mask, bool)
        new_idx = np.full(self.n_pt, -1, np.int32)
        new_idx[keep_pt_mask] = np.arange(int(keep_pt_mask.sum()), dtype=np.int32)
        om = keep_pt_mask[self.obs_pt]
        kw = dict(cam_intr=self.cam_intr, cam_Tbc=self.cam_Tbc, bf=self.bf, kf_pose=self.kf_pose, kf_vel=self.kf_vel,
                  kf_time=self.kf_time, kf_fixed=self.kf_fixed, pt_xyz=self.pt_xyz[keep_pt_mask],
                  rec_kf1=self.rec_kf1, rec_kf2=self.rec_kf2, rec_cam=self.rec_cam, rec_t=self.rec_t,
                  obs_u=self.obs_u[om], obs_v=self.obs_v[om],
                  obs_ur=None if self.obs_ur is None else self.obs_ur[om],
                  obs_inv_sigma2=self.obs_inv_sigma2[om], obs_rec=self.obs_rec[om], obs_pt=new_idx[self.obs_pt[om]],
                  obs_flags=self.obs_flags[om], prior_kf1=self.prior_kf1, prior_kf2=self.prior_kf2,
                  velp_kf=self.velp_kf, qc=self.qc, huber_mono=self.huber_mono, huber_stereo=self.huber_stereo,
                  huber_prior=self.huber_prior, lambda_init=self.lambda_init, linear_solver=self.linear_solver,
                  meta=self.meta)
        return Problem(**kw)
