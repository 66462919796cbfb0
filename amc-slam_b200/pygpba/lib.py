"""ctypes binding of libgpba.so (include/gpba.h).  Test / bench harness only: the product is the C ABI.

Fails loudly when the CUDA library is missing or no device is present -- there is no CPU fallback.
"""
import ctypes as C
import os

import numpy as np

from .problem import (CProblem, LmParams, LmTrace, StructureInfo, Thresholds, GPBA_N_STAGES)

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "libgpba.so")
_LIB = None

# every symbol include/gpba.h declares
SYMBOLS = [
    "gpba_create", "gpba_destroy", "gpba_last_error", "gpba_default_lm_params", "gpba_nccl_unique_id", "gpba_create_dist", "gpba_create_ex",
    "gpba_build_structure", "gpba_get_hpp_pattern", "gpba_get_hschur_pattern", "gpba_compute_errors", "gpba_build_system",
    "gpba_set_lambda", "gpba_restore_diagonal", "gpba_solve", "gpba_vector_size", "gpba_get_x", "gpba_get_b", "gpba_get_hpp",
    "gpba_get_hschur", "gpba_get_hll", "gpba_get_hpl", "gpba_oplus", "gpba_push", "gpba_pop", "gpba_discard_top",
    "gpba_optimize", "gpba_download_state", "gpba_edge_chi2", "gpba_edge_errors", "gpba_download_evaluated_state", "gpba_active_robust_chi2", "gpba_outlier_flags",
    "gpba_set_levels", "gpba_set_robust_kernel", "gpba_compute_errors_inactive", "gpba_rejection_rounds",
    "gpba_set_extrinsics", "gpba_get_extrinsics", "gpba_count_camera_observations", "gpba_calibrate_extrinsics",
    "gpba_stage_stats", "gpba_set_profiling", "gpba_reset_state", "gpba_get_stream", "gpba_schur_stats", "gpba_solver_stats", "gpba_symbolic_analyze", "gpba_factor_schedule_check", "gpba_pose_optimize", "gpba_vel_ransac", "gpba_pose_graph_optimize", "gpba_correct_points",
]


STAGE_NAMES = ["records", "residuals", "lin_landmarks", "lin_poses", "schur_prepare", "schur_pairs", "factorize", "tri_solve",
               "backsub_update", "collective", "schur_expand"]


class GpbaError(RuntimeError):
    pass


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise GpbaError(f"{LIB_PATH} is missing: build it with `python amc-slam_b200/build.py` (no CPU fallback exists)")
        L = C.CDLL(LIB_PATH)
        L.gpba_last_error.restype = C.c_char_p
        L.gpba_get_stream.restype = C.c_void_p
        L.gpba_pose_optimize.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 7
        L.gpba_vel_ransac.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 5
        L.gpba_create.argtypes = [C.POINTER(CProblem), C.c_int, C.POINTER(C.c_void_p)]
        L.gpba_create_dist.argtypes = [C.POINTER(CProblem), C.c_int, C.c_int, C.c_int, C.c_char_p, C.POINTER(C.c_void_p)]
        _LIB = L
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def default_lm_params():
    p = LmParams()
    lib().gpba_default_lm_params(C.byref(p))
    return p


def symbolic_analyze(n_pose, hs_row, hs_col, nd_depth=-1):
    """Host-only symbolic phase of the reduced-system factorization (no device needed): (perm, stats dict)."""
    r = np.ascontiguousarray(hs_row, np.int32); c = np.ascontiguousarray(hs_col, np.int32)
    perm = np.zeros(n_pose, np.int32); out = (C.c_int64 * 5)()
    rc = lib().gpba_symbolic_analyze(C.c_int32(n_pose), C.c_int32(len(r)), _p(r), _p(c), C.c_int32(nd_depth), _p(perm), out)
    if rc != 0:
        raise GpbaError(f"gpba_symbolic_analyze: {lib().gpba_last_error().decode()}")
    return perm, dict(tile_columns=out[0], levels=out[1], parts=out[2], tiles=out[3], update_pairs=out[4])


def factor_schedule_check(n_pose, hs_row, hs_col):
    """Host-only: task list of the persistent factorization kernel for a block pattern + its dependency invariants.
    Returns dict(tasks, chunks, products, panel_tasks, violations)."""
    r = np.ascontiguousarray(hs_row, np.int32); c = np.ascontiguousarray(hs_col, np.int32)
    out = (C.c_int64 * 5)()
    rc = lib().gpba_factor_schedule_check(C.c_int32(n_pose), C.c_int32(len(r)), _p(r), _p(c), out)
    if rc != 0:
        raise GpbaError(f"gpba_factor_schedule_check: {lib().gpba_last_error().decode()}")
    return dict(tasks=out[0], chunks=out[1], products=out[2], panel_tasks=out[3], violations=out[4])


def nccl_unique_id():
    buf = C.create_string_buffer(128)
    rc = lib().gpba_nccl_unique_id(buf)
    if rc != 0:
        raise GpbaError(f"gpba_nccl_unique_id: {lib().gpba_last_error().decode()}")
    return buf.raw


class GpBa:
    """One handle == one g2o::SparseOptimizer + BlockSolverX + OptimizationAlgorithmLevenberg instance."""

    def __init__(self, prob, device=-1, rank=0, nranks=1, nccl_id=None, async_upload=False):
        self.prob = prob      # keeps the host arrays alive (required until build_structure with async_upload)
        self._c = prob.to_c()
        self.L = lib()
        self.h = C.c_void_p()
        if async_upload:
            from .problem import CreateOptions, CREATE_ASYNC_UPLOAD
            o = CreateOptions(int(device), int(rank), int(nranks), nccl_id if nranks > 1 else None, CREATE_ASYNC_UPLOAD)
            rc = self.L.gpba_create_ex(C.byref(self._c), C.byref(o), C.byref(self.h))
        elif nranks > 1:
            rc = self.L.gpba_create_dist(C.byref(self._c), device, rank, nranks, nccl_id, C.byref(self.h))
        else:
            rc = self.L.gpba_create(C.byref(self._c), device, C.byref(self.h))
        self._ck(rc, "gpba_create")
        self.info = None

    def _ck(self, rc, what):
        if rc != 0:
            raise GpbaError(f"{what} failed ({rc}): {self.L.gpba_last_error().decode()}")

    def close(self):
        if getattr(self, "h", None):
            self.L.gpba_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- L1
    def build_structure(self):
        info = StructureInfo()
        self._ck(self.L.gpba_build_structure(self.h, C.byref(info)), "gpba_build_structure")
        self.info = info
        return info

    def _pattern(self, fn, n, what):
        r = np.zeros(n, np.int32); c = np.zeros(n, np.int32)
        self._ck(fn(self.h, _p(r), _p(c)), what)
        return r, c

    def hpp_pattern(self):
        return self._pattern(self.L.gpba_get_hpp_pattern, self.info.n_hpp, "gpba_get_hpp_pattern")

    def hschur_pattern(self):
        return self._pattern(self.L.gpba_get_hschur_pattern, self.info.n_hschur, "gpba_get_hschur_pattern")

    def compute_errors(self):
        chi = C.c_double()
        self._ck(self.L.gpba_compute_errors(self.h, C.byref(chi)), "gpba_compute_errors")
        return chi.value

    def build_system(self):
        self._ck(self.L.gpba_build_system(self.h), "gpba_build_system")

    def set_lambda(self, lam, backup=True):
        self._ck(self.L.gpba_set_lambda(self.h, C.c_double(lam), int(backup)), "gpba_set_lambda")

    def restore_diagonal(self):
        self._ck(self.L.gpba_restore_diagonal(self.h), "gpba_restore_diagonal")

    def solve(self):
        ok = C.c_int()
        self._ck(self.L.gpba_solve(self.h, C.byref(ok)), "gpba_solve")
        return bool(ok.value)

    def vector_size(self):
        n = C.c_int64()
        self._ck(self.L.gpba_vector_size(self.h, C.byref(n)), "gpba_vector_size")
        return n.value

    def x(self):
        a = np.zeros(self.vector_size()); self._ck(self.L.gpba_get_x(self.h, _p(a)), "gpba_get_x"); return a

    def b(self):
        a = np.zeros(self.vector_size()); self._ck(self.L.gpba_get_b(self.h, _p(a)), "gpba_get_b"); return a

    def hpp(self):
        a = np.zeros((self.info.n_hpp, 12, 12)); self._ck(self.L.gpba_get_hpp(self.h, _p(a)), "gpba_get_hpp"); return a

    def hschur(self):
        a = np.zeros((self.info.n_hschur, 12, 12)); bs = np.zeros(self.info.n_free_kf * 12)
        self._ck(self.L.gpba_get_hschur(self.h, _p(a), _p(bs)), "gpba_get_hschur")
        return a, bs

    def hll(self):
        a = np.zeros((self.info.n_active_pt, 3, 3)); self._ck(self.L.gpba_get_hll(self.h, _p(a)), "gpba_get_hll"); return a

    def hpl(self):
        beg = np.zeros(self.info.n_active_pt + 1, np.int64); pose = np.zeros(self.info.n_hpl, np.int32)
        blk = np.zeros((self.info.n_hpl, 12, 3))
        self._ck(self.L.gpba_get_hpl(self.h, _p(beg), _p(pose), _p(blk)), "gpba_get_hpl")
        return beg, pose, blk

    def oplus(self, x=None):
        self._ck(self.L.gpba_oplus(self.h, None if x is None else _p(np.ascontiguousarray(x, np.float64))), "gpba_oplus")

    def push(self):
        self._ck(self.L.gpba_push(self.h), "gpba_push")

    def pop(self):
        self._ck(self.L.gpba_pop(self.h), "gpba_pop")

    def discard_top(self):
        self._ck(self.L.gpba_discard_top(self.h), "gpba_discard_top")

    # ---- L2
    def optimize(self, iters=10, params=None, stop_flag=None):
        tr = LmTrace()
        self._ck(self.L.gpba_optimize(self.h, int(iters), stop_flag, C.byref(params) if params is not None else None,
                                      C.byref(tr)), "gpba_optimize")
        return tr

    def state(self):
        P = self.prob
        kp = np.zeros((P.n_kf, 7)); kv = np.zeros((P.n_kf, 6)); pt = np.zeros((P.n_pt, 3))
        self._ck(self.L.gpba_download_state(self.h, _p(kp), _p(kv), _p(pt)), "gpba_download_state")
        return kp, kv, pt

    def download_into(self, kp, kv, pt):
        self._ck(self.L.gpba_download_state(self.h, _p(kp), _p(kv), _p(pt)), "gpba_download_state")

    def edge_chi2(self):
        a = np.zeros(self.prob.n_obs); self._ck(self.L.gpba_edge_chi2(self.h, _p(a)), "gpba_edge_chi2"); return a

    def edge_errors(self):
        a = np.zeros((self.prob.n_obs, 3)); self._ck(self.L.gpba_edge_errors(self.h, _p(a)), "gpba_edge_errors"); return a

    def evaluated_state(self):
        P = self.prob
        kp = np.zeros((P.n_kf, 7)); kv = np.zeros((P.n_kf, 6)); tb = np.zeros((P.n_cam, 7))
        self._ck(self.L.gpba_download_evaluated_state(self.h, _p(kp), _p(kv), _p(tb)), "gpba_download_evaluated_state")
        return kp, kv, tb

    def active_robust_chi2(self):
        c = C.c_double(); self._ck(self.L.gpba_active_robust_chi2(self.h, C.byref(c)), "gpba_active_robust_chi2"); return c.value

    def outlier_flags(self, th=None):
        th = th or Thresholds.local_gpba()
        f = np.zeros(self.prob.n_obs, np.uint8)
        self._ck(self.L.gpba_outlier_flags(self.h, C.byref(th), _p(f)), "gpba_outlier_flags")
        return f

    def set_levels(self, level):
        self._ck(self.L.gpba_set_levels(self.h, _p(np.ascontiguousarray(level, np.uint8))), "gpba_set_levels")

    def set_robust_kernel(self, enabled):
        self._ck(self.L.gpba_set_robust_kernel(self.h, int(enabled)), "gpba_set_robust_kernel")

    def compute_errors_inactive(self):
        self._ck(self.L.gpba_compute_errors_inactive(self.h), "gpba_compute_errors_inactive")

    def rejection_rounds(self, n_rounds=4, iters=10, th=None, params=None):
        th = th or Thresholds.local_gpba()
        f = np.zeros(self.prob.n_obs, np.uint8)
        traces = (LmTrace * n_rounds)()
        self._ck(self.L.gpba_rejection_rounds(self.h, n_rounds, iters, C.byref(th),
                                              C.byref(params) if params is not None else None, _p(f), traces),
                 "gpba_rejection_rounds")
        return f, list(traces)

    # ---- extrinsic self-calibration
    def _ext_struct(self, free, prior_q, prior_info):
        from .problem import Extrinsics
        self._ext_keep = (np.ascontiguousarray(free, np.uint8),
                          None if prior_q is None else np.ascontiguousarray(prior_q, np.float64),
                          None if prior_info is None else np.ascontiguousarray(prior_info, np.float64))
        f, q, w = self._ext_keep
        return Extrinsics(f.ctypes.data_as(C.POINTER(C.c_uint8)),
                          None if q is None else q.ctypes.data_as(C.POINTER(C.c_double)),
                          None if w is None else w.ctypes.data_as(C.POINTER(C.c_double)))

    def set_extrinsics(self, free, prior_q=None, prior_info=None):
        e = self._ext_struct(free, prior_q, prior_info)
        self._ck(self.L.gpba_set_extrinsics(self.h, C.byref(e)), "gpba_set_extrinsics")

    def extrinsics(self):
        a = np.zeros((self.prob.n_cam, 7)); self._ck(self.L.gpba_get_extrinsics(self.h, _p(a)), "gpba_get_extrinsics"); return a

    def count_camera_observations(self):
        a = np.zeros(self.prob.n_cam, np.int64)
        self._ck(self.L.gpba_count_camera_observations(self.h, _p(a)), "gpba_count_camera_observations")
        return a

    def calibrate_extrinsics(self, candidates, prior_q=None, prior_info=None, min_obs=50, iters=10, params=None):
        e = self._ext_struct(candidates, prior_q, prior_info)
        tr = LmTrace()
        freed = np.zeros(self.prob.n_cam, np.uint8)
        self._ck(self.L.gpba_calibrate_extrinsics(self.h, C.byref(e), int(min_obs), int(iters), C.byref(params) if params is not None else None,
                                                  C.byref(tr), _p(freed)), "gpba_calibrate_extrinsics")
        return tr, freed

    # ---- measurement
    def set_profiling(self, on):
        self._ck(self.L.gpba_set_profiling(self.h, int(on)), "gpba_set_profiling")

    def stage_stats(self, reset=False):
        ms = (C.c_double * GPBA_N_STAGES)(); n = (C.c_int64 * GPBA_N_STAGES)()
        self._ck(self.L.gpba_stage_stats(self.h, ms, n, int(reset)), "gpba_stage_stats")
        names = STAGE_NAMES
        return {k: dict(ms=ms[i], launches=n[i]) for i, k in enumerate(names)}

    def schur_stats(self):
        a = (C.c_int64 * 4)()
        self._ck(self.L.gpba_schur_stats(self.h, a), "gpba_schur_stats")
        return dict(n_obs_pairs=a[0], n_record_pairs=a[1], n_items=a[2], n_contrib=a[3])

    def solver_stats(self):
        a = (C.c_int64 * 6)()
        self._ck(self.L.gpba_solver_stats(self.h, a), "gpba_solver_stats")
        return dict(tile_columns=a[0], levels=a[1], parts=a[2], tiles=a[3], tile_products=a[4], update_ctas=a[5])

    def stream(self):
        return self.L.gpba_get_stream(self.h)

    def reset_state(self, kf_pose=None, kf_vel=None, pt_xyz=None):
        P = self.prob
        kf_pose = P.kf_pose if kf_pose is None else kf_pose
        kf_vel = P.kf_vel if kf_vel is None else kf_vel
        pt_xyz = P.pt_xyz if pt_xyz is None else pt_xyz
        self._ck(self.L.gpba_reset_state(self.h, _p(np.ascontiguousarray(kf_pose)), _p(np.ascontiguousarray(kf_vel)),
                                         _p(np.ascontiguousarray(pt_xyz))), "gpba_reset_state")
