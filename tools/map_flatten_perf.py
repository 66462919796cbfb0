"""Host-side flattening throughput of the map mirror (csrc/gpba_map.cc): time of gpba_map_local_window /
gpba_map_global_window alone (no Python copies), beside the object-graph restatement (oracle/map_flatten.py) at C2."""
import sys, os, time, json, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "amc-slam_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from pygpba import mapmirror as MM, synth


def load(P, M):
    n_cam = P.n_cam
    cam_time = np.tile(P.kf_time[:, None], (1, n_cam)); cam_time[P.rec_kf2, P.rec_cam] = P.rec_t
    for k in range(P.n_kf):
        M.add_keyframe(k, k - 1, P.kf_pose[k], P.kf_vel[k], P.kf_time[k], cam_time[k])
    for j in range(P.n_pt):
        M.add_point(j, P.pt_xyz[j])
    kf2, cam = P.rec_kf2[P.obs_rec], P.rec_cam[P.obs_rec]
    o = np.argsort(kf2, kind="stable")
    if hasattr(M, "add_observations"):      # the product mirror: bulk hook; the restatement takes them one by one
        M.add_observations(kf2[o], cam[o], P.obs_pt[o], P.obs_u[o], P.obs_v[o], None, P.obs_inv_sigma2[o], P.obs_flags[o] & 1)
        return
    add = M.add_observation
    for i in o:
        add(kf2[i], cam[i], P.obs_pt[i], P.obs_u[i], P.obs_v[i], -1.0, P.obs_inv_sigma2[i], P.obs_flags[i] & 1)


def raw_time(M, fn, *args, reps=5):
    best = 1e9
    for _ in range(reps):
        h = C.c_void_p()
        t = time.perf_counter(); rc = fn(M.h, *args, C.byref(h)); dt = time.perf_counter() - t
        assert rc == 0
        n = M.L.gpba_window_problem(h).contents.n_obs
        M.L.gpba_window_destroy(h)
        best = min(best, dt)
    return best, n


out = {}
for name in sys.argv[1:] or ["c2", "c4"]:
    P = synth.make_problem(name)
    M = MM.MapMirror(P.cam_intr, P.cam_Tbc, P.bf, P.qc)
    t = time.time(); load(P, M); tl = time.time() - t
    L = M.L
    tg, ng = raw_time(M, L.gpba_map_global_window, C.c_int64(0))
    cov = np.zeros(0, np.int64)
    tw, nw = raw_time(M, L.gpba_map_local_window, C.c_int64(P.n_kf - 1), C.c_int32(0), None, C.c_int32(0))
    out[name] = dict(n_obs=int(P.n_obs), load_s=tl, global_ms=tg * 1e3, global_obs=int(ng), global_obs_per_s=ng / tg,
                     local_ms=tw * 1e3, local_obs=int(nw), local_obs_per_s=nw / tw)
    if name == "c2":
        import map_flatten
        R = map_flatten.RefMap(P.cam_intr, P.cam_Tbc, P.bf, P.qc)
        load(P, R)
        t = time.perf_counter(); r = R.global_window(0); out[name]["restatement_global_ms"] = (time.perf_counter() - t) * 1e3
        t = time.perf_counter(); r = R.local_window(P.n_kf - 1); out[name]["restatement_local_ms"] = (time.perf_counter() - t) * 1e3
    print(name, json.dumps(out[name]), flush=True)
