"""Mints the golden vectors under tests/golden/ from the CPU oracle (oracle/gpba_oracle.cc).

The reference holds no golden vectors, known-answer tests or fixtures for this path and cannot be built here
(SURVEY.md 0.5 / 0.6), so these files are not reference outputs: they freeze the oracle's own outputs on seeded inputs
(the oracle itself is pinned per layer against the reference's own sources compiled into oracle/_ref (tests/test_ref_pin.py)) so that (a) the oracle cannot drift silently (CPU test) and (b) the CUDA path is checked
against committed numbers and not only against a freshly compiled checker (GPU test).

    python tests/golden/make_golden.py          # rewrites tests/golden/*.npz

Each file stores the generator arguments (the inputs are re-generated from the seed and checked by a checksum),
the structure sizes and block patterns, the robust chi2 at the start, the LM trace of optimize(10), the final
keyframe / landmark state and the stored edge chi2.
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "amc-slam_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

CASES = {
    "tiny_local": dict(name="tiny"),
    "tiny_global": dict(name="tiny_global"),
    "loop_global": dict(name="loop", n_pt=600),
    "c1_outliers": dict(name="c1", n_pt=500, outliers=0.2, seed=31),
}


def input_checksum(P):
    h = hashlib.sha256()
    for f in ("cam_intr", "cam_Tbc", "kf_pose", "kf_vel", "kf_time", "kf_fixed", "pt_xyz", "rec_kf1", "rec_kf2", "rec_cam",
              "rec_t", "obs_u", "obs_v", "obs_inv_sigma2", "obs_rec", "obs_pt", "obs_flags", "prior_kf1", "prior_kf2", "velp_kf"):
        h.update(np.ascontiguousarray(getattr(P, f)).tobytes())
    return h.hexdigest()


def make_case(args):
    from pygpba import synth
    a = dict(args)
    return synth.make_problem(a.pop("name"), **a)


def run_oracle(P):
    import oracle_py
    from pygpba.problem import Thresholds
    o = oracle_py.Oracle(P)
    info = o.build_structure()
    hpp_r, hpp_c = o.hpp_pattern()
    hs_r, hs_c = o.hschur_pattern()
    chi0 = o.compute_errors()
    o2 = oracle_py.Oracle(P)
    tr = o2.optimize(10).summary()
    kp, kv, pt = o2.state()
    flags = o2.outlier_flags(Thresholds.local_gpba())
    return dict(
        sizes=np.array([info.n_free_kf, info.n_active_pt, info.n_active_obs, info.n_hpl, info.n_hpp, info.n_hschur], np.int64),
        hpp_rows=hpp_r, hpp_cols=hpp_c, hs_rows=hs_r, hs_cols=hs_c, chi2_start=np.float64(chi0),
        n_iters=np.int32(tr["n_iters"]), result=np.int32(tr["result"]), trials=np.array(tr["trials"], np.int32),
        chi2_before=np.array(tr["chi2_before"]), chi2_after=np.array(tr["chi2_after"]), lam=np.array(tr["lam"]),
        kf_pose=kp, kf_vel=kv, pt_xyz=pt, edge_chi2=o2.edge_chi2(), flags=flags)


if __name__ == "__main__":
    for key, args in CASES.items():
        P = make_case(args)
        out = run_oracle(P)
        out["input_sha256"] = np.array(input_checksum(P))
        np.savez_compressed(os.path.join(HERE, key + ".npz"), **out)
        print(key, "n_obs", P.n_obs, "iters", int(out["n_iters"]), "chi2", float(out["chi2_start"]), "->", float(out["chi2_after"][-1]))
