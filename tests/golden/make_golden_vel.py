"""Mints tests/golden/vel_*.npz from the CPU oracle of the velocity RANSAC (oracle/vel_ransac.h).  No fixtures exist
for Optimizer::OptimizeVel; these files freeze the oracle, whose EdgeVelReproj is pinned against the reference's
own G2oTypes.cc (tests/test_ref_pin.py) while the RANSAC loop around it is a restatement.

    python tests/golden/make_golden_vel.py
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "amc-slam_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

CASES = {"basic": dict(seed=91, n_match=500, n_hyp=23), "wide": dict(seed=92, A=4, n_match=900, n_hyp=40, outliers=0.35)}


def input_checksum(B):
    h = hashlib.sha256()
    for f in ("cam_intr", "cam_Tbc", "cam_dt", "last_pose", "vel_init", "obs_u", "obs_v", "obs_inv_sigma2", "obs_xw", "obs_cam", "samples"):
        h.update(np.ascontiguousarray(getattr(B, f)).tobytes())
    return h.hexdigest()


def pack(B, R):
    return dict(vel=R.vel.copy(), inliers=R.inliers.copy(), mask=R.mask.copy(), best=np.int32(R.best.value),
                n_iters=np.array([R.trace(h)["n_iters"] for h in range(B.n_hyp)], np.int32))


if __name__ == "__main__":
    import oracle_py
    from pygpba.velransac import make_vel_batch
    for key, args in CASES.items():
        B = make_vel_batch(**args)
        out = pack(B, oracle_py.vel_ransac(B))
        out["input_sha256"] = np.array(input_checksum(B))
        np.savez_compressed(os.path.join(HERE, "vel_" + key + ".npz"), **out)
        print(key, "best", int(out["best"]), "inliers", out["inliers"])
