"""Mints tests/golden/pose_*.npz from the CPU oracle of the pose-only GP optimisation (oracle/pose_only.h).

The reference holds no fixtures for Optimizer::PoseGPOptimizationFromeLastFrame either: the function as a whole is a restatement (its edges are pinned
against the reference's own G2oTypes.cc, tests/test_ref_pin.py) and is not checked
against the reference binary; these files freeze the oracle's outputs on seeded frames.

    python tests/golden/make_golden_pose.py
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
    "fixed": dict(n_frames=3, n_pt=300, A=2, outliers=0.15, seed=61, fix_prev=True),
    "free_stereo": dict(n_frames=2, n_pt=300, A=2, outliers=0.1, seed=62, fix_prev=False, stereo_fraction=0.5),
}


def input_checksum(B):
    h = hashlib.sha256()
    for f in B.FIELDS:
        a = getattr(B, f)
        if a is not None:
            h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def pack(B, R):
    from pygpba.pose import GPBA_POSE_ROUNDS
    n_iters, trials, before, after = [], [], [], []
    for f in range(B.n_frames):
        for rnd in range(GPBA_POSE_ROUNDS):
            t = R.trace(f, rnd)
            n_iters.append(t["n_iters"]); trials += list(t["trials"]); before += list(t["chi2_before"]); after += list(t["chi2_after"])
    return dict(cur_pose=R.cur_pose.copy(), cur_vel=R.cur_vel.copy(), prev_pose=R.prev_pose.copy(), prev_vel=R.prev_vel.copy(),
                outlier=R.outlier.copy(), n_inliers=R.n_inliers.copy(), n_iters=np.array(n_iters, np.int32),
                trials=np.array(trials, np.int32), chi2_before=np.array(before), chi2_after=np.array(after))


if __name__ == "__main__":
    import oracle_py
    from pygpba.pose import make_pose_batch
    for key, args in CASES.items():
        B = make_pose_batch(**args)
        out = pack(B, oracle_py.pose_optimize(B))
        out["input_sha256"] = np.array(input_checksum(B))
        np.savez_compressed(os.path.join(HERE, "pose_" + key + ".npz"), **out)
        print(key, "n_obs", B.n_obs, "inliers", out["n_inliers"], "iters", out["n_iters"])
