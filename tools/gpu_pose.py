"""Pose-only GP optimisation: frames/s of gpba_pose_optimize (host buffers in, host buffers out) beside the CPU oracle
on a sample of the same frames (run under gpurun)."""
import sys, os, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "amc-slam_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
from pygpba import pose as PO

n_frames = int(sys.argv[1]) if len(sys.argv) > 1 else 256
n_pt = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
t = time.time(); B = PO.make_pose_batch(n_frames=n_frames, n_pt=n_pt, A=2, outliers=0.1, seed=91, fix_prev=True); tg = time.time() - t
print(f"frames {B.n_frames} matches {B.n_obs} ({B.n_obs / B.n_frames:.0f} per frame) generated in {tg:.1f} s", flush=True)
best = 1e9
for rep in range(5):
    t = time.time(); R = PO.pose_optimize(B); dt = time.time() - t
    best = min(best, dt)
    print(f" rep{rep}: {dt * 1e3:.2f} ms  {B.n_frames / dt:.0f} frames/s", flush=True)
its = sum(R.trace(f, r)["n_iters"] for f in range(B.n_frames) for r in range(4))
trials = sum(sum(R.trace(f, r)["trials"]) for f in range(B.n_frames) for r in range(4))
truth = B.truth_outlier
print(f" LM iterations {its} trials {trials}  recall {(R.outlier[truth] > 0).mean():.3f}  false positives {(R.outlier[~truth] > 0).mean():.3f}")
one = B.slice(0)
lat = 1e9
for rep in range(20):
    t = time.time(); PO.pose_optimize(one); lat = min(lat, time.time() - t)
print(f" single frame ({one.n_obs} matches): {lat * 1e3:.3f} ms per call (host buffers in/out)")
import oracle_py
ns = min(16, B.n_frames)
sub = [B.slice(f) for f in range(ns)]
t = time.time()
for s in sub:
    oracle_py.pose_optimize(s)
dc = time.time() - t
print(json.dumps({"metric": "pose_only_frames_per_s", "gpu_e2e": B.n_frames / best, "cpu_port_1core": ns / dc, "frames": B.n_frames,
                  "matches_per_frame": B.n_obs / B.n_frames, "gpu_ms": best * 1e3, "single_frame_ms": lat * 1e3, "cpu_sample_frames": ns, "cpu_ms_per_frame": dc * 1e3 / ns}))
