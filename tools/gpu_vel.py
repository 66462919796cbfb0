"""Velocity RANSAC: gpba_vel_ransac (host buffers in/out) beside the CPU oracle on the same batch (run under gpurun)."""
import sys, os, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "amc-slam_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
from pygpba import velransac as VR
import oracle_py
out = {}
for n_hyp in (23, 148, 1184):
    B = VR.make_vel_batch(n_match=1400, n_hyp=n_hyp, A=2, outliers=0.2, seed=81)
    best = 1e9
    for rep in range(6):
        t = time.perf_counter(); R = VR.vel_ransac(B); dt = time.perf_counter() - t
        if rep: best = min(best, dt)
    t = time.perf_counter(); O = oracle_py.vel_ransac(B); dc = time.perf_counter() - t
    out[n_hyp] = dict(gpu_ms=best * 1e3, cpu_ms=dc * 1e3, hyp_per_s_gpu=n_hyp / best, hyp_per_s_cpu=n_hyp / dc, best_gpu=int(R.best.value), best_cpu=int(O.best.value),
                      inliers_best=int(R.inliers[R.best.value]))
    print(n_hyp, json.dumps(out[n_hyp]), flush=True)
