"""Wall-clock phases of the end-to-end call sequence (create -> optimize -> download -> destroy), run under gpurun."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "amc-slam_b200"))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from pygpba import lib as G
import bench

name = sys.argv[1] if len(sys.argv) > 1 else "c4"
ASYNC = len(sys.argv) > 2 and sys.argv[2] == "async"
P = bench.load_problem(name)
cudart = torch.cuda.cudart()
for a in (P.obs_u, P.obs_v, P.obs_inv_sigma2, P.obs_rec, P.obs_pt, P.obs_flags, P.pt_xyz, P.kf_pose, P.kf_vel):
    cudart.cudaHostRegister(a.ctypes.data, a.nbytes, 0)
kp = np.zeros((P.n_kf, 7)); kv = np.zeros((P.n_kf, 6)); pt = np.zeros((P.n_pt, 3))
for rep in range(4):
    t0 = time.perf_counter(); g = G.GpBa(P, async_upload=ASYNC); t1a = time.perf_counter(); torch.cuda.synchronize()
    t1 = time.perf_counter(); g.build_structure(); torch.cuda.synchronize()
    t2 = time.perf_counter(); tr = g.optimize(10); torch.cuda.synchronize()
    t3 = time.perf_counter(); g.download_into(kp, kv, pt)
    t4 = time.perf_counter(); g.close(); torch.cuda.synchronize()
    t5 = time.perf_counter()
    print(f"rep{rep}: create {1e3*(t1a-t0):.1f}+sync {1e3*(t1-t1a):.1f}  structure {1e3*(t2-t1):.1f}  optimize {1e3*(t3-t2):.1f}  download {1e3*(t4-t3):.1f}  destroy {1e3*(t5-t4):.1f}  total {1e3*(t5-t0):.1f} ms", flush=True)
