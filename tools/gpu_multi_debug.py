"""Compare the all-reduced reduced camera system of a 2-rank run with the single-GPU one (debug aid)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "amc-slam_b200"))
import numpy as np
import torch
import torch.distributed as dist
from pygpba import synth, lib as gl

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
if rank == 0:
    idt = torch.tensor(list(gl.nccl_unique_id()), dtype=torch.uint8, device="cuda")
dist.broadcast(idt, 0)
nccl_id = bytes(idt.cpu().tolist())
P = synth.make_problem("tiny_global", n_kf=400, n_pt=2000, obs_per_pt=8, seed=5)
g = gl.GpBa(P, device=local, rank=rank, nranks=world, nccl_id=nccl_id)
info = g.build_structure()
c = g.compute_errors()
g.build_system()
g.set_lambda(P.lambda_init)
okm = g.solve()
H, bs = g.hschur()
x = g.x()[:12 * info.n_free_kf]
if rank == 0:
    s = gl.GpBa(P, device=local)
    i0 = s.build_structure()
    c0 = s.compute_errors()
    s.build_system(); s.set_lambda(P.lambda_init); ok0 = s.solve()
    H0, bs0 = s.hschur()
    x0 = s.x()[:12 * i0.n_free_kf]
    print("chi2", c, c0, "ok", okm, ok0, "n_hs", info.n_hschur, i0.n_hschur, "n_lm", info.n_active_pt, i0.n_active_pt)
    d = np.abs(H - H0).reshape(len(H), -1).max(1)
    sc = np.abs(H0).max()
    bad = np.argsort(-d)[:8]
    r, cc = s.hschur_pattern()
    print("Hs max abs diff", d.max(), "scale", sc, "worst blocks", [(int(r[b]), int(cc[b]), float(d[b])) for b in bad])
    print("bs max diff", np.abs(bs - bs0).max(), "scale", np.abs(bs0).max(), "at", int(np.argmax(np.abs(bs - bs0))) // 12)
    print("x max diff", np.abs(x - x0).max(), "scale", np.abs(x0).max())
dist.barrier()
dist.destroy_process_group()
