"""Landmark-sharded multi-GPU run vs the single-GPU run of the same problem (torchrun, one rank per GPU)."""
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
ok = True
CASES = [("loop", {}), ("c1", {}), ("tiny_global", dict(n_kf=400, n_pt=20000, obs_per_pt=8, seed=5)),
         ("c5", dict(n_kf=2000, n_pt=100000))]
if os.environ.get("GPBA_MULTI_CASES") == "sweep":
    CASES = [("tiny_global", dict(n_kf=400, n_pt=n, obs_per_pt=8, seed=5)) for n in (2000, 4000, 9000, 20000)]
for name, kw in CASES:
    P = synth.make_problem(name, **kw)
    g = gl.GpBa(P, device=local, rank=rank, nranks=world, nccl_id=nccl_id)
    tr = g.optimize(10).summary()
    kp, kv, pt = g.state()
    c2 = g.edge_chi2()
    if rank == 0:
        s = gl.GpBa(P, device=local)
        ts = s.optimize(10).summary()
        kp0, kv0, pt0 = s.state()
        same = ts["n_iters"] == tr["n_iters"] and ts["trials"] == tr["trials"]
        rel = np.abs(np.array(ts["chi2_after"][:min(ts["n_iters"], tr["n_iters"])]) / np.array(tr["chi2_after"][:min(ts["n_iters"], tr["n_iters"])]) - 1).max()
        dp = np.abs(kp - kp0).max() if same else float("nan")
        dc = np.abs(c2 - s.edge_chi2()).max() / max(np.abs(c2).max(), 1e-300)
        print(f"{name} {kw}: n_obs {P.n_obs} single {ts['n_iters']} iters trials {ts['trials']} | x{world} {tr['n_iters']} iters trials {tr['trials']} | "
              f"chi2 rel {rel:.2e} pose maxabs {dp:.2e} pts {np.abs(pt - pt0).max() if same else float('nan'):.2e} edge chi2 rel {dc:.2e}", flush=True)
        print("   single chi2_after", [f"{v:.10g}" for v in ts["chi2_after"][:4]], "\n   multi  chi2_after", [f"{v:.10g}" for v in tr["chi2_after"][:4]], flush=True)
        ok = ok and same
    dist.barrier()
if rank == 0:
    print("MULTI-GPU PARITY", "OK" if ok else "MISMATCH")
dist.destroy_process_group()
