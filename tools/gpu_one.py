"""One optimize() of a config with a given iteration count (target of ncu captures; run under gpurun)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "amc-slam_b200"))
from pygpba import synth, lib as G
from pygpba.problem import SOLVER_PCG

name = sys.argv[1] if len(sys.argv) > 1 else "c2"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 2
P = synth.make_problem(name)
if len(sys.argv) > 3 and sys.argv[3] == "pcg":
    P.linear_solver = SOLVER_PCG
g = G.GpBa(P)
tr = g.optimize(iters).summary()
print(name, P.n_obs, tr["n_iters"], tr["total_trials"], tr["chi2_after"][: tr["n_iters"]])
