"""Essential-graph optimisation on the GPU vs the CPU oracle and the oracle's own reproducibility band (run under gpurun)."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "amc-slam_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import oracle_py as O
from pygpba import posegraph as PG
for n_kf, fs in ((60, True), (150, True), (150, False)):
    G = PG.make_pose_graph(n_kf=n_kf, seed=7, fix_scale=fs, scale_drift=0.0 if fs else 0.002)
    PG.optimize(G, 20)
    t = time.time(); og, tg = PG.optimize(G, 20); dt_g = time.time() - t
    t = time.time(); oc, tc = O.pose_graph_optimize(G, 20); dt_c = time.time() - t
    perm = np.random.default_rng(1).permutation(G.n_edge)
    G2 = PG.PoseGraph(G.sim3, G.fixed, G.edge_i[perm], G.edge_j[perm], G.edge_meas[perm], G.fix_scale)
    o2, t2 = O.pose_graph_optimize(G2, 20)
    a, b, c = tg.summary(), tc.summary(), t2.summary()
    print(f"n_kf {n_kf} fix_scale {fs} edges {G.n_edge}: gpu {dt_g*1e3:.1f} ms, oracle {dt_c*1e3:.1f} ms")
    print("  gpu    iters", a["n_iters"], a["trials"], ["%.8g" % x for x in a["chi2_after"]])
    print("  oracle iters", b["n_iters"], b["trials"], ["%.8g" % x for x in b["chi2_after"]])
    print("  shuffl iters", c["n_iters"], c["trials"], ["%.8g" % x for x in c["chi2_after"]])
    print("  chi2_before[0] rel dev %.2e" % (abs(a["chi2_before"][0] - b["chi2_before"][0]) / b["chi2_before"][0]))
    print("  gpu-oracle pose dev %.3e  oracle band %.3e" % (np.abs(og[:, 4:7] - oc[:, 4:7]).max(), np.abs(oc[:, 4:7] - o2[:, 4:7]).max()))

# a C4-sized essential graph on the device alone (the oracle's dense LDLT would take hours: it restates LinearSolverEigen by a dense solve)
G = PG.make_pose_graph(n_kf=1000, seed=7)
PG.optimize(G, 20)
t = time.time(); og, tg = PG.optimize(G, 20); dt = time.time() - t
a = tg.summary()
print(f"n_kf 1000 edges {G.n_edge}: gpu {dt*1e3:.1f} ms, iters {a['n_iters']} trials {a['trials']} chi2 {a['chi2_before'][0]:.6g} -> {a['chi2_after'][-1]:.6g}")
