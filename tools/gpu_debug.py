"""Stage-by-stage GPU-vs-oracle diff printer (debug aid; run under gpurun)."""
import sys, os, time, traceback
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "amc-slam_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
from pygpba import synth, lib as G
from pygpba.problem import SOLVER_PCG, SOLVER_DENSE_CHOL
import oracle_py as O


def rel(a, b):
    a = np.asarray(a); b = np.asarray(b)
    if a.shape != b.shape:
        return f"SHAPE {a.shape} vs {b.shape}"
    d = np.abs(a - b).max() if a.size else 0.0
    s = np.abs(b).max() if b.size else 1.0
    return f"maxabs {d:.3e} (scale {s:.3e}, rel {d / (s + 1e-300):.2e})"


def run(name, solver=SOLVER_DENSE_CHOL, iters=10, **kw):
    print(f"===== {name} solver={solver} {kw}", flush=True)
    P = synth.make_problem(name, **kw); P.linear_solver = solver
    g = G.GpBa(P); o = O.Oracle(P)
    ig, io = g.build_structure(), o.build_structure()
    print(" info gpu", [getattr(ig, f[0]) for f in ig._fields_], "cpu", [getattr(io, f[0]) for f in io._fields_])
    print(" hpp pattern eq", all(np.array_equal(a, b) for a, b in zip(g.hpp_pattern(), o.hpp_pattern())),
          "hs pattern eq", all(np.array_equal(a, b) for a, b in zip(g.hschur_pattern(), o.hschur_pattern())))
    cg, co = g.compute_errors(), o.compute_errors()
    print(f" chi2 gpu {cg:.12g} cpu {co:.12g} rel {abs(cg - co) / co:.2e}")
    print(" edge chi2", rel(g.edge_chi2(), o.edge_chi2()))
    g.build_system(); o.build_system()
    print(" hpp", rel(g.hpp(), o.hpp())); print(" hll", rel(g.hll(), o.hll()))
    print(" hpl", rel(g.hpl()[2], o.hpl()[2])); print(" b", rel(g.b(), o.b()))
    lam = P.lambda_init
    g.set_lambda(lam); o.set_lambda(lam)
    okg, oko = g.solve(), o.solve()
    print(" solve ok", okg, oko)
    print(" hs", rel(g.hschur()[0], o.hschur()[0])); print(" bs", rel(g.hschur()[1], o.hschur()[1]))
    xg, xo = g.x(), o.x()
    npz = ig.n_free_kf * 12
    print(" x poses", rel(xg[:npz], xo[:npz])); print(" x lms", rel(xg[npz:], xo[npz:]))
    g2 = G.GpBa(P); o2 = O.Oracle(P)
    t = time.time(); tg = g2.optimize(iters); tgpu = time.time() - t
    t = time.time(); tc = o2.optimize(iters); tcpu = time.time() - t
    a, b = tg.summary(), tc.summary()
    print(f" optimize gpu {tgpu:.3f}s cpu {tcpu:.3f}s")
    print("  gpu iters", a["n_iters"], a["result"], a["trials"], "cg", a["cg_iterations"]); print("  cpu iters", b["n_iters"], b["result"], b["trials"])
    print("  chi2 gpu", [f"{c:.9g}" for c in a["chi2_after"]]); print("  chi2 cpu", [f"{c:.9g}" for c in b["chi2_after"]])
    print("  lam gpu", [f"{c:.4g}" for c in a["lam"]]); print("  lam cpu", [f"{c:.4g}" for c in b["lam"]])
    sg, sc = g2.state(), o2.state()
    print("  pose t", rel(sg[0][:, 4:], sc[0][:, 4:]), " q", rel(sg[0][:, :4], sc[0][:, :4])); print("  vel", rel(sg[1], sc[1])); print("  pts", rel(sg[2], sc[2]))
    print("  stored chi2", rel(g2.edge_chi2(), o2.edge_chi2()))


if __name__ == "__main__" and "rounds" not in sys.argv:
    cases = [("tiny", SOLVER_DENSE_CHOL), ("tiny_global", SOLVER_DENSE_CHOL), ("c1", SOLVER_DENSE_CHOL), ("loop", SOLVER_DENSE_CHOL),
             ("tiny", SOLVER_PCG), ("loop", SOLVER_PCG)]
    for name, solver in cases:
        try:
            run(name, solver)
        except Exception:
            traceback.print_exc()


def rounds_debug():
    from pygpba.problem import Thresholds
    P = synth.make_problem("c1", n_pt=600, outliers=0.3, seed=33)
    g = G.GpBa(P); o = O.Oracle(P)
    th = Thresholds.local_gpba()
    for it in range(4):
        tg, tc = g.optimize(10).summary(), o.optimize(10).summary()
        print(f"--- round {it}: iters {tg['n_iters']}/{tc['n_iters']} trials {tg['trials']} / {tc['trials']}")
        print("   chi2_before rel", np.abs(np.array(tg['chi2_before']) - np.array(tc['chi2_before'])) / np.array(tc['chi2_before']))
        print("   chi2_after rel", np.abs(np.array(tg['chi2_after']) - np.array(tc['chi2_after'])) / np.array(tc['chi2_after']))
        sg, sc = g.state(), o.state()
        print("   pose t", rel(sg[0][:, 4:], sc[0][:, 4:]), " pts", rel(sg[2], sc[2]))
        g.compute_errors_inactive(); o.compute_errors_inactive()
        cg, co = g.edge_chi2(), o.edge_chi2()
        print("   stored chi2", rel(cg, co), "worst idx", int(np.argmax(np.abs(cg - co))))
        fg, fo = g.outlier_flags(th), o.outlier_flags(th)
        print("   flags differ", int((fg != fo).sum()), "of", len(fo), "flagged", int(fo.sum()))
        g.set_levels(fg); o.set_levels(fo)
        if it == 2:
            g.set_robust_kernel(0); o.set_robust_kernel(0)


if __name__ == "__main__" and "rounds" in sys.argv:
    rounds_debug()
