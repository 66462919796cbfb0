"""Stage breakdown of gpba_optimize on the BASELINE configs (run under gpurun)."""
import sys, os, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "amc-slam_b200"))
import numpy as np
from pygpba import synth, lib as G
from pygpba.problem import SOLVER_PCG, SOLVER_DENSE_CHOL


def run(name, solver=SOLVER_DENSE_CHOL, iters=10, reps=3, **kw):
    t = time.time(); P = synth.make_problem(name, **kw); P.linear_solver = solver
    print(f"===== {name} solver={solver} n_obs={P.n_obs} n_pt={P.n_pt} n_kf={P.n_kf} gen {time.time()-t:.1f}s", flush=True)
    t = time.time(); g = G.GpBa(P); t_create = time.time() - t
    t = time.time(); info = g.build_structure(); t_struct = time.time() - t
    print(f" create {t_create*1e3:.1f} ms  build_structure {t_struct*1e3:.1f} ms  free_kf {info.n_free_kf} hpl {info.n_hpl} hpp {info.n_hpp} hs {info.n_hschur}")
    for rep in range(reps):
        g.reset_state()
        g.set_profiling(rep == reps - 1)
        t = time.time(); tr = g.optimize(iters); dt = time.time() - t
        s = tr.summary()
        print(f" rep{rep}: optimize {dt*1e3:.2f} ms iters {s['n_iters']} trials {s['total_trials']} cg {s['cg_iterations']} chi2 {s['chi2_before'][0]:.6g}->{s['chi2_after'][s['n_iters']-1]:.6g}"
              f"  obs/s {P.n_obs*s['n_iters']/dt:.3e}  ms/iter {dt*1e3/s['n_iters']:.3f}")
    st = g.stage_stats(reset=True)
    tot = sum(v['ms'] for v in st.values())
    for k, v in st.items():
        print(f"   {k:16s} {v['ms']:9.3f} ms  {v['launches']:6d} launches  {100*v['ms']/max(tot,1e-9):5.1f}%")
    print(f"   sum of stages {tot:.3f} ms (profiled rep wall {dt*1e3:.2f} ms)")
    return g


def run_rejection_rounds(name="c3", reps=3):
    """C3: 4 x (initializeOptimization(0) + optimize(10)) with re-flagging after every round (Optimizer.cc:548-675 structure)."""
    P = synth.make_problem(name)
    print(f"===== {name} rejection rounds n_obs={P.n_obs} n_pt={P.n_pt} n_kf={P.n_kf} outliers={P.meta['outliers']}", flush=True)
    for rep in range(reps):
        g = G.GpBa(P)
        t = time.time(); flags, traces = g.rejection_rounds(4, 10); dt = time.time() - t
        its = sum(tr.n_iters for tr in traces)
        truth = P.truth["is_outlier"] if P.truth is not None else None
        msg = ""
        if truth is not None:
            msg = f" flagged {int(flags.sum())} (true outliers {int(truth.sum())}, recall {float((flags[truth] > 0).mean()):.3f}, false positives {float((flags[~truth] > 0).mean()):.4f})"
        print(f" rep{rep}: 4 rounds {dt*1e3:.2f} ms, {its} LM iterations, obs/s {P.n_obs*its/dt:.3e}{msg}", flush=True)
        g.close()


if __name__ == "__main__":
    which = sys.argv[1:] or ["c2", "c3", "c4"]
    for w in which:
        if w == "c4pcg":
            run("c4", SOLVER_PCG)
        elif w == "c3rr":
            run_rejection_rounds("c3")
        else:
            run(w)
