"""Comparison of a finished run with the committed oracle fixtures at the BASELINE.json sizes
(tests/golden/baseline_<name>.npz, minted by tests/golden/make_golden_baseline.py from the CPU oracle).

Used by the `-m gpu` tests and by bench.py (`parity_check` in its JSON line, at every N, so the multi-GPU path is checked
in a driver-run artefact).  Nothing here loads the CPU checker: the fixtures are plain arrays.

Tolerances are BASELINE.json's north_star: identical iteration / trial counts, final cost 1e-6 relative, poses 1e-6 m /
1e-7 rad, outlier flags bit-exact outside the 1e-6 chi2 band; widened only by the fixture's own reproducibility band
(tolerance = max(north-star, 10 x band), see make_golden_baseline.py) -- on C2/C3/C4 the band is orders of magnitude
below the north-star tolerance, so the north-star numbers are the ones that apply.
"""
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
POS_TOL, ROT_TOL, COST_RTOL, LAMBDA_RTOL, VEL_TOL, PT_TOL = 1e-6, 1e-7, 1e-6, 1e-5, 1e-5, 1e-5
BAND_FACTOR = 10.0


def fixture_path(name):
    return os.path.join(ROOT, "tests", "golden", f"baseline_{name}.npz")


def load(name):
    return np.load(fixture_path(name))


def rot_angle(qa, qb):
    s = np.sign(np.sum(qa * qb, axis=1))[:, None]
    return 2 * np.arcsin(np.minimum(1.0, np.linalg.norm(qa * s - qb, axis=1) / 2))


def tolerances(F):
    return dict(pos=max(POS_TOL, BAND_FACTOR * float(F["band_pos_m"])), rot=max(ROT_TOL, BAND_FACTOR * float(F["band_rot_rad"])),
                cost=max(COST_RTOL, BAND_FACTOR * float(F["band_chi2_rel"])), vel=max(VEL_TOL, BAND_FACTOR * float(F["band_vel"])),
                pt=max(PT_TOL, BAND_FACTOR * float(F["band_pt_m"])))


def compare(F, traces, state=None, flags=None, edge_chi2=None):
    """traces: list of LmTrace summaries (one per round); state = (kf_pose, kf_vel, pt_xyz) full arrays.
    Returns a dict of deviations and `ok`."""
    tol = tolerances(F)
    out = {"tolerance": tol}
    n_it = [int(t["n_iters"]) for t in traces]
    out["iters"] = n_it
    out["iters_equal"] = n_it == [int(x) for x in F["tr_n_iters"]]
    out["trials_equal"] = out["iters_equal"] and all(list(t["trials"]) == [int(x) for x in F["tr_trials"][r, :n]] for r, (t, n) in enumerate(zip(traces, n_it)))
    cost_dev, lam_dev = 0.0, 0.0
    if out["iters_equal"]:
        for r, (t, n) in enumerate(zip(traces, n_it)):
            for f in ("chi2_before", "chi2_after"):
                a, b = np.asarray(t[f][:n]), F["tr_" + f][r, :n]
                cost_dev = max(cost_dev, float(np.max(np.abs(a - b) / np.abs(b))))
            a, b = np.asarray(t["lam"][:n]), F["tr_lam"][r, :n]
            lam_dev = max(lam_dev, float(np.max(np.abs(a - b) / np.abs(b))))
    else:
        cost_dev = lam_dev = float("inf")
    out["cost_rel"] = cost_dev
    out["lambda_rel"] = lam_dev
    ok = out["iters_equal"] and out["trials_equal"] and cost_dev <= tol["cost"] and lam_dev <= LAMBDA_RTOL
    if state is not None:
        kp, kv, pt = state
        out["pos_m"] = float(np.abs(kp[:, 4:] - F["kf_pose"][:, 4:]).max())
        out["rot_rad"] = float(rot_angle(kp[:, :4], F["kf_pose"][:, :4]).max())
        out["vel"] = float(np.abs(kv - F["kf_vel"]).max())
        out["pt_m"] = float(np.abs(pt[F["pt_idx"]] - F["pt_xyz"]).max())
        ok = ok and out["pos_m"] <= tol["pos"] and out["rot_rad"] <= tol["rot"] and out["vel"] <= tol["vel"] and out["pt_m"] <= tol["pt"]
    if flags is not None:
        ref = np.unpackbits(F["flags_packed"])[:int(F["n_obs"])].astype(np.uint8)
        # north-star exclusion: observations whose chi2 lies within 1e-6 of a threshold
        from .problem import Thresholds
        th = Thresholds.local_gpba()
        near_idx, near_chi2 = F["near_idx"], F["near_chi2"]
        excl = near_idx[(np.abs(near_chi2 - th.chi2_mono) < 1e-6) | (np.abs(near_chi2 - th.chi2_mono_close) < 1e-6)]
        mask = np.ones(len(ref), bool)
        mask[excl] = False
        out["flags_differ"] = int((np.asarray(flags)[mask] != ref[mask]).sum())
        out["flags_excluded"] = int(len(excl))
        out["n_flagged"] = int(np.asarray(flags).sum())
        ok = ok and out["flags_differ"] == 0
    if edge_chi2 is not None:
        a, b = np.asarray(edge_chi2)[F["obs_idx"]], F["edge_chi2"]
        out["edge_chi2_rel"] = float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-2)))
        ok = ok and out["edge_chi2_rel"] <= 1e-4
    out["ok"] = bool(ok)
    return out
