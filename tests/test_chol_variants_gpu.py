"""The schedules of the reduced-system factorization agree with each other.

The factorization kernel reads its tuning switches from the environment once per process, so every variant runs in a
process of its own: the default (persistent dataflow kernel, diagonal factor published in wide levels only), the published
factor forced everywhere / nowhere (GPBA_CF_SPLIT_MIN), other chunk sizes, and the level-by-level launch sequence kept for
A/B comparisons (GPBA_CHOL_LEVELS).  Same seeded map, same LM trace to 1e-9 relative: they differ only in summation order."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r"""
import sys, os, json
sys.path.insert(0, os.path.join(%r, "amc-slam_b200"))
import numpy as np
from pygpba import synth, lib as G
P = synth.make_problem("tiny_global", n_kf=400, n_pt=12000, obs_per_pt=8, seed=41)
g = G.GpBa(P)
tr = g.optimize(4).summary()
kp, kv, pt = g.state()
st = g.solver_stats()
print(json.dumps({"chi2": list(map(float, tr["chi2_after"][: tr["n_iters"]])), "trials": list(map(int, tr["trials"][: tr["n_iters"]])),
                  "pos": kp[:, 4:].tolist(), "levels": int(st["levels"])}))
""" % ROOT

VARIANTS = {
    "default": {},
    "publish_everywhere": {"GPBA_CF_SPLIT_MIN": "0"},
    "publish_nowhere": {"GPBA_CF_SPLIT_MIN": "1000000"},
    "chunks_of_5": {"GPBA_LU_CHUNK": "5", "GPBA_LU_LATE_CHUNK": "0"},
    "level_by_level": {"GPBA_CHOL_LEVELS": "1"},
}


def run_variant(env_extra):
    env = dict(os.environ); env.update(env_extra)
    r = subprocess.run([sys.executable, "-c", SCRIPT], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    return json.loads(r.stdout.strip().splitlines()[-1])


def test_factorization_schedules_agree():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    ref = run_variant(VARIANTS["default"])
    assert ref["levels"] > 1 and len(ref["chi2"]) >= 2
    for name, env in VARIANTS.items():
        if name == "default":
            continue
        got = run_variant(env)
        assert got["trials"] == ref["trials"], name
        np.testing.assert_allclose(got["chi2"], ref["chi2"], rtol=1e-9, err_msg=name)
        assert np.abs(np.array(got["pos"]) - np.array(ref["pos"])).max() < 1e-8, name
