"""Host-only symbolic phase of the reduced-system factorization (csrc/gpba_order.h through gpba_symbolic_analyze): the
nested-dissection order, the tile-level symbolic factorization and the level schedule.  It replaces what
SimplicialLDLT::analyzePattern + AMD do for the reference (linear_solver_eigen.h:147-201); no device is needed."""
import numpy as np
import pytest

from pygpba import lib as gl

BPT = 4   # pose blocks per 48 x 48 tile


def ring_pattern(n, lap, band):
    """upper block pattern of a trajectory that drives the same loop twice: |i - j| <= band or |i - j - lap| <= band"""
    r, c = [], []
    for i in range(n):
        for j in range(i, n):
            if j - i <= band or abs(j - i - lap) <= band:
                r.append(i); c.append(j)
    return np.array(r, np.int32), np.array(c, np.int32)


def tile_fill(n, r, c, perm):
    """independent boolean right-looking elimination on the tile graph -> (#tiles, #levels)"""
    NT = (int(perm.max()) // BPT) + 1
    nz = np.zeros((NT, NT), bool)
    nz[np.arange(NT), np.arange(NT)] = True
    ti, tj = perm[c] // BPT, perm[r] // BPT
    lo, hi = np.minimum(ti, tj), np.maximum(ti, tj)
    nz[hi, lo] = True
    level = np.zeros(NT, int)
    for k in range(NT):
        rows = np.nonzero(nz[k + 1:, k])[0] + k + 1
        for a in rows:
            nz[a, rows[rows <= a]] = True
        level[rows] = np.maximum(level[rows], level[k] + 1)
    return int(np.tril(nz).sum()), int(level.max()) + 1


@pytest.mark.parametrize("depth", [0, 1, 3, -1])
def test_order_is_a_valid_tile_aligned_permutation(depth):
    n, (r, c) = 400, ring_pattern(400, 200, 12)
    perm, st = gl.symbolic_analyze(n, r, c, depth)
    assert len(set(perm.tolist())) == n and perm.min() >= 0            # injective into the padded position space
    assert st["tile_columns"] * BPT > perm.max()
    tiles, levels = tile_fill(n, r, c, perm)                            # the factor pattern and the schedule re-derived in numpy
    assert tiles == st["tiles"] and levels == st["levels"]
    if depth == 0:
        assert st["parts"] == 1 and st["levels"] == st["tile_columns"] == (n + BPT - 1) // BPT


def test_dissection_shortens_the_chain_of_a_loop_closed_trajectory():
    n, (r, c) = 999, ring_pattern(999, 500, 20)
    _, banded = gl.symbolic_analyze(n, r, c, 0)
    _, nd = gl.symbolic_analyze(n, r, c, -1)
    assert banded["levels"] == banded["tile_columns"] == 250
    assert nd["levels"] * 2.5 < banded["levels"]                        # the dependency chain is what the factorization waits for
    assert nd["update_pairs"] < 1.6 * banded["update_pairs"]            # ... bought with a bounded amount of extra work


def test_small_windows_keep_the_banded_order_and_bad_input_is_refused():
    n, (r, c) = 30, ring_pattern(30, 1000, 29)                          # a dense local window
    perm, st = gl.symbolic_analyze(n, r, c, -1)
    assert st["parts"] == 1 and st["tiles"] == 36 and sorted(perm.tolist()) == list(range(n))
    with pytest.raises(gl.GpbaError):
        gl.symbolic_analyze(4, np.array([2], np.int32), np.array([1], np.int32))     # lower-triangular entry
    with pytest.raises(gl.GpbaError):
        gl.symbolic_analyze(4, np.array([0], np.int32), np.array([7], np.int32))     # out of range
    perm, st = gl.symbolic_analyze(0, np.zeros(0, np.int32), np.zeros(0, np.int32))   # empty system
    assert st["tile_columns"] == 1 and st["tiles"] == 1


# ---- the task list of the persistent factorization kernel (k_chol_factor): its in-kernel waits are deadlock-free only if every
# task depends on tasks in FRONT of it in the list (they are claimed in list order by resident CTAs).  Checked on the host.
SCHEDULE_ENVS = [{}, {"GPBA_CF_SPLIT_MIN": "0"}, {"GPBA_CF_SPLIT_MIN": "1000000"}, {"GPBA_LU_CHUNK": "5", "GPBA_LU_LATE_CHUNK": "0"},
                 {"GPBA_LU_CHUNK": "1"}, {"GPBA_CHOL_ND_DEPTH": "0"}]


def _schedule_in_subprocess(env_extra, n, lap, band):
    """the chunking / publishing switches are read once per process"""
    import json, os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("import sys, json; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
            "from test_symbolic import ring_pattern\nfrom pygpba import lib as gl\n"
            "r, c = ring_pattern(%d, %d, %d)\nprint(json.dumps(gl.factor_schedule_check(%d, r, c)))\n"
            % (os.path.join(root, "amc-slam_b200"), os.path.join(root, "tests"), n, lap, band, n))
    env = dict(os.environ); env.update(env_extra)
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-1500:]
    return json.loads(out.stdout.strip().splitlines()[-1])


@pytest.mark.parametrize("env", SCHEDULE_ENVS)
def test_factor_task_list_only_depends_on_earlier_tasks(env):
    st = _schedule_in_subprocess(env, 999, 500, 20)
    assert st["violations"] == 0, st
    assert st["tasks"] == st["chunks"] + st["panel_tasks"] and st["products"] >= st["chunks"] > 0
    _, sym = gl.symbolic_analyze(999, *ring_pattern(999, 500, 20), int(env.get("GPBA_CHOL_ND_DEPTH", -1)))
    assert st["panel_tasks"] == sym["tiles"]                            # one panel task per non-zero tile of the factor


def test_factor_task_list_small_and_empty_systems():
    for n, lap, band in ((30, 1000, 29), (5, 1000, 1), (1, 1000, 0)):
        r, c = ring_pattern(n, lap, band)
        st = gl.factor_schedule_check(n, r, c)
        assert st["violations"] == 0 and st["panel_tasks"] >= 1, (n, st)
    st = gl.factor_schedule_check(0, np.zeros(0, np.int32), np.zeros(0, np.int32))
    assert st["violations"] == 0
    with pytest.raises(gl.GpbaError):
        gl.factor_schedule_check(4, np.array([2], np.int32), np.array([1], np.int32))
