"""N > 1 host logic on CPU (gloo, world_size 2): the landmark-sharded reduced camera system.

SURVEY.md §8e: landmarks (with all their observations) are partitioned across ranks; every rank assembles a partial
Hschur / bschur over the replicated block pattern, priors and the lambda damping of Hpp enter on rank 0 only, and one
all-reduce (sum) per LM trial gives every rank the full system.  Here each rank runs the CPU oracle on its shard
(the same split rule libgpba uses: contiguous ranges of the landmarks ordered by first keyframe, balanced by
observation count), the partial systems are summed with torch.distributed over gloo, and the result must equal the
unsharded oracle system.  The CUDA + NCCL version of the same contract is exercised by `bench.py --gpus N`.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pygpba import synth
from pygpba.problem import Problem


def shard_masks(P, nranks):
    """libgpba's split (Solver::build_structure): landmarks ordered by (first keyframe, point id), cut by observation count."""
    first_kf = np.full(P.n_pt, P.n_kf, np.int64)
    np.minimum.at(first_kf, P.obs_pt, P.rec_kf2[P.obs_rec])
    active = np.bincount(P.obs_pt, minlength=P.n_pt) > 0
    order = np.lexsort((np.arange(P.n_pt), first_kf))
    order = order[active[order]]
    cnt = np.concatenate([[0], np.cumsum(np.bincount(P.obs_pt, minlength=P.n_pt)[order])])
    total = cnt[-1]
    cuts = [0] + [int(np.searchsorted(cnt, total * r // nranks, side="left")) for r in range(1, nranks)] + [len(order)]
    masks = []
    for r in range(nranks):
        m = np.zeros(P.n_pt, bool)
        m[order[cuts[r]:cuts[r + 1]]] = True
        masks.append(m)
    return masks


def dense_system(oracle_mod, Q, lam, n_free):
    """Dense (12 n_free)^2 Hschur (upper blocks) + bschur of problem Q at its initial estimate."""
    o = oracle_mod.Oracle(Q)
    info = o.build_structure()
    o.compute_errors()
    o.build_system()
    o.set_lambda(lam)
    assert o.solve()
    H, bs = o.hschur()
    rows, cols = o.hschur_pattern()
    assert info.n_free_kf == n_free
    D = np.zeros((12 * n_free, 12 * n_free))
    for k, (r, c) in enumerate(zip(rows, cols)):
        D[12 * r:12 * r + 12, 12 * c:12 * c + 12] = H[k]
    return D, bs


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle_py
    P = synth.make_problem("tiny_global", n_pt=160, seed=21)
    lam = 0.37
    n_free = int((P.kf_fixed == 0).sum())
    mask = shard_masks(P, world)[rank]
    Q = P.subset_points(mask)
    if rank != 0:   # priors are replicated edges: they enter the sum exactly once, on rank 0
        Q = Problem(**{**Q.__dict__, "prior_kf1": [], "prior_kf2": [], "velp_kf": []})
    D, bs = dense_system(oracle_py, Q, lam, n_free)
    if rank != 0:   # so does the lambda damping of Hpp
        D[np.diag_indices_from(D)] -= lam
    t = torch.from_numpy(np.concatenate([D.ravel(), bs]))
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    if rank == 0:
        Dfull, bfull = dense_system(oracle_py, P, lam, n_free)
        got = t.numpy()
        ok = np.allclose(got[:D.size].reshape(D.shape), Dfull, rtol=1e-9, atol=1e-9 * np.abs(Dfull).max()) and \
            np.allclose(got[D.size:], bfull, rtol=1e-9, atol=1e-9 * np.abs(bfull).max())
        cover = int(sum(m.sum() for m in shard_masks(P, world)))
        out.put((bool(ok), cover, int((np.bincount(P.obs_pt, minlength=P.n_pt) > 0).sum())))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_reduced_system_sums_to_the_full_one(oracle_mod):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    ok, cover, n_active = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok
    assert cover == n_active          # the shards partition the active landmarks


def test_shards_are_balanced_and_disjoint():
    P = synth.make_problem("c1")
    for n in (2, 4, 8):
        masks = shard_masks(P, n)
        assert np.all(np.sum(masks, axis=0) <= 1)
        obs = np.array([m[P.obs_pt].sum() for m in masks])
        assert obs.sum() == P.n_obs and obs.max() <= 1.25 * obs.mean() + 64


# ---------------------------------------------------------------------------------------------------------------------
# The frame-rate paths shard without a data-path collective: frames (pose-only) and hypotheses (velocity RANSAC) are
# independent.  world_size 2 over gloo: every rank runs the oracle on its contiguous share of the frames, the results are
# gathered, and the union must be the unsharded result bit for bit.
def _pose_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle_py
    from pygpba import pose as PO
    B = PO.make_pose_batch(n_frames=7, n_pt=500, A=2, outliers=0.1, seed=63, fix_prev=True)
    sub, frames = PO.shard_frames(B, rank, world)
    R = oracle_py.pose_optimize(sub)
    pose = torch.zeros(B.n_frames, 7, dtype=torch.float64); inl = torch.zeros(B.n_frames, dtype=torch.int64)
    pose[torch.from_numpy(frames)] = torch.from_numpy(R.cur_pose)
    inl[torch.from_numpy(frames)] = torch.from_numpy(R.n_inliers.astype(np.int64))
    covered = torch.zeros(B.n_frames, dtype=torch.int64); covered[torch.from_numpy(frames)] = 1
    for t in (pose, inl, covered):
        dist.all_reduce(t, op=dist.ReduceOp.SUM)   # disjoint shards: the sum is the gather
    if rank == 0:
        full = oracle_py.pose_optimize(B)
        out.put((bool(np.array_equal(pose.numpy(), full.cur_pose)), bool(np.array_equal(inl.numpy(), full.n_inliers)),
                 covered.tolist(), [int(x) for x in np.diff(sub.obs_begin)]))
    dist.barrier()
    dist.destroy_process_group()


def test_pose_frames_shard_without_collectives(oracle_mod):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_pose_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    same_pose, same_inl, covered, _ = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert same_pose and same_inl
    assert covered == [1] * 7          # every frame on exactly one rank
