"""Measured agreement between the CPU oracle and the runs of the reference's own sources (tests/golden/ref_g2o_*.npz, minted by
oracle/_ref).  CPU only.  Writes the table committed as profiles/r02_reference_pin.txt.

    python tools/ref_pin_report.py > profiles/r02_reference_pin.txt
"""
import importlib.util
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for d in ("amc-slam_b200", "oracle", "tests"):
    sys.path.insert(0, os.path.join(ROOT, d))
import oracle_py as O  # noqa: E402
import test_whole_path_reference as T  # noqa: E402

mr = T.mr
print("Oracle (CPU restatement) against the reference's own sources run as they are (oracle/_ref/libamc_ref_g2o.so: g2o core,")
print("BlockSolverX, LinearSolverDense, Levenberg-Marquardt, AMC-SLAM's G2oTypes.cc / GaussianProcess.cc / Pose3utils.cc, compiled")
print("unmodified against stand-in Eigen / Sophus headers).  optimize(10) per case; deviations are maxima over the whole state.\n")
print("%-14s %8s %6s  %-30s %10s %10s %10s %10s %8s" % ("case", "n_obs", "iters", "LM trials (reference = oracle)", "cost rel", "pos [m]", "rot [rad]", "pt [m]", "flags"))
for key in mr.CASES:
    G = T.load(key); P = mr.make_case(key)
    o = O.Oracle(P); tr = o.optimize(mr.ITERS).summary()
    kp, kv, pt = o.state()
    ip, io = mr.samples(P)
    n = tr["n_iters"]
    acc = [i for i in range(n) if tr["chi2_after"][i] < tr["chi2_before"][i]]
    cost = max(abs(tr["chi2_after"][i] - G["chi2_stored"][i]) / G["chi2_stored"][i] for i in acc)
    fl = o.outlier_flags(T.Thresholds.local_gpba())
    ref_fl = np.unpackbits(G["flags_packed"])[:P.n_obs]
    same = tr["trials"] == [int(t) for t in G["trials"]] and n == int(G["n"])
    print("%-14s %8d %6d  %-30s %10.1e %10.1e %10.1e %10.1e %8s" % (
        key, P.n_obs, n, ("same " if same else "DIFFERENT ") + str(tr["trials"]), cost, np.abs(kp[:, 4:] - G["kf_pose"][:, 4:]).max(),
        T.angle(kp[:, :4], G["kf_pose"][:, :4]).max(), np.abs(pt[ip] - G["pt_xyz"]).max(), "%d/%d" % (int((fl != ref_fl).sum()), int(ref_fl.sum()))))
F = np.load(os.path.join(ROOT, "tests", "golden", "baseline_c2.npz")); G = np.load(os.path.join(ROOT, "tests", "golden", "ref_g2o_c2.npz"))
n = int(F["tr_n_iters"][0])
a, b = np.unpackbits(F["flags_packed"])[:int(F["n_obs"])], np.unpackbits(G["flags_packed"])[:int(F["n_obs"])]
print("%-14s %8d %6d  %-30s %10.1e %10.1e %10.1e %10.1e %8s   (oracle side = the committed fixture baseline_c2.npz)" % (
    "c2 (BASELINE)", int(F["n_obs"]), n, ("same " if list(F["tr_trials"][0][:n]) == list(G["trials"]) else "DIFFERENT ") + str([int(t) for t in G["trials"]]),
    np.abs(F["tr_chi2_after"][0][:n] - G["chi2_stored"]).max() / G["chi2_stored"].min(), np.abs(F["kf_pose"][:, 4:] - G["kf_pose"][:, 4:]).max(),
    T.angle(F["kf_pose"][:, :4], G["kf_pose"][:, :4]).max(), np.abs(F["pt_xyz"] - G["pt_xyz"]).max(), "%d/%d" % (int((a != b).sum()), int(b.sum()))))
p3 = os.path.join(ROOT, "tests", "golden", "ref_g2o_c3.npz")
if os.path.exists(p3):
    F = np.load(os.path.join(ROOT, "tests", "golden", "baseline_c3.npz")); G = np.load(p3)
    a, b = np.unpackbits(F["flags_packed"])[:int(F["n_obs"])], np.unpackbits(G["flags_packed"])[:int(F["n_obs"])]
    same = all([int(t) for t in G["round%d_trials" % r]] == [int(t) for t in F["tr_trials"][r][:int(F["tr_n_iters"][r])]] for r in range(4))
    cost = max(np.abs(F["tr_chi2_after"][r][:int(F["tr_n_iters"][r])] - G["round%d_chi2_stored" % r]).max() / G["round%d_chi2_stored" % r].min() for r in range(4))
    print("%-14s %8d %6s  %-30s %10.1e %10.1e %10.1e %10.1e %8s   (BASELINE C3 as stated: 30 %% outliers, 4 rejection rounds; oracle side = baseline_c3.npz)" % (
        "c3 (BASELINE)", int(F["n_obs"]), "+".join(str(int(x)) for x in F["tr_n_iters"]), ("same" if same else "DIFFERENT") + " in all 4 rounds", cost,
        np.abs(F["kf_pose"][:, 4:] - G["kf_pose"][:, 4:]).max(), T.angle(F["kf_pose"][:, :4], G["kf_pose"][:, :4]).max(),
        np.abs(F["pt_xyz"] - G["pt_xyz"]).max(), "%d/%d" % (int((a != b).sum()), int(b.sum()))))
for key in mr.SLOW_CASES:
    if not os.path.exists(os.path.join(ROOT, "tests", "golden", "ref_g2o_" + key + ".npz")):
        continue
    G = T.load(key); P = mr.make_case(key)
    o = O.Oracle(P, threads=8); tr = o.optimize(mr.ITERS).summary()
    kp, kv, pt = o.state(); ip, io = mr.samples(P); n = tr["n_iters"]
    acc = [i for i in range(n) if tr["chi2_after"][i] < tr["chi2_before"][i]]
    fl = o.outlier_flags(T.Thresholds.local_gpba()); ref_fl = np.unpackbits(G["flags_packed"])[:P.n_obs]
    print("%-14s %8d %6d  %-30s %10.1e %10.1e %10.1e %10.1e %8s   (%d keyframes, sparse block Cholesky in the oracle, dense LDLT in the reference run)" % (
        key, P.n_obs, n, ("same " if tr["trials"] == [int(t) for t in G["trials"]] else "DIFFERENT ") + str(tr["trials"]),
        max(abs(tr["chi2_after"][i] - G["chi2_stored"][i]) / G["chi2_stored"][i] for i in acc), np.abs(kp[:, 4:] - G["kf_pose"][:, 4:]).max(),
        T.angle(kp[:, :4], G["kf_pose"][:, :4]).max(), np.abs(pt[ip] - G["pt_xyz"]).max(), "%d/%d" % (int((fl != ref_fl).sum()), int(ref_fl.sum())), P.n_kf))
print("\nflags = observations classified differently by LocalGPBA's inlier check / observations the reference flags.")
# rejection rounds
G = np.load(os.path.join(ROOT, "tests", "golden", "ref_g2o_rounds_c3.npz")); P = mr.make_rounds_case()
o = O.Oracle(P); fl, trs = o.rejection_rounds(4, mr.ITERS); kp, kv, pt = o.state()
ref_fl = np.unpackbits(G["flags_packed"])[:P.n_obs]
print("\nC3 schedule (30 %% outliers, 4 rounds, kernels off after the third; %d observations): %d flagged by the reference, %d classified"
      " differently by the oracle; trials per round %s; pos %.1e m" % (P.n_obs, int(ref_fl.sum()), int((fl != ref_fl).sum()),
      [t.summary()["trials"] == [int(x) for x in G["round%d_trials" % i]] for i, t in enumerate(trs)], np.abs(kp[:, 4:] - G["kf_pose"][:, 4:]).max()))
# extrinsics
for key in sorted(mr.EXT_CASES):
    G = np.load(os.path.join(ROOT, "tests", "golden", "ref_g2o_" + key + ".npz"))
    P, free, q_ini, info3 = mr.make_ext_case(key)
    o = O.Oracle(P); t1 = o.optimize(mr.EXT_ITERS).summary(); o.set_extrinsics(G["freed"], q_ini, info3); t2 = o.optimize(mr.EXT_ITERS).summary()
    print("extrinsic self-calibration %-9s: released %s, stage trials same %s / %s, extrinsics %.1e m %.1e rad, poses %.1e m" % (
        key, [int(x) for x in G["freed"]], t1["trials"] == [int(x) for x in G["stage1_trials"]], t2["trials"] == [int(x) for x in G["stage2_trials"]],
        np.abs(o.extrinsics()[:, 4:] - G["Tbc"][:, 4:]).max(), T.angle(o.extrinsics()[:, :4], G["Tbc"][:, :4]).max(), np.abs(o.state()[0][:, 4:] - G["kf_pose"][:, 4:]).max()))
# pose-only, velocity, essential graph
from pygpba.pose import make_pose_batch  # noqa: E402
from pygpba.velransac import make_vel_batch  # noqa: E402
for key in sorted(mr.mgp.CASES):
    G = np.load(os.path.join(ROOT, "tests", "golden", "ref_g2o_pose_" + key + ".npz")); B = make_pose_batch(**mr.mgp.CASES[key])
    out = mr.mgp.pack(B, O.pose_optimize(B))
    print("pose-only %-12s: %d matches, %d classified differently, iterations / trials per round same %s, pose %.1e" % (
        key, B.n_obs, int((out["outlier"] != G["outlier"]).sum()), bool(np.array_equal(out["trials"], G["trials"]) and np.array_equal(out["n_iters"], G["n_iters"])),
        np.abs(out["cur_pose"] - G["cur_pose"]).max()))
for key in sorted(mr.mgv.CASES):
    G = np.load(os.path.join(ROOT, "tests", "golden", "ref_g2o_vel_" + key + ".npz")); B = make_vel_batch(**mr.mgv.CASES[key])
    R = O.vel_ransac(B); well = G["inliers"] >= 30
    print("velocity RANSAC %-6s: %d hypotheses, winner %d / %d, inlier masks differing %d, velocity %.1e" % (
        key, B.n_hyp, int(R.best.value), int(G["best"]), int((R.mask != G["mask"]).sum()), np.abs(R.vel[well] - G["vel"][well]).max()))
Z = np.load(os.path.join(ROOT, "tests", "golden", "ref_g2o_posegraph.npz"))
for key in sorted(mr.POSE_GRAPHS):
    Gp = mr.make_pose_graph(key); sim3, tr = O.pose_graph_optimize(Gp, mr.PG_ITERS)
    print("essential graph %-12s: %d keyframes, %d edges, start cost rel %.1e, final cost rel %.1e, pos %.1e m (numeric Jacobians: band 1e-7 m)" % (
        key, Gp.n_kf, Gp.n_edge, abs(tr.chi2_before[0] - float(Z[key + "_chi2_start"])) / float(Z[key + "_chi2_start"]),
        abs(min(tr.summary()["chi2_after"]) - Z[key + "_chi2_stored"].min()) / Z[key + "_chi2_stored"].min(), np.abs(sim3[:, 4:7] - Z[key + "_sim3"][:, 4:7]).max()))
