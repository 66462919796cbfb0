"""Seeded synthetic multi-camera GP-BA maps for configs C1..C5 of BASELINE.json.

One generator feeds the oracle and the GPU library the same arrays, so inputs are byte-identical.
Conventions follow the reference: body twist [lin; ang] (src/GaussianProcess.cc:15), Twb as unit
quaternion xyzw + translation, the last camera is the synchronous reference camera
(MultiKeyFrame::mTbc.back(), src/G2oTypes.cc:49), async camera c of keyframe k is captured at
t_k - U(0.1,0.9)*dt (src/KeyFrame.cc:132-139), float-typed inputs.
"""
import numpy as np

from .problem import Problem, OBS_CLOSE, SOLVER_DENSE_CHOL, SOLVER_PCG

# name: (async cams, keyframes, points, obs/pt, mode, dt, outlier fraction, seed)
CONFIGS = {
    "c1": dict(A=2, n_kf=10, n_pt=2450, obs_per_pt=11, mode="local", dt=0.1, outliers=0.0, seed=1),
    "c2": dict(A=4, n_kf=30, n_pt=20000, obs_per_pt=15, mode="local", dt=0.1, outliers=0.0, seed=2),
    "c3": dict(A=4, n_kf=50, n_pt=33334, obs_per_pt=15, mode="local", dt=0.1, outliers=0.3, seed=3),
    "c4": dict(A=5, n_kf=1000, n_pt=500000, obs_per_pt=10, mode="global", dt=0.1, outliers=0.0, seed=4),
    "c5": dict(A=5, n_kf=10000, n_pt=2000000, obs_per_pt=10, mode="global", dt=0.25, outliers=0.0, seed=5),
    # small cases for unit tests
    "tiny": dict(A=2, n_kf=5, n_pt=40, obs_per_pt=8, mode="local", dt=0.1, outliers=0.0, seed=11),
    "tiny_global": dict(A=2, n_kf=8, n_pt=120, obs_per_pt=8, mode="global", dt=0.1, outliers=0.0, seed=12),
    "loop": dict(A=3, n_kf=60, n_pt=3000, obs_per_pt=10, mode="global", dt=0.1, outliers=0.0, seed=13, lap=30),
}

IMG_W, IMG_H = 960.0, 600.0


# ------------------------------------------------------------------ vectorised quaternion / SE3 helpers
def quat_mul(a, b):
    ax, ay, az, aw = a[..., 0], a[..., 1], a[..., 2], a[..., 3]
    bx, by, bz, bw = b[..., 0], b[..., 1], b[..., 2], b[..., 3]
    q = np.stack([aw * bx + ax * bw + ay * bz - az * by,
                  aw * by + ay * bw + az * bx - ax * bz,
                  aw * bz + az * bw + ax * by - ay * bx,
                  aw * bw - ax * bx - ay * by - az * bz], -1)
    return q / np.linalg.norm(q, axis=-1, keepdims=True)


def quat_rot(q, p):
    qv = q[..., :3]
    uv = 2.0 * np.cross(qv, p)
    return p + q[..., 3:4] * uv + np.cross(qv, uv)


def quat_to_R(q):
    x, y, z, w = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    R = np.empty(q.shape[:-1] + (3, 3))
    R[..., 0, 0] = 1 - 2 * (y * y + z * z); R[..., 0, 1] = 2 * (x * y - z * w); R[..., 0, 2] = 2 * (x * z + y * w)
    R[..., 1, 0] = 2 * (x * y + z * w); R[..., 1, 1] = 1 - 2 * (x * x + z * z); R[..., 1, 2] = 2 * (y * z - x * w)
    R[..., 2, 0] = 2 * (x * z - y * w); R[..., 2, 1] = 2 * (y * z + x * w); R[..., 2, 2] = 1 - 2 * (x * x + y * y)
    return R


def hat(w):
    H = np.zeros(w.shape[:-1] + (3, 3))
    H[..., 0, 1] = -w[..., 2]; H[..., 0, 2] = w[..., 1]
    H[..., 1, 0] = w[..., 2]; H[..., 1, 2] = -w[..., 0]
    H[..., 2, 0] = -w[..., 1]; H[..., 2, 1] = w[..., 0]
    return H


def se3_exp(xi):
    """xi [...,6] = [rho; phi] -> (quat xyzw [...,4], t [...,3])."""
    rho, phi = xi[..., :3], xi[..., 3:]
    th2 = np.sum(phi * phi, -1)
    th = np.sqrt(th2)
    small = th < 1e-8
    ths = np.where(small, 1.0, th)
    imag = np.where(small, 0.5 - th2 / 48.0, np.sin(0.5 * ths) / ths)
    q = np.concatenate([imag[..., None] * phi, np.where(small, 1.0 - th2 / 8.0, np.cos(0.5 * ths))[..., None]], -1)
    q = q / np.linalg.norm(q, axis=-1, keepdims=True)
    a = np.where(small, 0.5, (1 - np.cos(ths)) / (ths * ths))
    b = np.where(small, 1.0 / 6.0, (ths - np.sin(ths)) / (ths ** 3))
    Om = hat(phi)
    V = np.eye(3) + a[..., None, None] * Om + b[..., None, None] * (Om @ Om)
    t = np.einsum("...ij,...j->...i", V, rho)
    return q, t


def se3_mul(qa, ta, qb, tb):
    return quat_mul(qa, qb), ta + quat_rot(qa, tb)


def _f32(a):
    return np.asarray(a, np.float32).astype(np.float64)


def _ring_extrinsics(n_cam):
    """Tbc for a ring of cameras: camera z forward / x right / y down, body x forward / y left / z up."""
    out = np.zeros((n_cam, 7))
    for c in range(n_cam):
        # reference camera (last) looks forward; async cameras spread over the remaining yaw angles
        yaw = 0.0 if c == n_cam - 1 else 2 * np.pi * (c + 1) / n_cam
        fwd = np.array([np.cos(yaw), np.sin(yaw), 0.0])
        right = np.array([np.sin(yaw), -np.cos(yaw), 0.0])
        down = np.array([0.0, 0.0, -1.0])
        R = np.stack([right, down, fwd], 1)  # columns = camera axes in body frame
        w = 0.5 * np.sqrt(max(0.0, 1 + R[0, 0] + R[1, 1] + R[2, 2]))
        if w > 1e-6:
            q = np.array([(R[2, 1] - R[1, 2]) / (4 * w), (R[0, 2] - R[2, 0]) / (4 * w), (R[1, 0] - R[0, 1]) / (4 * w), w])
        else:  # 180 deg
            x = np.sqrt(max(0.0, (1 + R[0, 0]) / 2)); y = np.sqrt(max(0.0, (1 + R[1, 1]) / 2)); z = np.sqrt(max(0.0, (1 + R[2, 2]) / 2))
            q = np.array([x, np.copysign(y, R[0, 1]), np.copysign(z, R[0, 2]), 0.0])
        t = 0.4 * fwd + np.array([0, 0, 1.2])
        q = _f32(q)
        out[c, :4] = q / np.linalg.norm(q)  # SE3f -> cast<double>() renormalises (so3.hpp:480-487)
        out[c, 4:] = _f32(t)
    return out


def make_problem(name="c1", *, seed=None, mode=None, linear_solver=None, **override):
    cfg = dict(CONFIGS[name])
    cfg.update(override)
    if seed is not None:
        cfg["seed"] = seed
    if mode is not None:
        cfg["mode"] = mode
    rng = np.random.default_rng(cfg["seed"])
    A, n_kf, n_pt_target, opp, dt = cfg["A"], cfg["n_kf"], cfg["n_pt"], cfg["obs_per_pt"], cfg["dt"]
    n_cam = A + 1
    is_global = cfg["mode"] == "global"

    # ---- cameras (float intrinsics, GeometricCamera.h:101)
    cam_intr = np.tile(_f32([500.0, 500.0, 480.0, 300.0]), (n_cam, 1))
    cam_intr[:, 0] += _f32(np.arange(n_cam) * 1.5)
    cam_intr[:, 1] += _f32(np.arange(n_cam) * 1.25)
    cam_intr = _f32(cam_intr)
    cam_Tbc = _ring_extrinsics(n_cam)
    bf = float(np.float32(501.7))

    # ---- ground-truth trajectory: piecewise-constant body twist, closes a lap every `lap` keyframes
    # global maps drive the same loop twice (second lap = the post-loop-closure revisit): C4 500 + 500
    # keyframes, C5 5000 + 5000 (a 5 km loop driven twice = 10 km); local windows never close a loop
    lap = cfg.get("lap", n_kf // 2 if is_global and n_kf >= 600 else 8 * n_kf)
    k = np.arange(n_kf)
    ph = 2 * np.pi * k / lap
    w0 = 2 * np.pi / (lap * dt)
    speed = 4.0 if dt <= 0.1 else 4.0  # 0.4 m/KF (C1-C4), 1 m/KF (C5)
    twist = np.stack([speed + 0.2 * np.sin(ph), 0.1 * np.sin(2 * ph), 0.05 * np.cos(ph),
                      0.01 * np.sin(ph), 0.01 * np.cos(2 * ph), w0 * (1 + 0.1 * np.sin(ph))], 1)
    kf_time = k * dt
    q_true = np.zeros((n_kf, 4)); q_true[0, 3] = 1.0
    t_true = np.zeros((n_kf, 3))
    dq, dtv = se3_exp(twist * dt)
    for i in range(1, n_kf):
        q_true[i], t_true[i] = se3_mul(q_true[i - 1], t_true[i - 1], dq[i - 1], dtv[i - 1])

    # ---- records: async cameras of KF k>=1 interpolate (k-1, k); the reference camera is synchronous
    rec_kf1, rec_kf2, rec_cam, rec_t = [], [], [], []
    rec_of = -np.ones((n_kf, n_cam), np.int64)
    frac = rng.uniform(0.1, 0.9, size=(n_kf, A))
    for kk in range(n_kf):
        if kk >= 1:
            for c in range(A):
                rec_of[kk, c] = len(rec_kf1)
                rec_kf1.append(kk - 1); rec_kf2.append(kk); rec_cam.append(c)
                rec_t.append(kf_time[kk] - frac[kk, c] * dt)
        rec_of[kk, A] = len(rec_kf1)
        rec_kf1.append(-1); rec_kf2.append(kk); rec_cam.append(A); rec_t.append(kf_time[kk])
    rec_kf1 = np.array(rec_kf1, np.int32); rec_kf2 = np.array(rec_kf2, np.int32)
    rec_cam = np.array(rec_cam, np.int32); rec_t = np.array(rec_t)
    n_rec = len(rec_kf1)
    # true capture pose of each record: T(t) = T_{k-1} exp((t - t_{k-1}) twist_{k-1})  (exact for piecewise-constant twist)
    base = np.where(rec_kf1 >= 0, rec_kf1, rec_kf2)
    tau = rec_t - kf_time[base]
    qd, td = se3_exp(twist[base] * tau[:, None])
    q_rec, t_rec = se3_mul(q_true[base], t_true[base], qd, td)
    q_wc, t_wc = se3_mul(q_rec, t_rec, cam_Tbc[rec_cam, :4], cam_Tbc[rec_cam, 4:])
    R_wc = quat_to_R(q_wc)  # Xc = R_wc^T (Xw - t_wc)

    # ---- points + visibility
    W = 2 * opp                                   # candidate keyframe window
    loop_frac = 0.03 if is_global else 0.0
    pts_l, obs_l = [], []                         # obs rows: (pt_local, rec, u, v, depth)
    n_have, chunk = 0, 50000
    n_gen_target = int(n_pt_target * 1.25) + 16
    while n_have < n_pt_target:
        m = min(chunk, n_gen_target)
        k0 = rng.integers(0, n_kf, size=m)
        # aim each point at a random camera's optical axis (+-0.6 rad) so that it is seen over the window
        cam_yaw = np.array([0.0 if c == n_cam - 1 else 2 * np.pi * (c + 1) / n_cam for c in range(n_cam)])
        ang = cam_yaw[rng.integers(0, n_cam, size=m)] + rng.uniform(-0.6, 0.6, m)
        dist = rng.uniform(2.0, 60.0, m)
        lateral = np.clip(dist * np.sin(ang), -30.0, 30.0)
        ahead = dist * np.cos(ang)
        height = rng.uniform(-5.0, 5.0, m) + 1.2
        pb = np.stack([ahead, lateral, height], 1)
        Xw = quat_rot(q_true[k0], pb) + t_true[k0]
        offs = np.arange(W) - W // 2
        kk = k0[:, None] + offs[None, :]                                   # [m, W]
        revisit = rng.uniform(size=m) < loop_frac
        kk2 = np.where(revisit[:, None], kk + lap, -1)
        kk = np.concatenate([kk, kk2], 1)                                  # [m, 2W]
        valid_k = (kk >= 0) & (kk < n_kf)
        kkc = np.clip(kk, 0, n_kf - 1)
        recs = rec_of[kkc]                                                 # [m, 2W, n_cam]
        valid = valid_k[:, :, None] & (recs >= 0)
        rr = np.where(valid, recs, 0)
        d = Xw[:, None, None, :] - t_wc[rr]
        Xc = np.einsum("...ji,...j->...i", R_wc[rr], d)
        z = Xc[..., 2]
        zs = np.where(z > 0.3, z, 1.0)
        u = cam_intr[rec_cam[rr], 0] * Xc[..., 0] / zs + cam_intr[rec_cam[rr], 2]
        v = cam_intr[rec_cam[rr], 1] * Xc[..., 1] / zs + cam_intr[rec_cam[rr], 3]
        vis = valid & (z > 0.5) & (z < 80.0) & (u > 8) & (u < IMG_W - 8) & (v > 8) & (v < IMG_H - 8)
        # thin to obs_per_pt per point: random priorities, keep the opp smallest among visible
        pri = np.where(vis, rng.uniform(size=vis.shape), 2.0).reshape(m, -1)
        order = np.argsort(pri, axis=1)[:, :opp]
        keep = np.zeros_like(pri, bool)
        np.put_along_axis(keep, order, True, axis=1)
        keep &= pri < 1.5
        keep = keep.reshape(vis.shape)
        nobs = keep.reshape(m, -1).sum(1)
        good = nobs >= 2
        good &= np.cumsum(good) <= (n_pt_target - n_have)
        ids = -np.ones(m, np.int64)
        ids[good] = n_have + np.arange(int(good.sum()))
        pi, wi, ci = np.nonzero(keep & good[:, None, None])              # sorted: point, window slot, camera
        # window slots of the revisit half come later in time already (kk + lap), so order = KF ascending
        obs_l.append(np.stack([ids[pi], recs[pi, wi, ci], u[pi, wi, ci], v[pi, wi, ci], z[pi, wi, ci]], 1))
        pts_l.append(Xw[good])
        n_have += int(good.sum())
    pt_true = np.concatenate(pts_l)[:n_pt_target]
    obs = np.concatenate(obs_l)
    obs_pt = obs[:, 0].astype(np.int32)
    obs_rec = obs[:, 1].astype(np.int32)
    n_obs = len(obs_pt)
    depth = obs[:, 4]

    # ---- measurements: + N(0, sigma_l^2), sigma_l = 1.2^octave, invSigma2 float (mvInvLevelSigma2)
    wts = 1.2 ** (-2.0 * np.arange(8)); wts /= wts.sum()
    octave = rng.choice(8, size=n_obs, p=wts)
    sigma = 1.2 ** octave
    u_meas = obs[:, 2] + rng.normal(size=n_obs) * sigma
    v_meas = obs[:, 3] + rng.normal(size=n_obs) * sigma
    is_out = np.zeros(n_obs, bool)
    if cfg["outliers"] > 0:
        is_out = rng.uniform(size=n_obs) < cfg["outliers"]
        u_meas = np.where(is_out, rng.uniform(0, IMG_W, n_obs), u_meas)
        v_meas = np.where(is_out, rng.uniform(0, IMG_H, n_obs), v_meas)
    obs_u, obs_v = _f32(u_meas), _f32(v_meas)               # cv::KeyPoint::pt is float
    inv_sigma2 = _f32(1.0 / (1.2 ** (2.0 * octave)))
    flags = np.where(depth < 10.0, OBS_CLOSE, 0).astype(np.uint8)

    # ---- initial estimate: truth (+) noise, then float-rounded like the map state
    dpose = np.concatenate([rng.normal(size=(n_kf, 3)) * 0.05, rng.normal(size=(n_kf, 3)) * np.deg2rad(0.5)], 1)
    dpose[0] = 0.0
    qn, tn = se3_exp(dpose)
    q_init, t_init = se3_mul(q_true, t_true, qn, tn)
    q_init = _f32(q_init); q_init /= np.linalg.norm(q_init, axis=1, keepdims=True)
    kf_pose = np.concatenate([q_init, _f32(t_init)], 1)
    kf_vel = _f32(twist + rng.normal(size=(n_kf, 6)) * 0.1 * np.array([1, 1, 1, 0.1, 0.1, 0.1]))
    kf_fixed = np.zeros(n_kf, np.uint8); kf_fixed[0] = 1
    pt_xyz = _f32(pt_true + rng.normal(size=pt_true.shape) * 0.1)

    # ---- priors: EdgeVelocity on every KF (inactive on the fixed one), EdgeGaussianPrior between consecutive KFs
    prior_kf1 = np.arange(0, n_kf - 1, dtype=np.int32)
    prior_kf2 = np.arange(1, n_kf, dtype=np.int32)
    velp_kf = np.arange(0, n_kf, dtype=np.int32)

    hub_mono = float(np.float32(np.sqrt(5.991)))           # const float thHuberMono (Optimizer.cc:138)
    hub_stereo = float(np.float32(np.sqrt(7.815)))
    if linear_solver is None:
        linear_solver = SOLVER_DENSE_CHOL
    prob = Problem(
        cam_intr=cam_intr, cam_Tbc=cam_Tbc, bf=bf, kf_pose=kf_pose, kf_vel=kf_vel, kf_time=kf_time, kf_fixed=kf_fixed,
        pt_xyz=pt_xyz, rec_kf1=rec_kf1, rec_kf2=rec_kf2, rec_cam=rec_cam, rec_t=rec_t,
        obs_u=obs_u, obs_v=obs_v, obs_ur=None, obs_inv_sigma2=inv_sigma2, obs_rec=obs_rec, obs_pt=obs_pt,
        obs_flags=flags, prior_kf1=prior_kf1, prior_kf2=prior_kf2, velp_kf=velp_kf,
        qc=[0.02, 0.02, 0.02, 0.002, 0.002, 0.002],       # Gaussian.Qc, orb_multicam.yaml:15
        huber_mono=hub_mono, huber_stereo=hub_stereo,
        huber_prior=21.026 if is_global else 0.0,          # Optimizer.cc:128-130 vs :903-910
        lambda_init=1e-5 if is_global else 1.0,            # Optimizer.cc:75 vs :854
        linear_solver=linear_solver,
        meta=dict(name=name, mode=cfg["mode"], seed=cfg["seed"], n_kf=n_kf, n_pt=len(pt_xyz), n_obs=n_obs,
                  n_cam=n_cam, n_rec=n_rec, lap=int(lap), outliers=float(cfg["outliers"])),
        truth=dict(kf_q=q_true, kf_t=t_true, kf_vel=twist, pt=pt_true, is_outlier=is_out),
    )
    return prob


def add_stereo(prob, fraction=0.5, seed=0, gp_fraction=0.0):
    """Turn a fraction of the synchronous (reference camera) observations into EdgeStereo (ur >= 0) and, with
    gp_fraction > 0, a fraction of the asynchronous ones into EdgeStereoGP (src/G2oTypes.cc:369-443)."""
    rng = np.random.default_rng(seed)
    ur = -np.ones(prob.n_obs)
    sync = prob.rec_kf1[prob.obs_rec] < 0
    draw = rng.uniform(size=prob.n_obs)
    pick = (sync & (draw < fraction)) | (~sync & (draw < gp_fraction))
    # ur = u - bf / z with z from the true geometry approximated through the current estimate is not needed for
    # parity tests: any plausible value works, use a depth of 8..40 m.
    z = rng.uniform(8.0, 40.0, prob.n_obs)
    ur[pick] = _f32(prob.obs_u[pick] - prob.bf / z[pick])
    prob.obs_ur = np.ascontiguousarray(ur)
    return prob
