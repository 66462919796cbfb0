"""Independent numpy/scipy mirror of the Lie-group / GP layer -- TEST INFRASTRUCTURE ONLY.

Deliberately NOT a transcription of oracle/lie.h: SE(3) exp/log go through scipy's generic matrix
exponential / logarithm of the 4x4 homogeneous matrix, the SE(3) Jacobians through their defining
power series J_l(xi) = sum_n ad(xi)^n / (n+1)!  (Barfoot, State Estimation for Robotics, eq. 7.79),
and the GP interpolation weights through the closed forms of SURVEY.md fact 0.8.  Used by
tests/test_oracle_math.py to pin the C++ restatement.
"""
import numpy as np
from scipy.linalg import expm, logm


def hat3(w):
    return np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0.0]])


def hat6(xi):
    """4x4 se(3) element, xi = [rho; phi] (translation first, src/GaussianProcess.cc:15)."""
    M = np.zeros((4, 4))
    M[:3, :3] = hat3(xi[3:])
    M[:3, 3] = xi[:3]
    return M


def vee6(M):
    return np.array([M[0, 3], M[1, 3], M[2, 3], M[2, 1], M[0, 2], M[1, 0]])


def exp_se3(xi):
    return expm(hat6(np.asarray(xi, float)))


def log_se3(T):
    return vee6(np.real(logm(T)))


def ad(xi):
    """curly hat: [[phi^, rho^],[0, phi^]] (src/Pose3utils.cc:111-118)."""
    A = np.zeros((6, 6))
    A[:3, :3] = hat3(xi[3:]); A[:3, 3:] = hat3(xi[:3]); A[3:, 3:] = hat3(xi[3:])
    return A


def Adj(T):
    R, t = T[:3, :3], T[:3, 3]
    A = np.zeros((6, 6))
    A[:3, :3] = R; A[3:, 3:] = R; A[:3, 3:] = hat3(t) @ R
    return A


def Jl_series(xi, terms=40):
    A = ad(np.asarray(xi, float))
    J = np.eye(6); P = np.eye(6); f = 1.0
    for n in range(1, terms):
        P = P @ A
        f *= (n + 1)
        J = J + P / f
    return J


def Jr_series(xi):
    return Jl_series(-np.asarray(xi, float))


def gp_weights(t1, t2, t):
    """(lambda11, lambda12, psi11, psi12) of SURVEY fact 0.8."""
    D = t2 - t1
    s = (t - t1) / D
    psi11 = 3 * s * s - 2 * s ** 3
    psi12 = D * (s ** 3 - s * s)
    return 1 - psi11, D * (s - 2 * s * s + s ** 3), psi11, psi12


def query_pose(T1, T2, v1, v2, t1, t2, t):
    l11, l12, p11, p12 = gp_weights(t1, t2, t)
    xi12 = log_se3(np.linalg.inv(T1) @ T2)
    arg = l12 * np.asarray(v1) + p11 * xi12 + p12 * (np.linalg.inv(Jr_series(xi12)) @ np.asarray(v2))
    return T1 @ exp_se3(arg)


def T_from7(p):
    x, y, z, w = p[:4]
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                  [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                  [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
    T = np.eye(4); T[:3, :3] = R; T[:3, 3] = p[4:7]
    return T


def reproj(T1, T2, v1, v2, t1, t2, t, Tbc, intr, Xw, gp=True):
    Twb = query_pose(T1, T2, v1, v2, t1, t2, t) if gp else T2
    Xc = np.linalg.inv(Twb @ Tbc) @ np.append(Xw, 1.0)
    return np.array([intr[0] * Xc[0] / Xc[2] + intr[2], intr[1] * Xc[1] / Xc[2] + intr[3]]), Xc[2]


def gp_edge_jacobians(T1, T2, v1, v2, t1, t2, t, Tbc, intr, bf, Xw, stereo):
    """Analytic Jacobians of the GP reprojection edges, composed from THIS module's pieces (scipy logm/expm, the power
    series of the SE(3) Jacobians, the closed-form GP weights) after the formulas of EdgeMonoGP / EdgeStereoGP::linearizeOplus
    (src/G2oTypes.cc:262-311 mono, :329-396 stereo; QueryPose with At1 / Pt1: src/GaussianProcess.cc:23-42).  Returns
    (J_kf1 [dim x 12], J_kf2 [dim x 12], J_point [dim x 3]) with the reference's block order [pose(6) | velocity(6)].
    The pose blocks carry the reference's first-order term -0.5 ad(v2), so they are NOT the derivative
    of the error; this function pins the restated FORMULAS, a numeric derivative cannot."""
    v1 = np.asarray(v1, float); v2 = np.asarray(v2, float)
    l11, l12, p11, p12 = gp_weights(t1, t2, t)
    I6 = np.eye(6); Z6 = np.zeros((6, 6))
    At1 = np.hstack([l11 * I6, l12 * I6]); Pt1 = np.hstack([p11 * I6, p12 * I6])
    xi12 = log_se3(np.linalg.inv(T1) @ T2)
    Jr_inv_xi12 = np.linalg.inv(Jr_series(xi12))
    dxi = At1 @ np.concatenate([np.zeros(6), v1]) + Pt1 @ np.concatenate([xi12, Jr_inv_xi12 @ v2])
    dT = exp_se3(dxi)
    Twb = T1 @ dT
    Tcb = np.linalg.inv(Tbc)
    Rcb = Tcb[:3, :3]
    Rbw = Twb[:3, :3].T
    Xb = (np.linalg.inv(Twb) @ np.append(Xw, 1.0))[:3]
    Xc = Rcb @ Xb + Tcb[:3, 3]
    x, y, z = Xc
    fx, fy = intr[0], intr[1]
    pj = np.array([[fx / z, 0, -fx * x / (z * z)], [0, fy / z, -fy * y / (z * z)]])
    if stereo:
        pj = np.vstack([pj, pj[0] + np.array([0, 0, bf / (z * z)])])
    SE3deriv = np.hstack([-Rcb, Rcb @ hat3(Xb)])
    J1 = -pj @ SE3deriv
    Ad_dT = Adj(exp_se3(-dxi))
    Jr_dxi = Jr_series(dxi)
    ad_v2 = ad(v2)
    ad_T12_inv = np.linalg.inv(Adj(exp_se3(xi12)))
    JinT1 = np.vstack([-Jr_inv_xi12 @ ad_T12_inv, -0.5 * ad_v2 @ (-Jr_inv_xi12 @ ad_T12_inv)])
    JinV1 = np.vstack([Z6, I6])
    JinT2 = np.vstack([Jr_inv_xi12, -0.5 * ad_v2 @ Jr_inv_xi12])
    JinV2 = np.vstack([Z6, Jr_inv_xi12])
    Ja = np.hstack([J1 @ (Jr_dxi @ Pt1 @ JinT1 + Ad_dT), J1 @ Jr_dxi @ At1 @ JinV1])
    Jj1 = J1 @ Jr_dxi @ Pt1
    Jb = np.hstack([Jj1 @ JinT2, Jj1 @ JinV2])
    Jp = -pj @ Rcb @ Rbw
    return Ja, Jb, Jp
