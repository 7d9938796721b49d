"""Independent numpy/scipy restatement of the BA linear algebra — the oracle's second opinion.

Deliberately different from oracle/ba_oracle.cpp: rotation through an explicit rotation
matrix, Jacobians by central differences through the manifold Plus, and the LM step from the
FULL damped normal equations (scipy sparse direct solve) instead of Schur elimination.
Cites the same reference lines: cost_factor_ceres.h:19-40, camera_model.hpp:93-210.
"""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from xrsfm_b200 import synth

HUBER_A = 5.99
MIN_DEPTH = 1e-2
NEG = 12.0


def project(model, intr, xy):
    x, y = xy
    if model == 0:
        f, cx, cy = intr[:3]
        return np.array([f * 2 * x + cx, f * 2 * y + cy])  # reference quirk: duv = xy
    if model == 1:
        fx, fy, cx, cy = intr[:4]
        return np.array([fx * 2 * x + cx, fy * 2 * y + cy])
    if model == 2:
        f, cx, cy, k = intr[:4]
        r2 = x * x + y * y
        return np.array([f * (x + x * k * r2) + cx, f * (y + y * k * r2) + cy])
    if model == 3:
        fx, fy, cx, cy, k = intr[:5]
        r2 = x * x + y * y
        return np.array([fx * (x + x * k * r2) + cx, fy * (y + y * k * r2) + cy])
    fx, fy, cx, cy, k1, k2, p1, p2 = intr[:8]
    r2 = x * x + y * y
    rad = k1 * r2 + k2 * r2 * r2
    du = x * rad + 2 * p1 * x * y + p2 * (r2 + 2 * x * x)
    dv = y * rad + 2 * p2 * x * y + p1 * (r2 + 2 * y * y)
    return np.array([fx * (x + du) + cx, fy * (y + dv) + cy])


def residual(q, t, X, model, intr, uv):
    R = synth.rotmat_from_quat(np.asarray(q) / np.linalg.norm(q))
    pc = R @ X + t
    if pc[2] < MIN_DEPTH:
        return np.array([NEG, NEG])
    return project(model, intr, pc[:2] / pc[2]) - uv


def plus_q(q, d):
    n = np.linalg.norm(d)
    if n == 0:
        return np.array(q, dtype=float)
    dq = np.concatenate([np.sin(n) / n * d, [np.cos(n)]])
    return synth.quat_mul(dq, q)


def huber(r):
    s = r @ r
    if s > HUBER_A ** 2:
        rt = np.sqrt(s)
        return 2 * HUBER_A * rt - HUBER_A ** 2, HUBER_A / rt
    return s, 1.0


def numeric_jacobians(q, t, X, model, intr, uv, h=1e-6):
    """Central differences in the tangent space: (Jd, Jt, JX), each 2x3, raw."""
    Jd, Jt, JX = np.zeros((2, 3)), np.zeros((2, 3)), np.zeros((2, 3))
    for k in range(3):
        e = np.zeros(3)
        e[k] = h
        Jd[:, k] = (residual(plus_q(q, e), t, X, model, intr, uv) -
                    residual(plus_q(q, -e), t, X, model, intr, uv)) / (2 * h)
        Jt[:, k] = (residual(q, t + e, X, model, intr, uv) - residual(q, t - e, X, model, intr, uv)) / (2 * h)
        JX[:, k] = (residual(q, t, X + e, model, intr, uv) - residual(q, t, X - e, model, intr, uv)) / (2 * h)
    return Jd, Jt, JX


class Layout:
    """Column layout of the reduced program: camera q/t tangent blocks, then points."""

    def __init__(self, sc):
        cams_used = np.zeros(sc.n_cams, bool)
        cams_used[sc.obs_cam] = True
        pts_used = np.zeros(sc.n_pts, bool)
        pts_used[sc.obs_pt] = True
        self.colq = -np.ones(sc.n_cams, int)
        self.colt = -np.ones(sc.n_cams, int)
        n = 0
        for c in range(sc.n_cams):
            if not cams_used[c]:
                continue
            if not sc.cam_q_fixed[c]:
                self.colq[c] = n
                n += 3
            if not sc.cam_t_fixed[c]:
                self.colt[c] = n
                n += 3
        self.nc = n
        self.colp = -np.ones(sc.n_pts, int)
        for p in range(sc.n_pts):
            if pts_used[p] and not sc.pt_fixed[p]:
                self.colp[p] = n
                n += 3
        self.n = n


def build_system(sc, lay):
    """Sparse robustified Jacobian (unscaled) and residual vector over active observations."""
    rows, cols, vals, res = [], [], [], []
    cost = 0.0
    r_i = 0
    for o in range(sc.n_obs):
        c, p = sc.obs_cam[o], sc.obs_pt[o]
        if lay.colq[c] < 0 and lay.colt[c] < 0 and lay.colp[p] < 0:
            continue
        q, t, X = sc.cam_q[c], sc.cam_t[c], sc.pts[p]
        model, intr = sc.intr_model[sc.cam_intr[c]], sc.intr[sc.cam_intr[c]]
        r = residual(q, t, X, model, intr, sc.obs_uv[o])
        if np.all(r == NEG):
            Jd = Jt = JX = np.zeros((2, 3))
        else:
            Jd, Jt, JX = numeric_jacobians(q, t, X, model, intr, sc.obs_uv[o])
        rho0, rho1 = huber(r)
        w = np.sqrt(rho1)
        cost += 0.5 * rho0
        for blk, col in ((Jd, lay.colq[c]), (Jt, lay.colt[c]), (JX, lay.colp[p])):
            if col < 0:
                continue
            for a in range(2):
                for k in range(3):
                    rows.append(r_i + a), cols.append(col + k), vals.append(w * blk[a, k])
        res.extend(w * r)
        r_i += 2
    J = sp.csr_matrix((vals, (rows, cols)), shape=(r_i, lay.n))
    return J, np.array(res), cost


def lm_step(sc, radius, scale=None):
    """One Ceres-style LM step from the full normal equations.

    Returns dict(delta, scale, model_cost_change, cost, J, r)."""
    lay = Layout(sc)
    J, r, cost = build_system(sc, lay)
    if scale is None:
        scale = 1.0 / (1.0 + np.sqrt(np.asarray(J.multiply(J).sum(axis=0)).ravel()))
    Js = J @ sp.diags(scale)
    diag = np.clip(np.asarray(Js.multiply(Js).sum(axis=0)).ravel(), 1e-6, 1e32)
    H = (Js.T @ Js + sp.diags(diag / radius)).tocsc()
    y = spla.spsolve(H, Js.T @ r)
    step = -y
    m = Js @ step
    model_cost_change = -m @ (r + m / 2)
    return dict(lay=lay, delta=step * scale, scale=scale, model_cost_change=model_cost_change,
                cost=cost, step=step)


def apply_delta(sc, lay, delta):
    out = sc.copy_state()
    for c in range(sc.n_cams):
        if lay.colq[c] >= 0:
            out.cam_q[c] = plus_q(sc.cam_q[c], delta[lay.colq[c]: lay.colq[c] + 3])
        if lay.colt[c] >= 0:
            out.cam_t[c] = sc.cam_t[c] + delta[lay.colt[c]: lay.colt[c] + 3]
    for p in range(sc.n_pts):
        if lay.colp[p] >= 0:
            out.pts[p] = sc.pts[p] + delta[lay.colp[p]: lay.colp[p] + 3]
    return out


def total_cost(sc):
    lay = Layout(sc)
    cost = 0.0
    for o in range(sc.n_obs):
        c, p = sc.obs_cam[o], sc.obs_pt[o]
        if lay.colq[c] < 0 and lay.colt[c] < 0 and lay.colp[p] < 0:
            continue
        r = residual(sc.cam_q[c], sc.cam_t[c], sc.pts[p], sc.intr_model[sc.cam_intr[c]],
                     sc.intr[sc.cam_intr[c]], sc.obs_uv[o])
        cost += 0.5 * huber(r)[0]
    return cost
