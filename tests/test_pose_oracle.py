"""CPU side of the pose-refinement path (src/geometry/pnp.cc:38-71): the BA oracle solves the one-camera problem, and
a numpy restatement of the ALGORITHM pose_refine.cu runs (6 x 6 normal equations accumulated per pose, model cost
change from y^T g - y^T H y / 2, Jacobi scaling, Ceres' trust-region schedule) reproduces the oracle's iterations.
The kernel itself is compared with the oracle in tests/test_pose_gpu.py; this file pins what it is compared with.
"""
import numpy as np
import pytest

from xrsfm_b200 import synth

from . import oracle_lib as O

POSE_OPTS = dict(max_iterations=10, function_tolerance=1e-6, parameter_tolerance=1e-8, gradient_tolerance=1e-10,
                 initial_radius=1e4)


def _oracle(batch, p, **kw):
    o = dict(POSE_OPTS)
    o.update(kw)
    sc = synth.pose_as_scene(batch, p)
    s = O.summary_dict(O.ba_solve(sc, O.ba_options(**o)))
    return sc.cam_q[0], sc.cam_t[0], s


def _normal_equations_lm(batch, p, **kw):
    """The loop of k_pose_refine, in numpy, with the oracle's per-observation model (xro_ba_eval_obs)."""
    o = dict(POSE_OPTS)
    o.update(kw)
    opts = O.ba_options(**o)
    lo, hi = int(batch["offsets"][p]), int(batch["offsets"][p + 1])
    idx = [i for i in range(lo, hi) if batch["inlier"][i]]
    model, intr = int(batch["intr_model"][p]), batch["intr"][p]
    q, t = batch["q"][p].copy(), batch["t"][p].copy()
    sc = np.ones(6)

    def linearise(q, t):
        H, g, c = np.zeros((6, 6)), np.zeros(6), 0.0
        for i in idx:
            r, Jd, Jt, _, rho0, _ = O.ba_eval_obs(q, t, batch["xyz"][i], model, intr, batch["uv"][i], opts, robustify=True)
            J = np.hstack([Jd, Jt]) * sc
            H += J.T @ J
            g += J.T @ r
            c += 0.5 * rho0
        return H, g, c

    def cost(q, t):
        return sum(0.5 * O.ba_eval_obs(q, t, batch["xyz"][i], model, intr, batch["uv"][i], opts)[4] for i in idx)

    def grad_max(q, g):
        return max(np.abs(q - O.quat_plus(q, -g[:3] / sc[:3])).max(), np.abs(g[3:] / sc[3:]).max())

    out = dict(num_lm_iterations=0, num_successful_steps=1, num_unsuccessful_steps=0, termination_type=1)
    H, g, x_cost = linearise(q, t)
    out["initial_cost"] = min_cost = x_cost
    sc = 1.0 / (1.0 + np.sqrt(np.diag(H)))
    H, g = H * np.outer(sc, sc), g * sc
    xnorm, gmax = np.sqrt(q @ q + t @ t), grad_max(q, g)
    radius, dec, it, invalid = o["initial_radius"], 2.0, 0, 0
    while True:
        if it >= o["max_iterations"]:
            out["termination_type"] = 1
            break
        if gmax <= o["gradient_tolerance"] or radius <= 1e-32:
            out["termination_type"] = 0
            break
        it += 1
        A = H + np.diag(np.clip(np.diag(H), 1e-6, 1e32) / radius)
        try:
            L = np.linalg.cholesky(A)
            y, ok = np.linalg.solve(L.T, np.linalg.solve(L, g)), True
        except np.linalg.LinAlgError:
            y, ok = np.zeros(6), False
        model_change = y @ g - 0.5 * y @ H @ y
        cq, ct = O.quat_plus(q, -y[:3] * sc[:3]), t - y[3:] * sc[3:]
        step_norm = np.sqrt(((q - cq) ** 2).sum() + ((t - ct) ** 2).sum())
        out["num_lm_iterations"] += 1
        cand = cost(cq, ct)
        if not (ok and np.isfinite(cand) and model_change > 0):
            invalid += 1
            if invalid >= 5:
                out["termination_type"] = 2
                break
            radius *= 0.5
            out["num_unsuccessful_steps"] += 1
            continue
        invalid = 0
        if step_norm <= o["parameter_tolerance"] * (xnorm + o["parameter_tolerance"]):
            out["termination_type"] = 0
            break
        change = x_cost - cand
        if abs(change) <= o["function_tolerance"] * x_cost:
            out["termination_type"] = 0
            break
        rho = change / model_change
        min_cost = min(min_cost, cand)
        if rho > 1e-3:
            q, t, x_cost = cq, ct, cand
            xnorm = np.sqrt(q @ q + t @ t)
            out["num_successful_steps"] += 1
            radius, dec = min(1e16, radius / max(1.0 / 3.0, 1.0 - (2.0 * rho - 1.0) ** 3)), 2.0
            H, g, _ = linearise(q, t)
            gmax = grad_max(q, g)
        else:
            radius, dec = radius / dec, dec * 2.0
            out["num_unsuccessful_steps"] += 1
    out["final_cost"] = min(out["initial_cost"], min_cost)
    return q, t, out


def test_pose_batch_generator_is_deterministic_and_consistent():
    a = synth.make_pose_batch(6, seed=5, behind_frac=0.05)
    b = synth.make_pose_batch(6, seed=5, behind_frac=0.05)
    for k in a:
        assert np.array_equal(a[k], b[k]), k
    assert a["offsets"][0] == 0 and np.all(np.diff(a["offsets"]) >= 20)
    assert set(a["intr_model"]) == {0, 1, 2, 3, 4}
    np.testing.assert_allclose(np.linalg.norm(a["q"], axis=1), 1.0, atol=1e-12)
    # at the true pose the unmasked measurements are the projections up to the noise, bar the outliers / points behind
    opts = O.ba_options()
    for p in range(6):
        lo, hi = int(a["offsets"][p]), int(a["offsets"][p + 1])
        res = np.array([O.ba_eval_obs(a["gt_q"][p], a["gt_t"][p], a["xyz"][i], int(a["intr_model"][p]), a["intr"][p],
                                      a["uv"][i], opts)[0] for i in range(lo, hi)])
        err = np.linalg.norm(res, axis=1)
        assert np.median(err) < 2.0 and (err < 5.0).mean() > 0.85


def test_oracle_refines_poses_towards_the_truth():
    batch = synth.make_pose_batch(10, seed=6)
    for p in range(10):
        q, t, s = _oracle(batch, p)
        assert s["termination_type"] == 0 and 2 <= s["num_lm_iterations"] <= 10
        assert s["final_cost"] < s["initial_cost"]
        assert np.linalg.norm(t - batch["gt_t"][p]) < np.linalg.norm(batch["t"][p] - batch["gt_t"][p])
        ang = 2 * np.arccos(min(1.0, abs(float(q @ batch["gt_q"][p]))))
        assert np.rad2deg(ang) < 0.2
        assert abs(np.linalg.norm(q) - 1.0) < 1e-12


@pytest.mark.parametrize("case", ["near", "far", "three_iterations", "tight_radius", "small_huber", "tiny"])
def test_normal_equation_loop_equals_oracle(case):
    kw = {}
    if case == "far":
        batch = synth.make_pose_batch(4, seed=12, rot_deg=8.0, trans_frac=0.2, outlier_frac=0.15, min_pts=20, max_pts=40)
    elif case == "tiny":
        batch = synth.make_pose_batch(3, seed=17, min_pts=4, max_pts=5, with_mask=False, outlier_frac=0.0)
    else:
        batch = synth.make_pose_batch(3, seed=13, max_pts=50, behind_frac=0.03)
        kw = {"three_iterations": dict(max_iterations=3), "tight_radius": dict(initial_radius=1e2),
              "small_huber": dict(huber_a=2.0)}.get(case, {})
    for p in range(batch["q"].shape[0]):
        oq, ot, os_ = _oracle(batch, p, **kw)
        q, t, s = _normal_equations_lm(batch, p, **kw)
        for k in ("num_lm_iterations", "num_successful_steps", "num_unsuccessful_steps", "termination_type"):
            assert s[k] == os_[k], (case, p, k)
        assert s["initial_cost"] == pytest.approx(os_["initial_cost"], rel=1e-10)
        assert s["final_cost"] == pytest.approx(os_["final_cost"], rel=1e-8, abs=1e-9)
        np.testing.assert_allclose(q, oq, rtol=0, atol=1e-8)
        np.testing.assert_allclose(t, ot, rtol=0, atol=1e-8)


def test_refine_poses_checks_its_arguments_before_the_library():
    from xrsfm_b200 import pnp
    b = synth.make_pose_batch(2, seed=3)
    with pytest.raises(TypeError):  # q must be updatable in place
        pnp.refine_poses(b["offsets"], b["uv"], b["xyz"], b["intr"], b["intr_model"], b["q"].tolist(), b["t"])
    with pytest.raises(ValueError):
        pnp.refine_poses(b["offsets"], b["uv"][:-1], b["xyz"], b["intr"], b["intr_model"], b["q"].copy(), b["t"].copy())
    with pytest.raises(ValueError):
        pnp.refine_poses(b["offsets"], b["uv"], b["xyz"], b["intr"][:, :4], b["intr_model"], b["q"].copy(), b["t"].copy())
    with pytest.raises(ValueError):
        pnp.refine_poses(b["offsets"], b["uv"], b["xyz"], b["intr"], b["intr_model"], b["q"].copy(), b["t"].copy(),
                         inlier_mask=b["inlier"][:-1])
    with pytest.raises(TypeError):
        pnp.make_options(no_such_option=1)
    o = pnp.make_options()
    assert (o.max_iterations, o.function_tolerance, o.parameter_tolerance, o.initial_radius) == (10, 1e-6, 1e-8, 1e4)
    assert o.huber_a == 5.99
