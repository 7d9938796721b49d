"""Batched pose refinement (xrb_pose_refine_batch, pose_refine.cu) against the BA oracle.

The reference refines a registered frame's pose with a ten-iteration Ceres solve over the inlier 2D-3D
correspondences (src/geometry/pnp.cc:38-71): the same cost functor, loss, parameterisation and minimiser as
bundle adjustment with ONE variable camera and constant points — which is how the oracle (oracle/ba_oracle.cpp,
xro_ba_solve) is asked to solve it here, pose by pose.  Tolerances: poses to 1e-8 absolute (unit quaternion,
translations of O(10)), costs to 1e-8 relative — the only difference between the two is summation order.
"""
import numpy as np
import pytest

from xrsfm_b200 import synth

from . import oracle_lib as O

POSE_OPTS = dict(max_iterations=10, function_tolerance=1e-6, parameter_tolerance=1e-8, gradient_tolerance=1e-10,
                 initial_radius=1e4)  # ceres::Solver::Options defaults + pnp.cc:57


def _oracle_batch(batch, **opts):
    o = dict(POSE_OPTS)
    o.update(opts)
    q, t, sums = batch["q"].copy(), batch["t"].copy(), []
    for p in range(q.shape[0]):
        sc = synth.pose_as_scene(batch, p)
        s = O.summary_dict(O.ba_solve(sc, O.ba_options(**o)))
        q[p], t[p] = sc.cam_q[0], sc.cam_t[0]
        sums.append(s)
    return q, t, sums


def _gpu_batch(batch, use_mask=True, **opts):
    from xrsfm_b200 import pnp
    q, t = batch["q"].copy(), batch["t"].copy()
    sums = pnp.refine_poses(batch["offsets"], batch["uv"], batch["xyz"], batch["intr"], batch["intr_model"], q, t,
                            inlier_mask=batch["inlier"] if use_mask else None, **opts)
    return q, t, sums


def _compare(batch, gq, gt, gs, oq, ot, os_, cost_rtol=1e-8, atol=1e-8):
    for p in range(gq.shape[0]):
        g, o = gs[p], os_[p]
        ctx = f"pose {p} (model {batch['intr_model'][p]}, {g['num_residuals'] // 2} correspondences)"
        assert g["num_residuals"] == o["num_residuals_reduced"], ctx
        assert g["termination_type"] == o["termination_type"], ctx
        assert g["num_lm_iterations"] == o["num_lm_iterations"], ctx
        assert g["num_successful_steps"] == o["num_successful_steps"], ctx
        assert g["initial_cost"] == pytest.approx(o["initial_cost"], rel=1e-12), ctx
        assert g["final_cost"] == pytest.approx(o["final_cost"], rel=cost_rtol, abs=1e-9), ctx
        np.testing.assert_allclose(gq[p], oq[p], rtol=0, atol=atol, err_msg=ctx)
        np.testing.assert_allclose(gt[p], ot[p], rtol=0, atol=atol, err_msg=ctx)


@pytest.mark.gpu
def test_pose_batch_equals_oracle():
    batch = synth.make_pose_batch(60, seed=11, behind_frac=0.02)
    gq, gt, gs = _gpu_batch(batch)
    oq, ot, os_ = _oracle_batch(batch)
    _compare(batch, gq, gt, gs, oq, ot, os_)
    # and it did something: every pose moved towards the truth
    err0 = np.linalg.norm(batch["t"] - batch["gt_t"], axis=1)
    err1 = np.linalg.norm(gt - batch["gt_t"], axis=1)
    assert np.all(err1 < err0)
    assert all(s["final_cost"] < s["initial_cost"] for s in gs)


@pytest.mark.gpu
def test_pose_batch_far_start_rejected_steps():
    """Starts far enough from the truth (8 degrees, 20 % of the depth) that the trust region has to work: rejected
    and invalid steps, radius changes, the iteration cap."""
    batch = synth.make_pose_batch(40, seed=12, rot_deg=8.0, trans_frac=0.2, outlier_frac=0.15, min_pts=20, max_pts=120)
    gq, gt, gs = _gpu_batch(batch)
    oq, ot, os_ = _oracle_batch(batch)
    _compare(batch, gq, gt, gs, oq, ot, os_, cost_rtol=1e-7, atol=1e-7)
    assert any(s["num_unsuccessful_steps"] > 0 or s["termination_type"] == 1 for s in gs)


@pytest.mark.gpu
def test_pose_fixed_iterations_and_other_options():
    batch = synth.make_pose_batch(16, seed=13)
    for opts in (dict(max_iterations=3), dict(max_iterations=2, fixed_iterations=1), dict(initial_radius=1e2),
                 dict(huber_a=2.0)):
        gq, gt, gs = _gpu_batch(batch, **opts)
        oq, ot, os_ = _oracle_batch(batch, **opts)
        # (fixed_iterations kept short: stepping on at the optimum makes accept / reject a round-off decision)
        _compare(batch, gq, gt, gs, oq, ot, os_, cost_rtol=1e-7, atol=1e-6 if "fixed_iterations" in opts else 1e-7)


@pytest.mark.gpu
def test_pose_matches_the_ba_engine():
    """The same problems through xrb_ba_solve (one variable camera, constant points): two product paths, one answer."""
    from xrsfm_b200 import ba
    batch = synth.make_pose_batch(6, seed=14)
    gq, gt, gs = _gpu_batch(batch)
    solver = ba.BASolver()
    for p in range(6):
        sc = synth.pose_as_scene(batch, p)
        s = solver.solve_scene(sc, **POSE_OPTS)
        assert s.termination_type == gs[p]["termination_type"]
        assert s.num_lm_iterations == gs[p]["num_lm_iterations"]
        assert s.final_cost == pytest.approx(gs[p]["final_cost"], rel=1e-8)
        np.testing.assert_allclose(sc.cam_q[0], gq[p], rtol=0, atol=1e-8)
        np.testing.assert_allclose(sc.cam_t[0], gt[p], rtol=0, atol=1e-8)


@pytest.mark.gpu
def test_pose_result_does_not_depend_on_the_batch():
    big = synth.make_pose_batch(700, seed=15, max_pts=150)  # more poses than resident CTAs x SMs / 2
    gq, gt, gs = _gpu_batch(big)
    for p in (0, 333, 699):
        lo, hi = int(big["offsets"][p]), int(big["offsets"][p + 1])
        one = dict(offsets=np.array([0, hi - lo], dtype=np.int64), uv=big["uv"][lo:hi].copy(), xyz=big["xyz"][lo:hi].copy(),
                   inlier=big["inlier"][lo:hi].copy(), intr=big["intr"][p:p + 1].copy(),
                   intr_model=big["intr_model"][p:p + 1].copy(), q=big["q"][p:p + 1].copy(), t=big["t"][p:p + 1].copy())
        q1, t1, s1 = _gpu_batch(one)
        assert np.array_equal(q1[0], gq[p]) and np.array_equal(t1[0], gt[p])
        assert s1[0] == gs[p]
    # every pose converged near its truth
    ang = 2 * np.arccos(np.minimum(1.0, np.abs(np.sum(gq * big["gt_q"], axis=1))))
    assert np.rad2deg(ang).max() < 0.5
    assert np.linalg.norm(gt - big["gt_t"], axis=1).max() < 0.5


@pytest.mark.gpu
def test_pose_edge_cases():
    from xrsfm_b200 import _lib, pnp
    # no poses at all
    assert len(pnp.refine_poses(np.zeros(1, dtype=np.int64), np.zeros((0, 2)), np.zeros((0, 3)), np.zeros((0, 8)),
                                np.zeros(0, dtype=np.int32), np.zeros((0, 4)), np.zeros((0, 3)))) == 0
    # a pose without correspondences, one whose correspondences are all masked out, and a regular one between them
    b = synth.make_pose_batch(3, seed=16, with_mask=False)
    lo1, hi1 = int(b["offsets"][1]), int(b["offsets"][2])
    offsets = np.array([0, 0, hi1 - lo1, hi1 - lo1 + 25], dtype=np.int64)
    lo2 = int(b["offsets"][2])
    uv = np.concatenate([b["uv"][lo1:hi1], b["uv"][lo2:lo2 + 25]])
    xyz = np.concatenate([b["xyz"][lo1:hi1], b["xyz"][lo2:lo2 + 25]])
    mask = np.ones(uv.shape[0], dtype=np.uint8)
    mask[hi1 - lo1:] = 0
    q, t = b["q"].copy(), b["t"].copy()
    sums = pnp.refine_poses(offsets, uv, xyz, b["intr"], b["intr_model"], q, t, inlier_mask=mask)
    for p in (0, 2):
        assert sums[p]["num_residuals"] == 0 and sums[p]["termination_type"] == 0
        assert sums[p]["num_lm_iterations"] == 0 and sums[p]["final_cost"] == 0.0
        assert np.array_equal(q[p], b["q"][p]) and np.array_equal(t[p], b["t"][p])
    one = dict(b, offsets=np.array([0, hi1 - lo1], dtype=np.int64), uv=b["uv"][lo1:hi1], xyz=b["xyz"][lo1:hi1],
               inlier=np.ones(hi1 - lo1, dtype=np.uint8), intr=b["intr"][1:2], intr_model=b["intr_model"][1:2],
               q=b["q"][1:2], t=b["t"][1:2])
    oq, ot, os_ = _oracle_batch(one)
    np.testing.assert_allclose(q[1], oq[0], rtol=0, atol=1e-8)
    np.testing.assert_allclose(t[1], ot[0], rtol=0, atol=1e-8)
    assert sums[1]["num_lm_iterations"] == os_[0]["num_lm_iterations"]
    # four or five correspondences: barely over-determined, the final cost is close to zero
    tiny = synth.make_pose_batch(4, seed=17, min_pts=4, max_pts=5, with_mask=False, outlier_frac=0.0)
    gq, gt, gs = _gpu_batch(tiny, use_mask=False)
    oq, ot, os_ = _oracle_batch(tiny)
    _compare(tiny, gq, gt, gs, oq, ot, os_, cost_rtol=1e-6, atol=1e-6)
    # bad arguments are refused, loudly
    with pytest.raises(_lib.XrbError):
        bad = b["intr_model"].copy()
        bad[0] = 9
        pnp.refine_poses(b["offsets"], b["uv"], b["xyz"], b["intr"], bad, b["q"].copy(), b["t"].copy())
    with pytest.raises(_lib.XrbError):
        off = b["offsets"].copy()
        off[1] = off[2] + 1
        pnp.refine_poses(off, b["uv"], b["xyz"], b["intr"], b["intr_model"], b["q"].copy(), b["t"].copy())
