"""Host mirror of the pose-refinement step of ``xrsfm::RegisterImage`` (src/geometry/pnp.cc:38-71).

The reference refines ONE frame's pose per call with a ten-iteration Ceres solve; the mapper calls it once per
registered frame.  ``refine_poses`` takes many frames at once — every pose is an independent 6-DoF problem, one CTA
each in a single kernel launch (xrsfm_b200/csrc/pose_refine.cu).  numpy arrays are host buffers only.
"""
import ctypes as C

import numpy as np

from . import _lib

TERMINATION = {0: "Convergence", 1: "No convergence", 2: "Failure"}

# xrb_pose_summary, one record per pose
SUMMARY_DTYPE = np.dtype([("num_residuals", "<i4"), ("num_lm_iterations", "<i4"), ("num_successful_steps", "<i4"),
                          ("num_unsuccessful_steps", "<i4"), ("termination_type", "<i4"), ("reserved", "<i4"),
                          ("initial_cost", "<f8"), ("final_cost", "<f8")])
assert SUMMARY_DTYPE.itemsize == C.sizeof(_lib.PoseSummary)


def make_options(**kw):
    """ceres::Solver::Options as pnp.cc:56-57 leaves them (defaults, max_num_iterations = 10) + cost constants."""
    o = _lib.BAOptions()
    _lib.lib().xrb_pose_default_options(C.byref(o))
    for k, v in kw.items():
        if not hasattr(o, k):
            raise TypeError(f"unknown option {k!r}")
        setattr(o, k, v)
    return o


def refine_poses(offsets, uv, xyz, intr, intr_model, q, t, inlier_mask=None, device=0, **opts):
    """Refine n poses in place.

    offsets[n+1] (int64) delimit each pose's correspondences in uv[total,2] (pixels) / xyz[total,3] (world);
    intr[n,8] / intr_model[n] are the (constant) cameras; q[n,4] (x,y,z,w) and t[n,3] are updated in place.
    Returns a structured array (SUMMARY_DTYPE, one ``xrb_pose_summary`` record per pose)."""
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    n = offsets.shape[0] - 1
    uv = np.ascontiguousarray(uv, dtype=np.float64)
    xyz = np.ascontiguousarray(xyz, dtype=np.float64)
    intr = np.ascontiguousarray(intr, dtype=np.float64)
    intr_model = np.ascontiguousarray(intr_model, dtype=np.int32)
    for name, a, dt, shape in (("q", q, np.float64, (n, 4)), ("t", t, np.float64, (n, 3))):
        if not (isinstance(a, np.ndarray) and a.dtype == dt and a.flags["C_CONTIGUOUS"] and a.shape == shape):
            raise TypeError(f"{name} must be a C-contiguous float64 array of shape {shape} (updated in place)")
    if intr.shape != (n, 8) or intr_model.shape != (n,):
        raise ValueError("intr must be [n, 8] and intr_model [n]")
    total = int(offsets[-1]) if n >= 0 else 0
    if uv.shape != (total, 2) or xyz.shape != (total, 3):
        raise ValueError("uv must be [total, 2] and xyz [total, 3] with total = offsets[-1]")
    mask_ptr = None
    if inlier_mask is not None:
        inlier_mask = np.ascontiguousarray(inlier_mask, dtype=np.uint8)
        if inlier_mask.shape != (total,):
            raise ValueError("inlier_mask must be [total]")
        mask_ptr = inlier_mask.ctypes.data
    o = make_options(**opts)
    sums = np.zeros(n, dtype=SUMMARY_DTYPE)
    _lib.check(_lib.lib().xrb_pose_refine_batch(device, n, offsets.ctypes.data, uv.ctypes.data, xyz.ctypes.data,
                                                mask_ptr, intr.ctypes.data, intr_model.ctypes.data, q.ctypes.data,
                                                t.ctypes.data, C.byref(o), sums.ctypes.data), "xrb_pose_refine_batch")
    return sums


def last_kernel_ms(device=0):
    """Device time of the kernel of the last refine_poses call on `device` (CUDA events)."""
    return _lib.lib().xrb_pose_last_kernel_ms(device)
