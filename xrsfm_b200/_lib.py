"""ctypes loader for libxrsfm_b200.so (the C ABI in include/xrsfm_b200.h).

Fails loudly: there is no CPU fallback.  If the shared library is missing, or no sm_100
GPU is usable when an engine object is created, an exception is raised.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libxrsfm_b200.so")

_lib = None


class XrbError(RuntimeError):
    pass


class BAProblem(C.Structure):
    _fields_ = [
        ("n_cams", C.c_int32), ("n_pts", C.c_int32), ("n_obs", C.c_int32), ("n_intr", C.c_int32),
        ("cam_q", C.c_void_p), ("cam_t", C.c_void_p), ("pts", C.c_void_p),
        ("intr", C.c_void_p), ("intr_model", C.c_void_p), ("cam_intr", C.c_void_p),
        ("obs_cam", C.c_void_p), ("obs_pt", C.c_void_p), ("obs_uv", C.c_void_p),
        ("cam_q_fixed", C.c_void_p), ("cam_t_fixed", C.c_void_p), ("pt_fixed", C.c_void_p),
    ]


class BAOptions(C.Structure):
    _fields_ = [
        ("max_iterations", C.c_int32),
        ("function_tolerance", C.c_double), ("parameter_tolerance", C.c_double),
        ("gradient_tolerance", C.c_double), ("initial_radius", C.c_double),
        ("huber_a", C.c_double), ("min_depth", C.c_double), ("neg_depth_residual", C.c_double),
        ("verbose", C.c_int32), ("fixed_iterations", C.c_int32),
    ]


class BAIteration(C.Structure):
    _fields_ = [
        ("iteration", C.c_int32), ("step_is_valid", C.c_int32), ("step_is_successful", C.c_int32),
        ("cost", C.c_double), ("cost_change", C.c_double), ("gradient_max_norm", C.c_double),
        ("step_norm", C.c_double), ("relative_decrease", C.c_double),
        ("trust_region_radius", C.c_double), ("model_cost_change", C.c_double),
    ]


class BASummary(C.Structure):
    _fields_ = [
        ("num_residuals_reduced", C.c_int32), ("num_effective_parameters_reduced", C.c_int32),
        ("num_successful_steps", C.c_int32), ("num_unsuccessful_steps", C.c_int32),
        ("termination_type", C.c_int32),
        ("initial_cost", C.c_double), ("final_cost", C.c_double), ("fixed_cost", C.c_double),
        ("total_time_in_seconds", C.c_double),
        ("linear_solver_seconds", C.c_double), ("residual_seconds", C.c_double),
        ("n_iterations_logged", C.c_int32), ("num_lm_iterations", C.c_int32),
        ("iterations", BAIteration * 128),
    ]


class PoseSummary(C.Structure):
    _fields_ = [("num_residuals", C.c_int32), ("num_lm_iterations", C.c_int32),
                ("num_successful_steps", C.c_int32), ("num_unsuccessful_steps", C.c_int32),
                ("termination_type", C.c_int32), ("reserved", C.c_int32),
                ("initial_cost", C.c_double), ("final_cost", C.c_double)]


class ColmapSizes(C.Structure):
    _fields_ = [("n_cameras", C.c_int32), ("n_frames", C.c_int32), ("n_points", C.c_int32),
                ("n_p2d", C.c_int64), ("n_obs", C.c_int64)]


ALLREDUCE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p)

# name -> (restype, argtypes); every symbol include/xrsfm_b200.h declares
SIGNATURES = {
    "xrb_abi_version": (C.c_int, []),
    "xrb_last_error": (C.c_char_p, []),
    "xrb_kernel_launch_count": (C.c_uint64, []),
    "xrb_match_create": (C.c_void_p, [C.c_int, C.c_int]),
    "xrb_match_destroy": (None, [C.c_void_p]),
    "xrb_match_max_features": (C.c_int, [C.c_void_p]),
    "xrb_match_set_descriptors": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int]),
    "xrb_match_get": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_float, C.c_float, C.c_int]),
    "xrb_match_upload_images": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "xrb_match_upload_packed": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "xrb_match_attach_device": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "xrb_match_pairs": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_float, C.c_float, C.c_int,
                                  C.c_int, C.c_void_p, C.c_void_p, C.c_int64]),
    "xrb_match_pairs_device": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_float, C.c_float,
                                         C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                         C.c_void_p]),
    "xrb_match_debug_dist_table": (C.c_int, [C.c_void_p, C.c_int]),
    "xrb_match_set_variant": (C.c_int, [C.c_void_p, C.c_int]),
    "xrb_ba_default_options": (None, [C.POINTER(BAOptions)]),
    "xrb_ba_create": (C.c_void_p, [C.c_int]),
    "xrb_ba_destroy": (None, [C.c_void_p]),
    "xrb_ba_set_exchange": (C.c_int, [C.c_void_p, C.c_int, C.c_int, ALLREDUCE_FN, C.c_void_p]),
    "xrb_nccl_unique_id": (C.c_int, [C.c_void_p]),
    "xrb_ba_comm_init": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    "xrb_ba_shard_range": (C.c_int, [C.c_int32, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "xrb_ba_solve": (C.c_int, [C.c_void_p, C.POINTER(BAProblem), C.POINTER(BAOptions),
                               C.POINTER(BASummary)]),
    "xrb_ba_load": (C.c_int, [C.c_void_p, C.POINTER(BAProblem)]),
    "xrb_ba_reset": (C.c_int, [C.c_void_p]),
    "xrb_ba_run": (C.c_int, [C.c_void_p, C.POINTER(BAOptions), C.POINTER(BASummary), C.c_void_p]),
    "xrb_ba_fetch": (C.c_int, [C.c_void_p, C.POINTER(BAProblem)]),
    "xrb_ba_residuals": (C.c_int, [C.c_void_p, C.c_void_p]),
    "xrb_ba_solve_batch": (C.c_int, [C.c_int, C.c_int, C.c_void_p, C.POINTER(BAOptions), C.c_void_p, C.c_int]),
    "xrb_pose_default_options": (None, [C.POINTER(BAOptions)]),
    "xrb_pose_refine_batch": (C.c_int, [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(BAOptions), C.c_void_p]),
    "xrb_pose_last_kernel_ms": (C.c_double, [C.c_int]),
    "xrb_ba_filter_points3d": (C.c_int, [C.c_void_p, C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.c_void_p, C.c_void_p]),
    "xrb_ba_profile": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "xrb_ba_profile_detail": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "xrb_debug_chol_trace": (C.c_int, [C.c_int, C.c_void_p, C.c_int]),
    "xrb_debug_tile_solve": (C.c_int, [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "xrb_fm_default_options": (None, [C.c_void_p]),
    "xrb_fm_loransac_batch": (C.c_int, [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_void_p]),
    "xrb_debug_fm_samples": (C.c_int, [C.c_int, C.c_int, C.c_void_p]),
    "xrb_debug_chol_plan": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                                      C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int]),
    "xrb_debug_column_order": (C.c_int, [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "xrb_ftr_scan": (C.c_int, [C.c_char_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "xrb_ftr_read": (C.c_int, [C.c_char_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "xrb_ftr_write": (C.c_int, [C.c_char_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "xrb_match_upload_ftr": (C.c_int, [C.c_void_p, C.c_char_p]),
    "xrb_fp_scan": (C.c_int, [C.c_char_p, C.c_void_p, C.c_void_p]),
    "xrb_fp_read": (C.c_int, [C.c_char_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                              C.c_void_p, C.c_void_p]),
    "xrb_fp_write": (C.c_int, [C.c_char_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                               C.c_void_p, C.c_void_p]),
    "xrb_colmap_scan": (C.c_int, [C.c_char_p, C.c_void_p]),
    "xrb_colmap_read_problem": (C.c_int, [C.c_char_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                          C.c_void_p]),
    "xrb_colmap_write_updated": (C.c_int, [C.c_char_p, C.c_char_p, C.c_void_p, C.c_void_p]),
}


def lib():
    """Load (once) and return the ctypes handle; raise if the extension is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise XrbError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; "
                "g.build()'` (xrsfm_b200 has no CPU fallback)")
        h = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(h, name)  # AttributeError here == ABI drift, let it surface
            fn.restype = res
            fn.argtypes = args
        _lib = h
    return _lib


def last_error():
    return lib().xrb_last_error().decode("utf-8", "replace")


def check(rc, what):
    if rc != 0:
        raise XrbError(f"{what} failed with status {rc}: {last_error()}")
