"""ctypes bindings of oracle/liboracle.so and oracle/_ref/libxrref_match.so.

Test infrastructure: only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline /
--impl reference legs import this module."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "liboracle.so")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libxrref_match.so")

_lib = None
_ref = None


def build():
    subprocess.check_call(["make", "-C", ORACLE_DIR, "-s", "liboracle.so"])


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(ORACLE_SO):
            build()
        h = C.CDLL(ORACLE_SO)
        h.xro_dist_of_dot.restype = C.c_float
        h.xro_dist_of_dot.argtypes = [C.c_int]
        h.xro_accept.restype = C.c_int
        h.xro_accept.argtypes = [C.c_int, C.c_int, C.c_float, C.c_float]
        h.xro_dot_matrix.restype = None
        h.xro_dot_matrix.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        h.xro_match_pair.restype = C.c_int
        h.xro_match_pair.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_float, C.c_float,
                                     C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        h.xro_match_pairs.restype = C.c_int
        h.xro_match_pairs.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_float,
                                      C.c_float, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int64]
        _lib = h
    return _lib


def match_pair(d1, d2, distmax=0.7, ratiomax=0.8, mbm=1, max_match=16384, want_m=False):
    """Oracle for one pair -> matches [n,2] (and m12, m21 when want_m)."""
    h = load()
    d1 = np.ascontiguousarray(d1, dtype=np.uint8)
    d2 = np.ascontiguousarray(d2, dtype=np.uint8)
    n1, n2 = d1.shape[0], d2.shape[0]
    out = np.zeros((max(1, min(max_match, max(n1, 1))), 2), dtype=np.uint32)
    m12 = np.full(max(n1, 1), -2, dtype=np.int32)
    m21 = np.full(max(n2, 1), -2, dtype=np.int32)
    n = h.xro_match_pair(n1, d1.ctypes.data, n2, d2.ctypes.data, distmax, ratiomax, mbm, max_match,
                         out.ctypes.data, m12.ctypes.data, m21.ctypes.data)
    if want_m:
        return out[:n].copy(), m12[:n1], m21[:n2]
    return out[:n].copy()


def load_ref():
    """The reference's own CUDA kernels (GPU box only)."""
    global _ref
    if _ref is None:
        if not os.path.exists(REF_SO):
            return None
        h = C.CDLL(REF_SO)
        h.xrref_match_pair.restype = C.c_int
        h.xrref_match_pair.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_float, C.c_float,
                                       C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        _ref = h
    return _ref


def ref_match_pair(d1, d2, distmax=0.7, ratiomax=0.8, mbm=1, max_match=16384, want_m=False):
    h = load_ref()
    d1 = np.ascontiguousarray(d1, dtype=np.uint8)
    d2 = np.ascontiguousarray(d2, dtype=np.uint8)
    n1, n2 = d1.shape[0], d2.shape[0]
    out = np.zeros((max(1, min(max_match, max(n1, 1))), 2), dtype=np.uint32)
    m12 = np.full(max(n1, 1), -2, dtype=np.int32)
    m21 = np.full(max(n2, 1), -2, dtype=np.int32)
    n = h.xrref_match_pair(n1, d1.ctypes.data, n2, d2.ctypes.data, distmax, ratiomax, mbm, max_match,
                           out.ctypes.data, m12.ctypes.data, m21.ctypes.data)
    if n < 0:
        raise RuntimeError("reference kernels returned -1 (CUDA error)")
    if want_m:
        return out[:n].copy(), m12[:n1], m21[:n2]
    return out[:n].copy()


# ------------------------------------------------------------------------------------------
# BA oracle
# ------------------------------------------------------------------------------------------
def _ba_types():
    from xrsfm_b200 import _lib as L
    return L


def ba_options(**kw):
    """xrb_ba_options with Ceres defaults + the reference's constants; kw overrides."""
    L = _ba_types()
    o = L.BAOptions()
    h = load()
    h.xro_ba_default_options.restype = None
    h.xro_ba_default_options(C.byref(o))
    for k, v in kw.items():
        setattr(o, k, v)
    return o


GBA_ACCURATE = dict(max_iterations=50, function_tolerance=1e-5, parameter_tolerance=1e-6)  # ba_solver.cc:626-629
GBA_FAST = dict(max_iterations=20, function_tolerance=1e-4, parameter_tolerance=1e-5)      # :630-634
KGBA = dict(max_iterations=20, function_tolerance=1e-4, parameter_tolerance=1e-5,
            initial_radius=1e6)                                                            # :667-670


def ba_problem(scene):
    """ctypes xrb_ba_problem aliasing the numpy arrays of a synth.BAScene (kept alive by scene)."""
    L = _ba_types()
    p = L.BAProblem()
    p.n_cams, p.n_pts, p.n_obs, p.n_intr = scene.n_cams, scene.n_pts, scene.n_obs, scene.n_intr
    for name in ("cam_q", "cam_t", "pts", "intr", "intr_model", "cam_intr", "obs_cam", "obs_pt",
                 "obs_uv", "cam_q_fixed", "cam_t_fixed", "pt_fixed"):
        a = scene[name]
        assert a.flags["C_CONTIGUOUS"], name
        setattr(p, name, a.ctypes.data)
    return p


def ba_solve(scene, opts, n_threads=0):
    """Run the oracle in place on `scene`; returns the summary struct."""
    L = _ba_types()
    h = load()
    h.xro_ba_solve.restype = C.c_int
    h.xro_ba_solve.argtypes = [C.POINTER(L.BAProblem), C.POINTER(L.BAOptions), C.POINTER(L.BASummary),
                               C.c_int]
    s = L.BASummary()
    p = ba_problem(scene)
    rc = h.xro_ba_solve(C.byref(p), C.byref(opts), C.byref(s), n_threads)
    assert rc == 0
    return s


def ba_residuals(scene, opts):
    L = _ba_types()
    h = load()
    h.xro_ba_residuals.restype = C.c_int
    h.xro_ba_residuals.argtypes = [C.POINTER(L.BAProblem), C.POINTER(L.BAOptions), C.c_void_p]
    out = np.zeros((scene.n_obs, 2))
    p = ba_problem(scene)
    h.xro_ba_residuals(C.byref(p), C.byref(opts), out.ctypes.data)
    return out


def ba_cost(scene, opts):
    L = _ba_types()
    h = load()
    h.xro_ba_cost.restype = C.c_double
    h.xro_ba_cost.argtypes = [C.POINTER(L.BAProblem), C.POINTER(L.BAOptions)]
    p = ba_problem(scene)
    return h.xro_ba_cost(C.byref(p), C.byref(opts))


def ba_eval_obs(q, t, X, model, intr, uv, opts, robustify=False):
    """-> (r[2], Jd[2,3], Jt[2,3], JX[2,3], rho0, depth_branch)"""
    L = _ba_types()
    h = load()
    h.xro_ba_eval_obs.restype = C.c_int
    h.xro_ba_eval_obs.argtypes = [C.c_void_p] * 3 + [C.c_int, C.c_void_p, C.c_void_p, C.POINTER(L.BAOptions),
                                                     C.c_int, C.c_void_p]
    arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in (q, t, X, intr, uv)]
    intr8 = np.zeros(8)
    intr8[: len(arrs[3])] = arrs[3]
    out = np.zeros(21)
    br = h.xro_ba_eval_obs(arrs[0].ctypes.data, arrs[1].ctypes.data, arrs[2].ctypes.data, model,
                           intr8.ctypes.data, arrs[4].ctypes.data, C.byref(opts), int(robustify),
                           out.ctypes.data)
    return out[:2], out[2:8].reshape(2, 3), out[8:14].reshape(2, 3), out[14:20].reshape(2, 3), out[20], br


def quat_plus(q, d):
    h = load()
    h.xro_quat_plus.restype = None
    h.xro_quat_plus.argtypes = [C.c_void_p] * 3
    q = np.ascontiguousarray(q, dtype=np.float64)
    d = np.ascontiguousarray(d, dtype=np.float64)
    out = np.zeros(4)
    h.xro_quat_plus(q.ctypes.data, d.ctypes.data, out.ctypes.data)
    return out


def summary_dict(s):
    its = []
    for i in range(s.n_iterations_logged):
        it = s.iterations[i]
        its.append({k: getattr(it, k) for k, _ in it._fields_})
    d = {k: getattr(s, k) for k, _ in s._fields_ if k != "iterations"}
    d["iterations"] = its
    return d
