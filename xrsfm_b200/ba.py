"""Host mirror of the reference's bundle-adjustment surface (path B).

``BASolver`` keeps the names and option sets of ``xrsfm::BASolver``
(src/optimization/ba_solver.h:14-30):

* ``GBA(problem, accurate=True, fix_all_frames=False)``  — ba_solver.cc:594-638
* ``KGBA(problem)``                                      — ba_solver.cc:640-678 (solver part)
* ``LBA``-style problems are expressed through ``pt_fixed`` / ``cam_t_fixed`` masks
  (ba_solver.cc:358-391, 552-584) and ``solve(..., max_iterations=5, ...)``.

The reference mutates an AoS ``Map``; here the problem is the flat SoA the C ABI takes
(``xrb_ba_problem``, SURVEY.md Appendix B describes the Map -> SoA flattening a C++ shim
performs — see xrsfm_b200/shim/ba_solver_b200.h).  All arithmetic runs in
``libxrsfm_b200.so``; numpy arrays are host buffers only.
"""
import ctypes as C

import numpy as np

from . import _lib

# option sets: ba_solver.cc:626-634 and :667-670 on top of InitSolverOptions (:70-77)
GBA_ACCURATE = dict(max_iterations=50, function_tolerance=1e-5, parameter_tolerance=1e-6)
GBA_FAST = dict(max_iterations=20, function_tolerance=1e-4, parameter_tolerance=1e-5)
KGBA_OPTIONS = dict(max_iterations=20, function_tolerance=1e-4, parameter_tolerance=1e-5,
                    initial_radius=1e6)
LBA_OPTIONS = dict(max_iterations=5, function_tolerance=1e-4, parameter_tolerance=1e-5)  # :586-590

TERMINATION = {0: "Convergence", 1: "No convergence", 2: "Failure"}  # ba_solver.cc:44-65

_FIELDS = ("cam_q", "cam_t", "pts", "intr", "intr_model", "cam_intr", "obs_cam", "obs_pt",
           "obs_uv", "cam_q_fixed", "cam_t_fixed", "pt_fixed")
_DTYPES = dict(cam_q=np.float64, cam_t=np.float64, pts=np.float64, intr=np.float64,
               intr_model=np.int32, cam_intr=np.int32, obs_cam=np.int32, obs_pt=np.int32,
               obs_uv=np.float64, cam_q_fixed=np.uint8, cam_t_fixed=np.uint8, pt_fixed=np.uint8)


def make_options(**kw):
    o = _lib.BAOptions()
    _lib.lib().xrb_ba_default_options(C.byref(o))
    for k, v in kw.items():
        if not hasattr(o, k):
            raise TypeError(f"unknown BA option {k!r}")
        setattr(o, k, v)
    return o


def make_problem(scene):
    """xrb_ba_problem aliasing the arrays of `scene` (mapping with the xrb_ba_problem field
    names; arrays must be C-contiguous with the documented dtypes — they are updated in place)."""
    p = _lib.BAProblem()
    p.n_cams, p.n_pts = int(scene["n_cams"]), int(scene["n_pts"])
    p.n_obs, p.n_intr = int(scene["n_obs"]), int(scene["n_intr"])
    for name in _FIELDS:
        a = scene[name]
        if a is None:
            setattr(p, name, None)
            continue
        if not (isinstance(a, np.ndarray) and a.flags["C_CONTIGUOUS"] and a.dtype == _DTYPES[name]):
            raise TypeError(f"{name}: need a C-contiguous {np.dtype(_DTYPES[name]).name} array")
        setattr(p, name, a.ctypes.data)
    return p


def print_solver_summary(s):
    """PrintSolverSummary (ba_solver.cc:14-68), same fields and layout."""
    n = max(s.num_residuals_reduced, 1)
    print(f"{'Residuals : ':>16}{s.num_residuals_reduced}")
    print(f"{'Parameters : ':>16}{s.num_effective_parameters_reduced}")
    print(f"{'Iterations : ':>16}{s.num_successful_steps + s.num_unsuccessful_steps}")
    print(f"{'Time : ':>16}{s.total_time_in_seconds:g} [s]")
    print(f"{'Initial cost : ':>16}{np.sqrt(s.initial_cost / n):.6g} [px]")
    print(f"{'Final cost : ':>16}{np.sqrt(s.final_cost / n):.6g} [px]")
    print(f"{'Termination : ':>16}{TERMINATION.get(s.termination_type, 'Unknown')}\n")


class BASolver:
    """B200 engine behind the reference's BASolver call surface."""

    def __init__(self, device=0):
        self._device = device
        self._h = None
        self._hook = None  # keep the ctypes callback alive

    # -- reference-named entry points -----------------------------------------------------
    def GBA(self, scene, accurate=True, fix_all_frames=False, verbose=False):
        """ba_solver.cc:594-638. Gauge: translations of the first two cameras are expected
        fixed in scene['cam_t_fixed'] (init_id1/init_id2, :611-614); fix_all_frames freezes
        every pose (:616-621)."""
        sc = scene
        if fix_all_frames:
            sc = dict(scene)
            sc["cam_q_fixed"] = np.ones(scene["n_cams"], dtype=np.uint8)
            sc["cam_t_fixed"] = np.ones(scene["n_cams"], dtype=np.uint8)
        s = self.solve_scene(sc, **(GBA_ACCURATE if accurate else GBA_FAST))
        if verbose:
            print_solver_summary(s)
        return s

    def KGBA(self, scene, verbose=False):
        """Solver part of ba_solver.cc:640-678 (key-frame selection and UpdateByRefFrame stay
        with the caller, which passes only key-frames as cameras)."""
        s = self.solve_scene(scene, **KGBA_OPTIONS)
        if verbose:
            print_solver_summary(s)
        return s

    def LBA(self, scene, verbose=False):
        """Local BA option set (ba_solver.cc:586-591); the caller supplies the local window and
        the constant-point / constant-translation masks (:380-382, :552-584)."""
        s = self.solve_scene(scene, **LBA_OPTIONS)
        if verbose:
            print_solver_summary(s)
        return s

    # -- flat interface ------------------------------------------------------------------
    def solve_scene(self, scene, **opts):
        self._ensure()
        p = make_problem(scene)
        o = make_options(**opts)
        s = _lib.BASummary()
        _lib.check(_lib.lib().xrb_ba_solve(self._h, C.byref(p), C.byref(o), C.byref(s)), "xrb_ba_solve")
        return s

    def load(self, scene):
        self._ensure()
        self._problem = make_problem(scene)
        self._scene = scene
        _lib.check(_lib.lib().xrb_ba_load(self._h, C.byref(self._problem)), "xrb_ba_load")

    def reset(self):
        _lib.check(_lib.lib().xrb_ba_reset(self._h), "xrb_ba_reset")

    def run(self, stream=None, **opts):
        o = make_options(**opts)
        s = _lib.BASummary()
        _lib.check(_lib.lib().xrb_ba_run(self._h, C.byref(o), C.byref(s), stream), "xrb_ba_run")
        return s

    def fetch(self):
        _lib.check(_lib.lib().xrb_ba_fetch(self._h, C.byref(self._problem)), "xrb_ba_fetch")

    def residuals(self):
        out = np.zeros((self._problem.n_obs, 2))
        _lib.check(_lib.lib().xrb_ba_residuals(self._h, out.ctypes.data), "xrb_ba_residuals")
        return out

    def profile(self):
        ms = (C.c_double * 6)()
        ln = (C.c_int64 * 6)()
        _lib.check(_lib.lib().xrb_ba_profile(self._h, ms, ln), "xrb_ba_profile")
        names = ("schur", "solve", "backsub", "cost", "exchange", "run")
        return {n: (ms[i], ln[i]) for i, n in enumerate(names)}

    @staticmethod
    def solve_batch(scenes, device=0, n_workers=8, **opts):
        """xrb_ba_solve_batch: independent (local-BA sized) problems solved concurrently on one device."""
        probs = (_lib.BAProblem * len(scenes))(*[make_problem(sc) for sc in scenes])
        sums = (_lib.BASummary * len(scenes))()
        o = make_options(**opts)
        _lib.check(_lib.lib().xrb_ba_solve_batch(device, len(scenes), probs, C.byref(o), sums, n_workers),
                   "xrb_ba_solve_batch")
        return list(sums)

    def filter_points3d(self, max_re, deg):
        """Point3dProcessor::FilterPoints3d (track_processor.cc:321-349) on the solver's current state.
        Returns (keep_obs, pt_outlier, pt_error, pt_angle, (num_filtered1, num_filtered2))."""
        n_obs, n_pts = int(self._problem.n_obs), int(self._problem.n_pts)
        keep = np.zeros(max(1, n_obs), dtype=np.uint8)
        out = np.zeros(max(1, n_pts), dtype=np.uint8)
        err = np.zeros(max(1, n_pts))
        ang = np.zeros(max(1, n_pts))
        counts = np.zeros(2, dtype=np.int32)
        _lib.check(_lib.lib().xrb_ba_filter_points3d(self._h, float(max_re), float(deg), keep.ctypes.data, out.ctypes.data,
                                                     err.ctypes.data, ang.ctypes.data, counts.ctypes.data),
                   "xrb_ba_filter_points3d")
        return keep[:n_obs], out[:n_pts], err[:n_pts], ang[:n_pts], (int(counts[0]), int(counts[1]))

    def profile_detail(self):
        out = (C.c_double * 20)()
        _lib.check(_lib.lib().xrb_ba_profile_detail(self._h, out, 20), "xrb_ba_profile_detail")
        keys = ("lin_ms", "gather_ms", "cam_blocks_ms", "solves", "n_blocks", "n_incidences", "nc", "half_bandwidth",
                "parts", "tile_columns", "tiles", "tiles_original", "plan_flops", "depth_factor", "depth_backward",
                "chains", "schur_window_ctas", "schur_window_stride", "camera_span", "longest_track")
        return dict(zip(keys, list(out)))

    def comm_init(self, rank, world, broadcast_bytes):
        """Native multi-GPU exchange (xrb_ba_comm_init): the library owns an NCCL communicator and
        all-reduces on the solver's stream.  `broadcast_bytes(buf: bytearray | None) -> bytes` ships
        rank 0's 128-byte NCCL id to every rank (any transport: torch.distributed, MPI, a file)."""
        self._ensure()
        ident = (C.c_uint8 * 128)()
        if rank == 0:
            _lib.check(_lib.lib().xrb_nccl_unique_id(ident), "xrb_nccl_unique_id")
        raw = broadcast_bytes(bytes(ident) if rank == 0 else None)
        ident = (C.c_uint8 * 128).from_buffer_copy(raw)
        _lib.check(_lib.lib().xrb_ba_comm_init(self._h, ident, rank, world), "xrb_ba_comm_init")

    def set_exchange(self, rank, world, allreduce):
        """allreduce(ptr: int, count: int, stream: int) -> None: in-place SUM over ranks of `count`
        doubles at device address `ptr`, ordered on the CUDA stream `stream` (see
        include/xrsfm_b200.h xrb_ba_set_exchange)."""
        self._ensure()

        def _cb(buf, count, stream, _user):
            try:
                allreduce(int(buf), int(count), int(stream or 0))
                return 0
            except Exception as e:  # noqa: BLE001 - surfaced as XRB_ERR_COMM
                print("exchange hook failed:", e)
                return 1

        self._hook = _lib.ALLREDUCE_FN(_cb)
        _lib.check(_lib.lib().xrb_ba_set_exchange(self._h, rank, world, self._hook, None),
                   "xrb_ba_set_exchange")

    # -- plumbing ------------------------------------------------------------------------
    def _ensure(self):
        if self._h is None:
            h = _lib.lib().xrb_ba_create(self._device)
            if not h:
                raise _lib.XrbError("xrb_ba_create failed: " + _lib.last_error())
            self._h = h

    def __del__(self):
        try:
            if self._h is not None:
                _lib.lib().xrb_ba_destroy(self._h)
                self._h = None
        except Exception:
            pass
