#!/usr/bin/env python
"""bench.py — headline benchmark of xrsfm_b200 (contract: see the task statement / DESIGN.md §4).

Headline line (BASELINE.json configs[1], C2): bundle adjustment on the synthetic
500-camera / 200k-point / 2M-observation scene, metric = LM iterations per second.
One "step" = one Levenberg–Marquardt iteration (linear solve through the Schur complement +
candidate evaluation); the timed region runs EXACTLY K of them on state that
is already resident in HBM (`xrb_ba_run`, fixed_iterations), bracketed by CUDA events on the
stream the kernels are launched on.  `e2e` times the reference-facing C-ABI call
`xrb_ba_solve` with HOST buffers (upload + K iterations + download).

The same JSON line carries a `matching` object with the second hot path (BASELINE config C3
per-pair size: 4096 x 4096 x 128-D uint8 descriptors, pairs/s).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--path ba|match]

N > 1 is launched by torchrun (one rank per GPU): BA shards the points and all-reduces the
reduced camera system once per linear solve (strong scaling); matching shards the pair list
(no collective).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured"
    return 6650.0, "fallback"  # B200_PROFILING.md fallback


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])), mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except (ValueError, IndexError):
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def host_cores():
    """CPU threads this process may really use: affinity mask and cgroup quota, not the
    box's core count (the GPU box reports 128 cores but runs jobs under a quota)."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        q, per = open("/sys/fs/cgroup/cpu.max").read().split()
        if q != "max":
            n = max(1, min(n, int(int(q) / int(per))))
    except (OSError, ValueError):
        pass
    return n


def dist_env():
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if os.environ.get("XRB_BENCH_ONE_GPU"):  # dry run of the N > 1 control flow on a 1-GPU box (gloo)
        local_rank = 0
    return rank, world, local_rank


class _CudaPtr:
    """Zero-copy view of a raw device pointer for torch (via __cuda_array_interface__)."""

    def __init__(self, ptr, count):
        self.__cuda_array_interface__ = {"shape": (count,), "typestr": "<f8", "data": (ptr, False), "version": 2}


# ------------------------------------------------------------------------------------------
# BA arm
# ------------------------------------------------------------------------------------------
_SCENE_KEYS = ("cam_q", "cam_t", "pts", "intr", "intr_model", "cam_intr", "obs_cam", "obs_pt", "obs_uv",
               "cam_q_fixed", "cam_t_fixed", "pt_fixed")

# option sets of the reference for each configuration
GBA_ACCURATE = dict(function_tolerance=1e-5, parameter_tolerance=1e-6)                    # ba_solver.cc:626-629
KGBA = dict(function_tolerance=1e-4, parameter_tolerance=1e-5, initial_radius=1e6)         # ba_solver.cc:667-670
BA_CONFIGS = {
    "C2": dict(options=GBA_ACCURATE, max_iterations=50,
               workload="C2: synthetic 500-camera / 200k-point / 2M-obs global BA (GBA accurate options), "
                        "SIMPLE_RADIAL, Huber 5.99, 2 translations fixed"),
    "C4": dict(options=KGBA, max_iterations=20,
               workload="C4: KITTI-shaped sequential scene, 2.7k cams / 1M pts / 10M obs (KGBA options: radius 1e6, "
                        "<= 20 iterations), banded reduced camera system"),
    "C5": dict(options=GBA_ACCURATE, max_iterations=50,
               workload="C5: 1DSfM-shaped clustered scene, 5k cams (one camera model per image) / 1.5M pts / 12M obs, "
                        "global BA (GBA accurate options)"),
}


def make_scene_cached(name, scale=1.0):
    from xrsfm_b200 import synth
    cache = f"/tmp/xrsfm_b200_{name}_{scale}.npz"
    if os.path.exists(cache):
        try:
            z = np.load(cache)
            sc = synth.BAScene({k: np.ascontiguousarray(z[k]) for k in _SCENE_KEYS})
            sc.n_cams, sc.n_pts, sc.n_obs, sc.n_intr = (int(z["dims"][i]) for i in range(4))
            return sc
        except Exception:
            pass
    sc = synth.make_scene(name, scale)
    try:  # atomic publish: several ranks may build the scene at the same time
        tmp = f"{cache}.{os.getpid()}.tmp.npz"
        np.savez(tmp, dims=np.array([sc.n_cams, sc.n_pts, sc.n_obs, sc.n_intr]), **{k: sc[k] for k in _SCENE_KEYS})
        os.replace(tmp, cache)
    except Exception:
        pass
    return sc


def make_c2(scale=1.0):
    return make_scene_cached("C2", scale)


def ba_gather_unique_bytes(detail, n_obs):
    """UNIQUE HBM bytes of one k_gather launch: every 144-byte observation record read once, the
    incidence index pairs (8 B) and the block table (12 B) read once, every 6x6 block of S written once.
    The kernel requests each record k_p - 1 times; the re-reads are L2 business, not algorithmic bytes."""
    return int(144 * n_obs + 8 * detail["n_incidences"] + (288 + 12) * detail["n_blocks"])


def ba_lin_bytes(sc):
    """k_lin: observation stream 24 B/obs read + one 144-byte record written per observation;
    per point 24 B read and 96 B (V^-1, g, h) written."""
    return 168 * sc.n_obs + 120 * sc.n_pts


def ba_iteration_bytes(sc, detail):
    """bytes per LM iteration = 72 O + 96 P + 8 (2 nnz(S) + 12 C)."""
    nnz = 36.0 * detail["n_blocks"] + 21.0 * sc.n_cams
    return 72.0 * sc.n_obs + 96.0 * sc.n_pts + 8.0 * (2.0 * nnz + 12.0 * sc.n_cams)


def ba_chol_flops(nc, bw):
    """Cholesky + forward/backward substitution, banded: n*bw^2 (dense: n^3/3) FMAs x 2."""
    if bw >= nc - 1:
        return 2.0 * (nc ** 3 / 6.0 + nc ** 2)
    return 2.0 * (nc * bw * bw / 2.0 + 2.0 * nc * bw)


def load_fp64_peak():
    """Measured FP64 FMA-pipe peak of this pool's B200 (tools/fp64_peak.cu, committed copy)."""
    p = os.path.join(ROOT, "profiles", "r02_fp64_peak.json")
    try:
        d = json.load(open(p))
        return float(d["dfma_tflops"]), float(d["dmma_m8n8k4_tflops"]), "measured (profiles/r02_fp64_peak.json)"
    except Exception:
        return 37.0, 37.0, "nominal"


def setup_exchange(solver, rank, world, local_rank):
    """Multi-GPU: the library's own NCCL communicator (xrb_ba_comm_init); torch.distributed only
    ships the 128-byte id."""
    if world <= 1:
        return
    import torch
    import torch.distributed as dist

    def bcast(raw):
        t = torch.zeros(128, dtype=torch.uint8, device=f"cuda:{local_rank}")
        if rank == 0:
            t.copy_(torch.frombuffer(bytearray(raw), dtype=torch.uint8))
        dist.broadcast(t, 0)
        return bytes(t.cpu().numpy().tobytes())

    solver.comm_init(rank, world, bcast)


def parity_check(sc, cfg, n_iters, gpu_log):
    """Outside every timed region: the first iterations of the CPU oracle on the same scene against
    the GPU's iteration log (cost 1e-7 relative, same accept/reject decisions)."""
    from tests import oracle_lib as ol
    ref = sc.copy_state()
    s = ol.ba_solve(ref, ol.ba_options(max_iterations=n_iters, **cfg["options"]), host_cores())
    n = min(s.n_iterations_logged, len(gpu_log), n_iters + 1)
    worst = 0.0
    for i in range(n):
        a, b = s.iterations[i], gpu_log[i]
        if a.step_is_successful != b["step_is_successful"]:
            return {"parity_checked": False, "why": f"accept/reject differs at iteration {i}"}
        worst = max(worst, abs(a.cost - b["cost"]) / max(abs(a.cost), 1e-300))
    return {"parity_checked": bool(worst < 1e-7), "iterations_compared": n, "max_rel_cost_diff": worst,
            "against": "CPU oracle (oracle/ba_oracle.cpp), same scene and options"}


def run_ba(args, rank, world, local_rank, name="C2"):
    import torch
    from xrsfm_b200 import _lib, ba
    cfg = BA_CONFIGS[name]
    opts = cfg["options"]
    torch.cuda.set_device(local_rank)
    sc = make_scene_cached(name, args.scale)
    solver = ba.BASolver(device=local_rank)
    solver._ensure()
    setup_exchange(solver, rank, world, local_rank)
    lib = _lib.lib()
    stream = torch.cuda.current_stream().cuda_stream

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if world <= 1:
            return v
        import torch.distributed as dist
        t = torch.tensor([v], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # -- resident path: load once, W warm-up iterations, then exactly K timed iterations
    solver.load(sc)
    if args.warmup > 0:
        solver.run(stream=stream, max_iterations=args.warmup, fixed_iterations=1, **opts)
    solver.reset()
    clocks = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        clocks.start()
    launches0 = lib.xrb_kernel_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    summ = solver.run(stream=stream, max_iterations=args.steps, fixed_iterations=1, **opts)
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    launches = lib.xrb_kernel_launch_count() - launches0
    clk = clocks.stop() if rank == 0 else None
    prof = solver.profile()
    detail = solver.profile_detail()
    assert summ.num_lm_iterations == args.steps, (summ.num_lm_iterations, args.steps)

    # -- to-convergence run with the reference's option set (reported; its log feeds the parity check)
    solver.reset()
    conv = solver.run(stream=stream, max_iterations=cfg["max_iterations"], **opts)
    conv_log = [{"cost": conv.iterations[i].cost, "step_is_successful": conv.iterations[i].step_is_successful}
                for i in range(conv.n_iterations_logged)]
    barrier()

    # -- e2e: host buffers through xrb_ba_solve (upload + structure build + K iterations + download)
    def pinned_scene():
        """The scene in PINNED host memory (what the contract's e2e copies from): same arrays, page-locked."""
        w = sc.copy_state()
        for k in _SCENE_KEYS:
            a = np.ascontiguousarray(sc[k])
            t = torch.empty(a.shape, dtype=torch.from_numpy(a[:0].copy()).dtype).pin_memory()
            t.copy_(torch.from_numpy(a))
            w[k] = t.numpy()
            w.setdefault("_pins", []).append(t)  # keep the pinned tensors alive
        return w

    work = pinned_scene()
    h2d = sum(sc[k].nbytes for k in _SCENE_KEYS)
    d2h = sum(sc[k].nbytes for k in ("cam_q", "cam_t", "pts"))
    barrier()
    t0 = time.perf_counter()
    s_e2e = solver.solve_scene(work, max_iterations=args.steps, fixed_iterations=1, **opts)
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    # the same call for a REAL solve: to convergence, nothing amortised
    work2 = pinned_scene()
    barrier()
    t0 = time.perf_counter()
    s_full = solver.solve_scene(work2, max_iterations=cfg["max_iterations"], **opts)
    barrier()
    full_s = max_over_ranks(time.perf_counter() - t0)
    if rank != 0:
        return None

    nsolve = max(1.0, detail["solves"])
    kern_ms = {"k_lin": detail["lin_ms"] / nsolve, "k_gather": detail["gather_ms"] / nsolve,
               "k_cam_blocks_exposed": detail["cam_blocks_ms"] / nsolve, "tile_cholesky+backsolve": prof["solve"][0] / nsolve}
    per_it = {k: v[0] / nsolve for k, v in prof.items() if k != "run"}
    out = {
        "metric": "BA LM-iterations/sec", "value": args.steps / (ms * 1e-3), "unit": "LM-iterations/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": cfg["workload"], "name": name,
                   "n_cams": sc.n_cams, "n_pts": sc.n_pts, "n_obs": sc.n_obs, "scale": args.scale,
                   "parallelism": f"points sharded over {world} GPU(s), cameras replicated, one in-library "
                                  "ncclAllReduce(ncclDouble, ncclSum) of the packed reduced camera system per solve",
                   "l2_note": "no explicit L2 flush: every iteration streams a working set larger than the 126 MB L2 "
                              "(observation records 144 B/obs + observations 24 B/obs + reduced system + points)",
                   "reduced_system": {"dimension": int(detail["nc"]), "half_bandwidth": int(detail["half_bandwidth"]),
                                      "blocks": int(detail["n_blocks"])},
                   "iterations_to_convergence": conv.num_lm_iterations,
                   "termination": ba.TERMINATION.get(conv.termination_type),
                   "final_rms_px": float(np.sqrt(conv.final_cost / max(1, conv.num_residuals_reduced))),
                   "phase_ms_per_iteration": per_it},
        "e2e": {"value": args.steps / e2e_s, "unit": "LM-iterations/s", "h2d_bytes_per_step": h2d / args.steps,
                "d2h_bytes_per_step": d2h / args.steps, "call": "xrb_ba_solve (host buffers in/out)",
                "iterations": s_e2e.num_lm_iterations,
                "to_convergence": {"value": s_full.num_lm_iterations / full_s, "unit": "LM-iterations/s",
                                   "iterations": s_full.num_lm_iterations, "seconds": full_s,
                                   "note": "one real solve through xrb_ba_solve: upload + structure build + every "
                                           "iteration to the reference's tolerances + download, nothing amortised"}},
        "gpu_launches": int(launches),
        "clocks": clk,
        "kernel_ms_per_solve": kern_ms,
        "_scene": sc, "_detail": detail, "_conv_log": conv_log, "_cfg": cfg,
    }
    return out


def ba_rooflines(out):
    """roofline = the step's dominant kernel; HBM figures use UNIQUE bytes (never more than DRAM traffic)."""
    sc, detail, kern_ms = out["_scene"], out["_detail"], out["kernel_ms_per_solve"]
    hbm_peak, peak_src = load_peaks()
    dfma, dmma, fsrc = load_fp64_peak()
    step_ms = out["ms_per_step"]
    chol_ms = kern_ms["tile_cholesky+backsolve"]
    chol_tf = ba_chol_flops(detail["nc"], detail["half_bandwidth"]) / (chol_ms * 1e-3) / 1e12
    gat_b = ba_gather_unique_bytes(detail, sc.n_obs)
    gat_ach = gat_b / (kern_ms["k_gather"] * 1e-3) / 1e9
    it_b = ba_iteration_bytes(sc, detail)
    roof = {"kernel": "k_tile_cholesky + k_tile_backsolve (sparse tile Cholesky of the reduced camera system: task DAG "
                      "over resident CTAs, tile products on the FP64 tensor pipe)",
            "bound": "tensor", "bound_note": "FP64 tensor pipe (mma.sync.m8n8k4.f64, DMMA) for the tile products; the "
                                             "diagonal chain (POTRF of 64 x 64 tiles) is dependency-latency bound",
            "achieved": chol_tf, "peak": dmma, "peak_source": fsrc, "unit": "TFLOP/s", "frac": chol_tf / dmma,
            "share_of_step": chol_ms / step_ms, "ms_per_launch": chol_ms, "traffic": None,
            "algorithmic_flops_per_launch": ba_chol_flops(detail["nc"], detail["half_bandwidth"]),
            "executed_flops_per_launch": detail["plan_flops"],
            "plan": {"column_order_parts": int(detail["parts"]), "tile_columns": int(detail["tile_columns"]),
                     "tiles": int(detail["tiles"]), "tiles_original": int(detail["tiles_original"]),
                     "longest_dependency_path_tasks": int(detail["depth_factor"]), "chains": int(detail["chains"])},
            "tensor_pipe_note": f"FP64 tensor (DMMA m8n8k4) peak measured at {dmma:.1f} TFLOP/s, DFMA {dfma:.1f} on this part; "
                                "algorithmic flops = n^3/3 (dense) or n bw^2 (band) + substitutions, the fill of a "
                                "dissected band is not counted"}
    out["roofline"] = roof
    out["roofline_hbm"] = {
        "kernel": "k_gather (Schur complement: per-block gather of the observation records)", "bound": "hbm",
        "achieved": gat_ach, "peak": hbm_peak, "peak_source": peak_src, "unit": "GB/s", "frac": gat_ach / hbm_peak,
        "share_of_step": kern_ms["k_gather"] / step_ms, "ms_per_launch": kern_ms["k_gather"],
        "algorithmic_bytes_per_launch": gat_b, "traffic": None,
        "traffic_note": "measured dram__bytes of this kernel: see the ncu summary under profiles/ (r01c: 1.355e9 B per "
                        "launch at C2 against 0.41e9 unique: the records are re-read k_p - 1 times, about half from L2)"}
    out["roofline_iteration"] = {
        "bound": "hbm", "unit": "GB/s", "bytes_per_iteration_survey_8d": it_b,
        "achieved": it_b / (step_ms * 1e-3) / 1e9, "peak": hbm_peak, "frac": it_b / (step_ms * 1e-3) / 1e9 / hbm_peak}
    out["lin_roofline"] = {"kernel": "k_lin", "bound": "hbm", "achieved": ba_lin_bytes(sc) / (kern_ms["k_lin"] * 1e-3) / 1e9,
                           "peak": hbm_peak, "unit": "GB/s"}


def lba_latency(local_rank, n_windows=40):
    """Per-solve latency of BASolver::LBA-sized problems (ba_solver.cc:523-591: a window of <= 8 frames, 5
    iterations, 1e-4 / 1e-5) through xrb_ba_solve with host buffers — the call a mapper makes once per
    registered frame (incremental_mapper.cc:71).  Single GPU."""
    from xrsfm_b200 import ba, synth
    sc = synth.make_sequential_scene(8, 3000, 6, 4242)
    sc.pt_fixed[::3] = 1                        # SetUpLBA keeps well-triangulated points constant (:380-382)
    solver = ba.BASolver(device=local_rank)
    opts = dict(max_iterations=5, function_tolerance=1e-4, parameter_tolerance=1e-5)
    for _ in range(3):
        solver.solve_scene(sc.copy_state(), **opts)
    t0 = time.perf_counter()
    its = 0
    for _ in range(n_windows):
        its += solver.solve_scene(sc.copy_state(), **opts).num_lm_iterations
    dt = time.perf_counter() - t0
    # the same windows through the batched entry point (8 engines on this device)
    batch = [sc.copy_state() for _ in range(4 * n_windows)]
    ba.BASolver.solve_batch(batch[:16], device=local_rank, n_workers=8, **opts)
    t0 = time.perf_counter()
    ba.BASolver.solve_batch(batch, device=local_rank, n_workers=8, **opts)
    dtb = time.perf_counter() - t0
    return {"ms_per_solve": dt / n_windows * 1e3, "solves_per_s": n_windows / dt, "lm_iterations_per_solve": its / n_windows,
            "window": {"frames": int(sc.n_cams), "points": int(sc.n_pts), "observations": int(sc.n_obs)},
            "call": "xrb_ba_solve (load + <= 5 LM iterations + fetch, host buffers)",
            "batched": {"solves_per_s": len(batch) / dtb, "windows": len(batch), "engines": 8,
                        "call": "xrb_ba_solve_batch (same windows, 8 engines on one device)"}}


def pose_refine_rate(local_rank, n_poses=4096, with_cpu=True):
    """Pose refinement after PnP (pnp.cc:38-71: ten-iteration Ceres solve per registered frame) as ONE launch over
    n_poses frames through xrb_pose_refine_batch with host buffers (copies inside the timed region), beside the
    same problems through the BA engine one by one (xrb_ba_solve) and the CPU oracle on a bounded sample."""
    from xrsfm_b200 import ba, pnp, synth
    batch = synth.make_pose_batch(n_poses, seed=99, max_pts=300)
    args = (batch["offsets"], batch["uv"], batch["xyz"], batch["intr"], batch["intr_model"])
    pnp.refine_poses(*args, batch["q"].copy(), batch["t"].copy(), inlier_mask=batch["inlier"], device=local_rank)
    best = None
    for _ in range(3):
        q, t = batch["q"].copy(), batch["t"].copy()
        t0 = time.perf_counter()
        sums = pnp.refine_poses(*args, q, t, inlier_mask=batch["inlier"], device=local_rank)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    its = int(sums["num_lm_iterations"].sum())
    out = {"poses_per_s": n_poses / best, "ms_per_batch": best * 1e3, "kernel_ms": pnp.last_kernel_ms(local_rank), "poses": n_poses,
           "correspondences": int(batch["offsets"][-1]), "lm_iterations_per_pose": its / n_poses,
           "converged": int((sums["termination_type"] == 0).sum()), "gpu_launches": 1,
           "call": "xrb_pose_refine_batch (one launch, host buffers)"}
    solver = ba.BASolver(device=local_rank)
    opts = dict(max_iterations=10, function_tolerance=1e-6, parameter_tolerance=1e-8)
    scenes = [synth.pose_as_scene(batch, p) for p in range(24)]
    for sc in scenes[:4]:
        solver.solve_scene(sc.copy_state(), **opts)
    t0 = time.perf_counter()
    for sc in scenes[4:]:
        solver.solve_scene(sc, **opts)
    out["ba_engine_per_solve"] = {"poses_per_s": 20 / (time.perf_counter() - t0), "call": "xrb_ba_solve, one pose at a time"}
    if with_cpu:
        from tests import oracle_lib as ol
        o = ol.ba_options(**opts)
        sample = [synth.pose_as_scene(batch, p) for p in range(3000)]
        t0 = time.perf_counter()
        for sc in sample:
            ol.ba_solve(sc, o, 1)
        out["cpu_baseline"] = {"value": len(sample) / (time.perf_counter() - t0), "unit": "poses/s", "cores": 1, "kind": "port",
                               "sample": "3000 of the same poses through oracle/ba_oracle.cpp (one thread each, as the "
                                         "mapper calls it)"}
    return out


def strip_private(out):
    for k in [k for k in out if k.startswith("_")]:
        del out[k]
    return out


def cpu_baseline_ba(args, sample_iters=2):
    """The CPU oracle (Ceres-faithful port; the reference's Ceres cannot be built here) on the
    host cores, same C2 scene, bounded sample of LM iterations."""
    from tests import oracle_lib as ol
    sc = make_scene_cached(args.config, args.scale)
    cores = host_cores()
    t0 = time.perf_counter()
    s = ol.ba_solve(sc, ol.ba_options(max_iterations=sample_iters, fixed_iterations=1, **BA_CONFIGS[args.config]["options"]), cores)
    dt = time.perf_counter() - t0
    return {"value": s.num_lm_iterations / dt, "unit": "LM-iterations/s", "cores": cores, "kind": "port",
            "sample": f"{s.num_lm_iterations} LM iterations of the {args.config} scene with the CPU oracle "
                      f"(oracle/ba_oracle.cpp, OpenMP {cores} threads), {dt:.1f} s"}


# ------------------------------------------------------------------------------------------
# matching arm
# ------------------------------------------------------------------------------------------
def gen_descriptors_torch(n_images, n_feat, seed, device):
    """Same distribution as synth.make_images, generated on the device (plumbing only)."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    window, n_pool = 2 * n_feat, max(2 * n_feat + 1, 100 * n_images)
    gam = torch.distributions.Gamma(torch.tensor(0.6, device=device), torch.tensor(1.0, device=device))
    torch.manual_seed(seed)
    pool = gam.sample((n_pool, 128)).float()
    pool[torch.rand((n_pool, 128), device=device, generator=g) < 0.4] = 0
    pool[:, 0] += 1e-3
    root = torch.sqrt(pool / pool.sum(1, keepdim=True))
    out = torch.empty((n_images, n_feat, 128), dtype=torch.uint8, device=device)
    for i in range(n_images):
        centre = int(i * (n_pool / n_images))
        idx = (centre - window // 2 + torch.randperm(window, device=device, generator=g)[:n_feat]) % n_pool
        d = root[idx] + 0.02 * torch.randn((n_feat, 128), device=device, generator=g)
        out[i] = torch.clamp(torch.round(512.0 * torch.clamp(d, min=0)), 0, 255).to(torch.uint8)
    return out


def run_match(args, rank, world, local_rank, n_feat=4096):
    n_images = args.match_images
    import torch
    from xrsfm_b200 import _lib, matching, synth
    torch.cuda.set_device(local_rank)
    dev = f"cuda:{local_rank}"
    lib = _lib.lib()
    block = gen_descriptors_torch(n_images, n_feat, 20260926, dev)
    pairs_all = synth.sequential_pairs(n_images, window=19, n_retrieval=5, seed=2)
    pairs = pairs_all[rank::world].copy()  # pair-sharded, no collective
    m = matching.SiftMatchGPU(n_feat)
    m.SetLanguage(matching.SiftMatchGPU.SIFTMATCH_CUDA_DEVICE0 + local_rank)  # feature_processing.cc:66-71
    assert m.VerifyContextGL() == 1, _lib.last_error()
    offs = np.arange(n_images + 1, dtype=np.int64) * n_feat
    _lib.check(lib.xrb_match_attach_device(m._h, n_images, offs.ctypes.data, block.data_ptr()), "attach")
    pd = torch.from_numpy(pairs).to(dev)
    counts = torch.zeros(pairs.shape[0], dtype=torch.int32, device=dev)
    out = torch.zeros((pairs.shape[0], n_feat, 2), dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream().cuda_stream

    def one_pass():
        _lib.check(lib.xrb_match_pairs_device(m._h, pairs.shape[0], pd.data_ptr(), 0.7, 0.8, 1, 16384,
                                              counts.data_ptr(), out.data_ptr(), n_feat, st), "pairs_device")

    for _ in range(max(1, args.warmup)):
        one_pass()
    torch.cuda.synchronize()
    reps = max(1, args.steps // 4)
    launches0 = lib.xrb_kernel_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        one_pass()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    launches = (lib.xrb_kernel_launch_count() - launches0) // reps
    # e2e: host (pinned) descriptors in, host match lists out, through xrb_match_pairs
    host_block = torch.empty((n_images * n_feat, 128), dtype=torch.uint8).pin_memory()
    host_block.copy_(block.view(-1, 128))
    hb = host_block.numpy()
    m2 = matching.SiftMatchGPU(n_feat)
    m2.SetLanguage(matching.SiftMatchGPU.SIFTMATCH_CUDA_DEVICE0 + local_rank)
    assert m2.VerifyContextGL() == 1
    t0 = time.perf_counter()
    m2.upload_packed(offs, hb)
    off, mm = m2.match_pairs(pairs)
    e2e_s = time.perf_counter() - t0
    # per-pair compat path: what the UNCHANGED feature_processing.cc:118-154 drives through the façade —
    # two host->device descriptor copies and one blocking read-back per pair
    compat = None
    if rank == 0:
        import ctypes as C
        n_cp = min(200, pairs.shape[0])
        buf = np.zeros((n_feat, 2), dtype=np.uint32)
        lib.xrb_match_set_descriptors.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int]
        def one_pair(a, b):
            lib.xrb_match_set_descriptors(m2._h, 0, n_feat, hb[a * n_feat:].ctypes.data, -1)
            lib.xrb_match_set_descriptors(m2._h, 1, n_feat, hb[b * n_feat:].ctypes.data, -1)
            return lib.xrb_match_get(m2._h, n_feat, buf.ctypes.data, 0.7, 0.8, 1)
        for a, b in pairs[:5]:
            one_pair(int(a), int(b))
        t0 = time.perf_counter()
        tot = 0
        for a, b in pairs[:n_cp]:
            tot += one_pair(int(a), int(b))
        dt = time.perf_counter() - t0
        compat = {"value": n_cp / dt, "unit": "pairs/s", "pairs": int(n_cp), "mean_matches": tot / n_cp,
                  "call": "xrb_match_set_descriptors x2 + xrb_match_get per pair (pinned host buffers, blocking) — the "
                          "like-for-like of reference_cuda_kernels below"}
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([ms, e2e_s], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_s = float(t[0]), float(t[1])
        n_tot = torch.tensor([pairs.shape[0]], device=dev)
        dist.all_reduce(n_tot)
        total_pairs = int(n_tot.item())
    else:
        total_pairs = pairs.shape[0]
    if rank != 0:
        return None
    peak, peak_src = load_peaks()
    alg = 132 * 2 * n_feat  # 128*(n1+n2) read + 4*(n1+n2) written per pair
    per_pair_s = ms * 1e-3 / pairs.shape[0]
    return {
        "metric": "SIFT match-pairs/sec", "value": total_pairs / (ms * 1e-3), "unit": "pairs/s",
        "ms_per_pass": ms, "pairs_per_pass": total_pairs, "n_gpus": world, "dtype": "u8",
        "config": {"workload": f"C3-shaped: {n_images} images x {n_feat} x 128-D uint8, "
                               f"{pairs_all.shape[0]} pairs (window 19 + 5 pseudo-retrieval), distmax 0.7, "
                               f"ratio 0.8, mutual best; descriptor set ({n_images * n_feat * 128 / 1e6:.0f} MB) > L2",
                   "variant": int(lib.xrb_match_set_variant(m._h, 0))},
        "e2e": {"value": total_pairs / e2e_s, "unit": "pairs/s",
                "h2d_bytes_per_step": int(hb.nbytes + pairs.nbytes), "d2h_bytes_per_step": int(mm.nbytes + off.nbytes),
                "call": "xrb_match_upload_packed + xrb_match_pairs (pinned host buffers)"},
        "gpu_launches": int(launches),
        "roofline": {"kernel": "score kernel (fused dot tiles + top-2 filter)", "bound": "hbm",
                     "achieved": alg / per_pair_s / 1e9, "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                     "frac": alg / per_pair_s / 1e9 / peak, "traffic": None,
                     "int8_tops": 2 * 128 * n_feat * n_feat / per_pair_s / 1e12,
                     "note": "binding roof is the integer-MAC pipe (tcgen05 kind::i8), not HBM: see DESIGN.md §M.4"},
        "mean_matches_per_pair": float(np.diff(off).mean()) if len(off) > 1 else 0.0,
        "per_pair_compat": compat,
    }


def cpu_baseline_match(n_feat=4096, seconds=12.0):
    from tests import oracle_lib as ol
    from xrsfm_b200 import synth
    imgs, _ = synth.make_images(4, n_feat, seed=3)
    cores = host_cores()
    n, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        ol.match_pair(imgs[n % 3], imgs[n % 3 + 1])
        n += 1
    dt = time.perf_counter() - t0
    out = {"value": n / dt, "unit": "pairs/s", "cores": cores, "kind": "port",
           "sample": f"{n} pairs of 4096x4096 with the CPU oracle (oracle/match_oracle.c, OpenMP {cores} threads)"}
    return out


def reference_cuda_matcher(n_feat=4096, n=40):
    """Baseline B of BASELINE.md: the reference's own CUDA kernels (ProgramCU.cu compiled
    verbatim into oracle/_ref) driven with blocking copies like SiftMatchCU.cpp, same GPU."""
    from tests import oracle_lib as ol
    from xrsfm_b200 import synth
    if ol.load_ref() is None:
        return None
    imgs, _ = synth.make_images(4, n_feat, seed=3)
    for i in range(3):
        ol.ref_match_pair(imgs[i], imgs[i + 1])
    t0 = time.perf_counter()
    for i in range(n):
        ol.ref_match_pair(imgs[i % 3], imgs[i % 3 + 1])
    dt = time.perf_counter() - t0
    return {"value": n / dt, "unit": "pairs/s", "kind": "reference CUDA kernels (SiftGPU, sm_100a recompile)",
            "sample": f"{n} pairs, host buffers, blocking H2D/D2H per pair"}


# ------------------------------------------------------------------------------------------
def reference_arm(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path on the host cores.
    The reference's BA is Ceres (un-vendored, not buildable here) -> the Ceres-faithful oracle
    port; rank 0 alone runs it.  Same workload, steps and warm-up as the native arm (one LM iteration
    of C2 is under a second on the box's cores, so nothing has to be cut)."""
    if rank != 0:
        return None
    from tests import oracle_lib as ol
    cfg = BA_CONFIGS["C2"]
    sc = make_c2(args.scale)
    cores = host_cores()
    if args.warmup > 0:
        warm = sc.copy_state()
        ol.ba_solve(warm, ol.ba_options(max_iterations=args.warmup, fixed_iterations=1, **cfg["options"]), cores)
    t0 = time.perf_counter()
    s = ol.ba_solve(sc, ol.ba_options(max_iterations=args.steps, fixed_iterations=1, **cfg["options"]), cores)
    dt = time.perf_counter() - t0
    v = s.num_lm_iterations / dt
    return {
        "impl": "reference", "metric": "BA LM-iterations/sec", "value": v, "unit": "LM-iterations/s",
        "n_gpus": world, "steps": s.num_lm_iterations, "warmup": args.warmup,
        "ms_per_step": dt / s.num_lm_iterations * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": cfg["workload"], "name": "C2",
                   "n_cams": sc.n_cams, "n_pts": sc.n_pts, "n_obs": sc.n_obs, "scale": args.scale},
        "cpu_baseline": {"value": v, "unit": "LM-iterations/s", "cores": cores, "kind": "port",
                         "sample": f"{s.num_lm_iterations} LM iterations of the C2 scene, "
                                   "CPU oracle = Ceres-faithful restatement; Ceres itself is un-vendored and "
                                   "not buildable in this image"},
        "e2e": {"value": v, "unit": "LM-iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--path", default="both", choices=["both", "ba", "match"])
    ap.add_argument("--config", default="C2", choices=sorted(BA_CONFIGS), help="BA scene of the headline line")
    ap.add_argument("--no-c4", action="store_true", help="skip the C4 object next to the C2 headline")
    ap.add_argument("--scale", type=float, default=1.0, help="shrink the BA scene (tests only; 1.0 = the named config)")
    ap.add_argument("--match-images", type=int, default=2000, help="images of the matching leg (C3 = 2000)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank, world, local_rank = dist_env()
    if args.impl == "reference":
        out = reference_arm(args, rank, world)
        if out is not None:
            print(json.dumps(out))
        return
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — xrsfm_b200 has no CPU fallback")
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group(os.environ.get("XRB_BENCH_BACKEND", "nccl"))
    out = None
    if args.path in ("both", "ba"):
        out = run_ba(args, rank, world, local_rank, args.config)
        c4 = None
        if args.config == "C2" and not args.no_c4 and args.scale == 1.0:
            c4 = run_ba(args, rank, world, local_rank, "C4")
        if rank == 0:
            ba_rooflines(out)
            if not args.no_cpu_baseline:
                out.update(parity_check(out["_scene"], out["_cfg"], 3, out["_conv_log"]))
            if world == 1:
                # secondary objects: a failure here is reported in place, it must not take the headline with it
                for key, fn in (("lba_latency", lambda: lba_latency(local_rank)),
                                ("pose_refine", lambda: pose_refine_rate(local_rank, with_cpu=not args.no_cpu_baseline))):
                    try:
                        out[key] = fn()
                    except Exception as e:  # noqa: BLE001
                        out[key] = {"error": f"{type(e).__name__}: {e}"}
            if c4 is not None:
                ba_rooflines(c4)
                keep = ("value", "unit", "ms_per_step", "steps", "n_gpus", "config", "e2e", "kernel_ms_per_solve",
                        "roofline", "gpu_launches")
                out["c4"] = {k: c4[k] for k in keep}
            strip_private(out)
    mt = None
    if args.path in ("both", "match"):
        mt = run_match(args, rank, world, local_rank)
    if rank == 0:
        if out is None:  # --path match: the matching line becomes the headline
            out = dict(mt)
            out.update({"steps": args.steps, "warmup": args.warmup, "higher_is_better": True, "scaling": "weak",
                        "vs_baseline": None, "data": "synthetic", "ms_per_step": mt["ms_per_pass"]})
            if not args.no_cpu_baseline:
                out["cpu_baseline"] = cpu_baseline_match()
                out["reference_cuda_kernels"] = reference_cuda_matcher()
        else:
            if mt is not None:
                out["matching"] = mt
            if not args.no_cpu_baseline:
                out["cpu_baseline"] = cpu_baseline_ba(args)
                if mt is not None:
                    out["matching"]["cpu_baseline"] = cpu_baseline_match()
                    out["matching"]["reference_cuda_kernels"] = reference_cuda_matcher()
        print(json.dumps(out))
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
