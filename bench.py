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

GBA_ACCURATE = dict(function_tolerance=1e-5, parameter_tolerance=1e-6)  # ba_solver.cc:626-629


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
def make_c2(scale=1.0):
    from xrsfm_b200 import synth
    cache = f"/tmp/xrsfm_b200_C2_{scale}.npz"
    keys = ("cam_q", "cam_t", "pts", "intr", "intr_model", "cam_intr", "obs_cam", "obs_pt", "obs_uv",
            "cam_q_fixed", "cam_t_fixed", "pt_fixed")
    if os.path.exists(cache):
        try:
            z = np.load(cache)
            sc = synth.BAScene({k: np.ascontiguousarray(z[k]) for k in keys})
            sc.n_cams, sc.n_pts, sc.n_obs, sc.n_intr = (int(z["dims"][i]) for i in range(4))
            return sc
        except Exception:
            pass
    sc = synth.make_scene("C2", scale)
    try:  # atomic publish: several ranks may build the scene at the same time
        tmp = f"{cache}.{os.getpid()}.tmp.npz"
        np.savez(tmp, dims=np.array([sc.n_cams, sc.n_pts, sc.n_obs, sc.n_intr]), **{k: sc[k] for k in keys})
        os.replace(tmp, cache)
    except Exception:
        pass
    return sc


def ba_gather_bytes(detail):
    """Algorithmic HBM bytes of ONE launch of k_gather: per (block, point)
    incidence two 144-byte observation records + the 8-byte index pair, per block the 6x6
    result written once (288 B) + 12 B of block table."""
    return int(detail["n_incidences"] * (2 * 144 + 8) + detail["n_blocks"] * (288 + 12))


def ba_lin_bytes(sc):
    """k_lin: observation stream 24 B/obs read + one 144-byte record written per observation;
    per point 24 B read and 96 B (V^-1, g, h) written."""
    return 168 * sc.n_obs + 120 * sc.n_pts


def ba_chol_flops(nc, bw):
    """Blocked Cholesky + forward/backward substitution, banded: n*bw^2 (dense: n^3/3) FMAs x 2."""
    if bw >= nc - 1:
        return 2.0 * (nc ** 3 / 6.0 + nc ** 2)
    return 2.0 * (nc * bw * bw / 2.0 + 2.0 * nc * bw)


def run_ba(args, rank, world, local_rank):
    import torch
    from xrsfm_b200 import _lib, ba
    torch.cuda.set_device(local_rank)
    sc = make_c2(args.scale)
    solver = ba.BASolver(device=local_rank)
    solver._ensure()
    if world > 1:
        import torch.distributed as dist

        def allreduce(ptr, count):
            t = torch.as_tensor(_CudaPtr(ptr, count), device=f"cuda:{local_rank}")
            dist.all_reduce(t)
            torch.cuda.current_stream().synchronize()

        solver.set_exchange(rank, world, allreduce)
    lib = _lib.lib()
    stream = torch.cuda.current_stream().cuda_stream

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    # -- resident path: load once, W warm-up iterations, then exactly K timed iterations
    solver.load(sc)
    if args.warmup > 0:
        solver.run(stream=stream, max_iterations=args.warmup, fixed_iterations=1, **GBA_ACCURATE)
    solver.reset()
    clocks = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        clocks.start()
    launches0 = lib.xrb_kernel_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    summ = solver.run(stream=stream, max_iterations=args.steps, fixed_iterations=1, **GBA_ACCURATE)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = lib.xrb_kernel_launch_count() - launches0
    clk = clocks.stop() if rank == 0 else None
    prof = solver.profile()
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    assert summ.num_lm_iterations == args.steps, (summ.num_lm_iterations, args.steps)

    # -- to-convergence run (reported, not timed as the metric)
    solver.reset()
    conv = solver.run(stream=stream, max_iterations=50, **GBA_ACCURATE)
    barrier()

    # -- e2e: host buffers through xrb_ba_solve (upload + K iterations + download)
    work = sc.copy_state()
    for k in ("cam_q", "cam_t", "pts"):
        work[k] = np.ascontiguousarray(work[k])
    h2d = sum(sc[k].nbytes for k in ("cam_q", "cam_t", "pts", "intr", "intr_model", "cam_intr", "obs_cam",
                                      "obs_pt", "obs_uv", "cam_q_fixed", "cam_t_fixed", "pt_fixed"))
    d2h = sum(sc[k].nbytes for k in ("cam_q", "cam_t", "pts"))
    barrier()
    t0 = time.perf_counter()
    s_e2e = solver.solve_scene(work, max_iterations=args.steps, fixed_iterations=1, **GBA_ACCURATE)
    barrier()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())

    if rank != 0:
        return None
    peak, peak_src = load_peaks()
    detail = solver.profile_detail() if rank == 0 else {}
    nsolve = max(1.0, detail["solves"])
    kern_ms = {"k_lin": detail["lin_ms"] / nsolve, "k_gather": detail["gather_ms"] / nsolve,
               "k_cam_blocks": detail["cam_blocks_ms"] / nsolve, "cholesky_graph": prof["solve"][0] / nsolve}
    gather_bytes = ba_gather_bytes(detail)
    ach = gather_bytes / (kern_ms["k_gather"] * 1e-3) / 1e9
    chol_tflops = ba_chol_flops(detail["nc"], detail["half_bandwidth"]) / (kern_ms["cholesky_graph"] * 1e-3) / 1e12
    out = {
        "metric": "BA LM-iterations/sec", "value": args.steps / (ms * 1e-3), "unit": "LM-iterations/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": "C2: synthetic 500-camera / 200k-point / 2M-obs global BA (GBA accurate options), "
                               "SIMPLE_RADIAL, Huber 5.99, 2 translations fixed",
                   "n_cams": sc.n_cams, "n_pts": sc.n_pts, "n_obs": sc.n_obs, "scale": args.scale,
                   "parallelism": f"points sharded over {world} GPU(s), cameras replicated, "
                                  "SUM all-reduce of the reduced camera system per solve",
                   "l2_note": "no explicit L2 flush: every iteration streams a working set larger than the 126 MB L2 "
                              "(observations 48 MB + per-observation records 288 MB + reduced system 72 MB + points)",
                   "iterations_to_convergence": conv.num_lm_iterations,
                   "termination": ba.TERMINATION.get(conv.termination_type),
                   "final_rms_px": float(np.sqrt(conv.final_cost / max(1, conv.num_residuals_reduced))),
                   "phase_ms_per_iteration": {k: v[0] / max(1, args.steps + 1) for k, v in prof.items() if k != "run"}},
        "e2e": {"value": args.steps / e2e_s, "unit": "LM-iterations/s", "h2d_bytes_per_step": h2d / args.steps,
                "d2h_bytes_per_step": d2h / args.steps, "call": "xrb_ba_solve (host buffers in/out)",
                "iterations": s_e2e.num_lm_iterations},
        "gpu_launches": int(launches),
        "clocks": clk,
        "roofline": {"kernel": "k_gather (Schur complement: per-block gather of the observation records)",
                     "bound": "hbm", "achieved": ach, "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                     "frac": ach / peak,
                     # dram__bytes_read.sum + dram__bytes_write.sum of one k_gather launch on this
                     # very scene, from the ncu --set full capture summarised in profiles/
                     "traffic": 1.355e9 if args.scale == 1.0 else None,
                     "traffic_source": "profiles/r01c_gather_and_update_summary.md",
                     "algorithmic_bytes_per_launch": gather_bytes,
                     "ms_per_launch": kern_ms["k_gather"],
                     "note": "HBM-bound kernel of the iteration; the largest share of the time is the FP64 "
                             "Cholesky (see roofline_fp64), which is bound by the FP64 FMA pipe, not HBM"},
        "roofline_fp64": {"kernel": "blocked Cholesky + substitutions (CUDA graph)", "bound": "fp64 FMA pipe",
                          "achieved": chol_tflops, "peak": 37.0, "peak_source": "nominal (B200 FP64, no measured figure "
                          "in MEASURED_PEAKS.json)", "unit": "TFLOP/s", "frac": chol_tflops / 37.0,
                          "ms_per_launch": kern_ms["cholesky_graph"]},
        "kernel_ms_per_solve": kern_ms,
        "lin_roofline": {"kernel": "k_lin", "bound": "hbm", "achieved": ba_lin_bytes(sc) / (kern_ms["k_lin"] * 1e-3) / 1e9,
                         "peak": peak, "unit": "GB/s"},
    }
    return out


def cpu_baseline_ba(args, sample_iters=2):
    """The CPU oracle (Ceres-faithful port; the reference's Ceres cannot be built here) on the
    host cores, same C2 scene, bounded sample of LM iterations."""
    from tests import oracle_lib as ol
    sc = make_c2(args.scale)
    cores = host_cores()
    t0 = time.perf_counter()
    s = ol.ba_solve(sc, ol.ba_options(max_iterations=sample_iters, fixed_iterations=1, **GBA_ACCURATE), cores)
    dt = time.perf_counter() - t0
    return {"value": s.num_lm_iterations / dt, "unit": "LM-iterations/s", "cores": cores, "kind": "port",
            "sample": f"{s.num_lm_iterations} LM iterations of the C2 scene with the CPU oracle "
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
    port; rank 0 alone runs it."""
    if rank != 0:
        return None
    from tests import oracle_lib as ol
    sc = make_c2(args.scale)
    cores = host_cores()
    steps = max(1, min(args.steps, 4))  # bounded sample: ~2-3 s per LM iteration on 8 cores
    warm = sc.copy_state()
    ol.ba_solve(warm, ol.ba_options(max_iterations=min(1, args.warmup), fixed_iterations=1, **GBA_ACCURATE), cores)
    t0 = time.perf_counter()
    s = ol.ba_solve(sc, ol.ba_options(max_iterations=steps, fixed_iterations=1, **GBA_ACCURATE), cores)
    dt = time.perf_counter() - t0
    v = s.num_lm_iterations / dt
    return {
        "impl": "reference", "metric": "BA LM-iterations/sec", "value": v, "unit": "LM-iterations/s",
        "n_gpus": world, "steps": s.num_lm_iterations, "warmup": min(1, args.warmup),
        "ms_per_step": dt / s.num_lm_iterations * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "C2: synthetic 500-camera / 200k-point / 2M-obs global BA (GBA accurate options)",
                   "n_cams": sc.n_cams, "n_pts": sc.n_pts, "n_obs": sc.n_obs, "scale": args.scale},
        "cpu_baseline": {"value": v, "unit": "LM-iterations/s", "cores": cores, "kind": "port",
                         "sample": f"{s.num_lm_iterations} LM iterations (of {args.steps} requested) of the C2 scene, "
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
    ap.add_argument("--scale", type=float, default=1.0, help="shrink C2 (tests only; 1.0 = the named config)")
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
        out = run_ba(args, rank, world, local_rank)
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
