#!/usr/bin/env python
"""Stand-alone check + timing of the reduced-camera-system solver (xrb_debug_tile_solve) against
numpy on random SPD systems: dense and banded, ragged sizes.  GPU box:  python tools/chol_check.py"""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from xrsfm_b200 import _lib  # noqa: E402


def spd(n, bw, rng):
    """Dense: B B^T + eps I.  Banded: random symmetric band made diagonally dominant."""
    if bw >= n - 1:
        B = rng.standard_normal((n, n))
        return B @ B.T + n * 1e-2 * np.eye(n)
    A = np.zeros((n, n))
    rowsum = np.zeros(n)
    for d in range(1, bw + 1):
        v = rng.standard_normal(n - d)
        idx = np.arange(n - d)
        A[idx + d, idx] = v
        A[idx, idx + d] = v
        rowsum[idx] += np.abs(v)
        rowsum[idx + d] += np.abs(v)
    A[np.arange(n), np.arange(n)] = rowsum + 1.0
    return A


def ref_solve(A, b, bw):
    n = A.shape[0]
    if bw >= n - 1:
        return np.linalg.solve(A, b)
    from scipy.linalg import solveh_banded
    ab = np.zeros((bw + 1, n))
    for d in range(bw + 1):
        ab[d, : n - d] = A[np.arange(d, n), np.arange(n - d)]
    return solveh_banded(ab, b, lower=True)


def main():
    import torch
    torch.cuda.init()
    lib = _lib.lib()
    rng = np.random.default_rng(5)
    cases = [(64, 63), (100, 99), (192, 191), (500, 499), (1000, 999), (2994, 2993), (700, 70), (3000, 59), (8000, 59),
             (5000, 300)]
    if len(sys.argv) > 1:
        cases = [tuple(int(v) for v in a.split(",")) for a in sys.argv[1:]]
    for n, bw in cases:
        A = spd(n, bw, rng)
        b = rng.standard_normal(n)
        x = np.zeros(n)
        ms = C.c_double(0)
        trace = os.environ.get("XRB_TRACE")
        if trace:
            lib.xrb_debug_chol_trace(1, None, 0)
        t0 = time.perf_counter()
        rc = lib.xrb_debug_tile_solve(n, bw, A.ctypes.data, b.ctypes.data, x.ctypes.data, 3, C.byref(ms))
        wall = time.perf_counter() - t0
        if rc != 0:
            print(f"n={n} bw={bw}: FAILED rc={rc}: {_lib.last_error()}", flush=True)
            continue
        if trace:
            buf = np.zeros(4096 + 16, dtype=np.int64)
            lib.xrb_debug_chol_trace(0, buf.ctypes.data, 4096 + 16)
            st = buf[4096:].astype(np.float64)
            if st[6] > 0:
                print(f"   workers ({st[6]:.0f}): busy span {st[5] / st[6]:.0f} cycles each, waiting on flags {st[0] / st[6]:.0f}; "
                      f"P tasks {st[3]:.0f} x {st[1] / max(1, st[3]):.0f} cycles, U tasks {st[4]:.0f} x {st[2] / max(1, st[4]):.0f} cycles (waits included)")
            nt = (n + 63) // 64
            nt = min(nt, 256)
            t = buf[: nt * 16].reshape(nt, 16).astype(np.float64)
            m = t[2: max(3, nt - 2)]
            seg = {"wait+load": (0, 1), "trsm": (1, 7), "store L+publish": (7, 2), "rhs": (2, 9), "syrk gemm": (9, 10),
                   "D-=acc(+pad)": (10, 3), "potrf": (3, 4), "y (warp 0)": (4, 8), "y+store": (4, 5), "publish": (5, 6)}
            print("   chain, mean cycles per block column:", ", ".join(f"{nm} {(m[:, b] - m[:, a]).mean():.0f}" for nm, (a, b) in seg.items()),
                  f"| step {np.diff(t[:, 0]).mean():.0f}")
            if os.environ.get("XRB_TRACE") == "2":
                print("   wait+load per step:", " ".join(f"{v / 1e3:.0f}k" for v in (t[:, 1] - t[:, 0])))
                print("   step length       :", " ".join(f"{v / 1e3:.0f}k" for v in np.diff(t[:, 0])))
        ref = ref_solve(A, b, bw)
        err = np.abs(x - ref).max() / max(1e-300, np.abs(ref).max())
        flops = (n ** 3 / 3.0 if bw >= n - 1 else n * bw * bw) + 4.0 * n * min(n, bw)
        print(f"n={n} bw={bw}: rel err {err:.2e}  {ms.value:.3f} ms  {flops / ms.value / 1e9:.2f} TFLOP/s  (wall {wall:.1f} s)",
              flush=True)


if __name__ == "__main__":
    main()
