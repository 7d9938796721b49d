#!/usr/bin/env python
"""Stand-alone check + timing of the reduced-camera-system solver (xrb_debug_tile_solve) against
numpy on random SPD systems: dense, banded (natural and dissected column order), random sparse tile
patterns, ragged sizes.  GPU box:  python tools/chol_check.py  [n,bw[,nd] ...]   XRB_TRACE=1|2 for the
chain timeline."""
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


def dissect(A, b, bw):
    """Apply the library's column order for a band of 6-column cameras (xrb_debug_column_order): returns the
    permuted, padded system and the map back."""
    n = A.shape[0]
    V = n // 6
    assert V * 6 == n
    lib = _lib.lib()
    w = np.full(V, 6, dtype=np.int32)
    start = np.zeros(V, dtype=np.int32)
    n_pad, parts = C.c_int32(0), C.c_int32(0)
    _lib.check(lib.xrb_debug_column_order(V, w.ctypes.data, bw, 1, start.ctypes.data, C.byref(n_pad), C.byref(parts)), "order")
    new = (start[:, None] + np.arange(6)[None, :]).reshape(-1)  # natural column -> new column
    Ap = np.eye(n_pad.value)
    Ap[np.ix_(new, new)] = A
    bp = np.zeros(n_pad.value)
    bp[new] = b
    return Ap, bp, new, parts.value


def main():
    import torch
    torch.cuda.init()
    lib = _lib.lib()
    rng = np.random.default_rng(5)
    cases = [(64, 63, 0), (100, 99, 0), (192, 191, 0), (500, 499, 0), (1000, 999, 0), (2994, 2993, 0), (700, 70, 0),
             (3000, 59, 0), (3000, 59, 1), (9000, 59, 0), (9000, 59, 1), (4998, 300, 1), (-1500, 0, 0)]
    if len(sys.argv) > 1:
        cases = [tuple(int(v) for v in (a.split(",") + ["0", "0"])[:3]) for a in sys.argv[1:]]
    for n, bw, nd in cases:
        label = f"n={n} bw={bw}" + (" dissected" if nd else "")
        if n < 0:  # random sparse tile pattern
            n = -n
            nt = (n + 63) // 64
            A = np.zeros((n, n))
            keep = np.tril(rng.random((nt, nt)) < 0.25)
            for i in range(nt):
                for j in range(i + 1):
                    if keep[i, j] or i == j:
                        r0, r1, c0, c1 = i * 64, min(n, i * 64 + 64), j * 64, min(n, j * 64 + 64)
                        A[r0:r1, c0:c1] = rng.standard_normal((r1 - r0, c1 - c0))
            A = np.tril(A)
            A = A + A.T
            A[np.arange(n), np.arange(n)] = np.abs(A).sum(axis=1) + 1.0
            bw = n - 1
            label = f"n={n} random sparse tiles"
        else:
            A = spd(n, bw, rng)
        b = rng.standard_normal(n)
        ref = np.linalg.solve(A, b) if n <= 6000 else None
        if ref is None:
            from scipy.linalg import solveh_banded
            ab = np.zeros((bw + 1, n))
            for d in range(bw + 1):
                ab[d, : n - d] = A[np.arange(d, n), np.arange(n - d)]
            ref = solveh_banded(ab, b, lower=True)
        As, bs, back, bw_s = A, b, None, bw
        if nd:
            As, bs, back, parts = dissect(A, b, bw)
            bw_s = As.shape[0] - 1
            label += f" ({parts} parts, {As.shape[0]} padded)"
        ns = As.shape[0]
        As = np.ascontiguousarray(As)
        x = np.zeros(ns)
        ms = C.c_double(0)
        trace = os.environ.get("XRB_TRACE")
        if trace:
            lib.xrb_debug_chol_trace(1, None, 0)
        t0 = time.perf_counter()
        rc = lib.xrb_debug_tile_solve(ns, bw_s, As.ctypes.data, bs.ctypes.data, x.ctypes.data, 3, C.byref(ms))
        wall = time.perf_counter() - t0
        if rc != 0:
            print(f"{label}: FAILED rc={rc}: {_lib.last_error()}", flush=True)
            continue
        if trace:
            buf = np.zeros(4096 + 16, dtype=np.int64)
            lib.xrb_debug_chol_trace(0, buf.ctypes.data, 4096 + 16)
            st = buf[4096:].astype(np.float64)
            if st[6] > 0:
                print(f"   workers ({st[6]:.0f}): busy span {st[5] / st[6]:.0f} cycles each, waiting on flags {st[0] / st[6]:.0f}; "
                      f"P tasks {st[3]:.0f} x {st[1] / max(1, st[3]):.0f} cycles, U tasks {st[4]:.0f} x {st[2] / max(1, st[4]):.0f} cycles (waits included)")
            nf = min((ns + 63) // 64, 256)
            t = buf[: nf * 16].reshape(nf, 16).astype(np.float64)
            m = t[2: max(3, nf - 2)]
            m = m[m[:, 7] > 0] if (m[:, 7] > 0).any() else m
            seg = {"wait+load": (0, 1), "trsm": (1, 7), "store L": (7, 2), "rhs": (2, 9), "syrk": (9, 10),
                   "D-=acc+publish": (10, 3), "potrf": (3, 4), "[fac16": (3, 11), "rows below": (11, 12), "next cols]": (12, 13), "store": (4, 5), "publish": (5, 6),
                   "task": (0, 6)}
            print("   chain, mean cycles per F task:", ", ".join(f"{nm} {(m[:, b_] - m[:, a_]).mean():.0f}" for nm, (a_, b_) in seg.items()))
            if trace == "2":
                print("   wait+load per task:", " ".join(f"{v / 1e3:.0f}k" for v in (t[:, 1] - t[:, 0])))
                print("   task length       :", " ".join(f"{v / 1e3:.0f}k" for v in (t[:, 6] - t[:, 0])))
        xs = x[back] if back is not None else x
        err = np.abs(xs - ref).max() / max(1e-300, np.abs(ref).max())
        flops = (n ** 3 / 3.0 if bw >= n - 1 else n * bw * bw) + 4.0 * n * min(n, bw)
        print(f"{label}: rel err {err:.2e}  {ms.value:.3f} ms  {flops / ms.value / 1e9:.2f} TFLOP/s  (wall {wall:.1f} s)",
              flush=True)


if __name__ == "__main__":
    main()
