"""Timeline of the reduced-camera-system factorisation inside the replayed CUDA graph.

Arms xrb_debug_chol_trace, runs a few LM iterations on the C2 scene and prints, per kernel of
ba_chol.cu, how long its recording CTA ran (%globaltimer) and, for the fused diagonal CTA, where
its cycles went.  Usage (GPU box):  python tools/chol_trace.py [iterations]
"""
import ctypes as C
import sys

import numpy as np

sys.path.insert(0, ".")
import bench  # noqa: E402
from xrsfm_b200 import _lib, ba  # noqa: E402

NAMES = {0: "chol_diag", 1: "chol_panel", 2: "chol_update<128>", 3: "chol_update<64>",
         4: "chol_update<64>+diag", 5: "chol_backsolve"}


def main():
    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    sc = bench.make_c2(1.0)
    s = ba.BASolver(device=0)
    s.load(sc)
    s.run(max_iterations=3, fixed_iterations=1, **bench.GBA_ACCURATE)  # warm-up, graph capture
    s.reset()
    lib = _lib.lib()
    _lib.check(lib.xrb_debug_chol_trace(1, None, 0), "trace on")
    s.run(max_iterations=iters, fixed_iterations=1, **bench.GBA_ACCURATE)
    buf = np.zeros((4096, 12), dtype=np.int64)
    n = lib.xrb_debug_chol_trace(0, buf.ctypes.data, 4096)
    assert n >= 0, _lib.last_error()
    r = buf[:n]
    print(f"{n} records over {iters} LM iterations")
    for kid in sorted(set(r[:, 0].tolist())):
        m = r[r[:, 0] == kid]
        d = (m[:, 3] - m[:, 2]) * 1e-3
        print(f"{NAMES.get(kid, kid):24s} n={len(m):4d}  mean {d.mean():7.2f} us  median {np.median(d):7.2f}  "
              f"max {d.max():7.2f}  sum/iter {d.sum() / iters:8.1f} us")
        if kid in (0, 4):
            ph = m[:, 4:].mean(axis=0)
            lab = ["gemm", "warp 16x16 factor+inv", "inner panel", "inner trailing", "inverse off-diag",
                   "C update + stage", "store", "-"]
            print("    phase cycles (mean): " + ", ".join(f"{a} {b:.0f}" for a, b in zip(lab, ph) if b))
    # critical chain of one solve: the last `per` records of panels and fused tiles
    chain = r[np.isin(r[:, 0], (1, 4))]
    chain = chain[np.argsort(chain[:, 2])]
    if len(chain) > 8:
        t = chain[-60:]
        gaps = (t[1:, 2] - t[:-1, 3]) * 1e-3
        kinds = [f"{NAMES[a][5:11]}->{NAMES[b][5:11]}" for a, b in zip(t[:-1, 0], t[1:, 0])]
        for k in sorted(set(kinds)):
            g = np.array([x for x, y in zip(gaps, kinds) if y == k])
            print(f"gap {k:18s} n={len(g):3d} mean {g.mean():6.2f} us  median {np.median(g):6.2f}")
        span = (t[-1, 3] - t[0, 2]) * 1e-3
        print(f"last {len(t)} chain kernels span {span:.1f} us -> {span / (len(t) / 2):.1f} us per block column")


if __name__ == "__main__":
    main()
