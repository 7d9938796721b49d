#!/usr/bin/env python
"""GPU box: a C4-shaped scene solved with the dissected and the natural column order, and by the CPU
oracle; prints the three iteration logs side by side.   python tools/order_check.py [scale] [iters]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import oracle_lib as ol  # noqa: E402
from xrsfm_b200 import ba, synth  # noqa: E402

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.15
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20
sc = synth.make_scene("C4", scale)
opts = dict(ol.KGBA)
opts["max_iterations"] = iters
logs = {}
for order in ("nd", "natural"):
    os.environ["XRB_BA_ORDER"] = order
    s = ba.BASolver()
    w = sc.copy_state()
    t0 = time.perf_counter()
    summ = s.solve_scene(w, **opts)
    d = s.profile_detail()
    print(order, f"{time.perf_counter() - t0:.2f}s parts={d['parts']:.0f} nt={d['tile_columns']:.0f} tiles={d['tiles']:.0f}/{d['tiles_original']:.0f} "
          f"depth={d['depth_factor']:.0f} chains={d['chains']:.0f} term={summ.termination_type} its={summ.num_lm_iterations}")
    logs[order] = [(summ.iterations[i].cost, summ.iterations[i].step_is_successful) for i in range(summ.n_iterations_logged)]
    logs[order + "_state"] = w
t0 = time.perf_counter()
ref = sc.copy_state()
so = ol.ba_solve(ref, ol.ba_options(**opts), os.cpu_count() or 1)
print(f"oracle {time.perf_counter() - t0:.1f}s term={so.termination_type} its={so.num_lm_iterations}")
logs["oracle"] = [(so.iterations[i].cost, so.iterations[i].step_is_successful) for i in range(so.n_iterations_logged)]
n = max(len(logs[k]) for k in ("nd", "natural", "oracle"))
for i in range(n):
    row = []
    for k in ("nd", "natural", "oracle"):
        row.append(f"{logs[k][i][0]:.10e} {logs[k][i][1]}" if i < len(logs[k]) else "-")
    print(i, " | ".join(row))
for k in ("nd", "natural"):
    w = logs[k + "_state"]
    print(k, "max |dq|", np.abs(w.cam_q - ref.cam_q).max(), "max |dt|", np.abs(w.cam_t - ref.cam_t).max(), "max |dX|",
          np.abs(w.pts - ref.pts).max())
