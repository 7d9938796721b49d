"""Regenerates tests/golden/ba_trace_*.json from the CPU oracle (oracle/ba_oracle.cpp).

The reference pins no BA results (no tests; Ceres un-vendored), so these traces freeze OUR
Ceres-faithful oracle: a deliberate change to the oracle must be accompanied by rerunning
this script and explaining the diff.  Usage:  python tests/golden/make_ba_golden.py
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from tests import oracle_lib as ol  # noqa: E402
from xrsfm_b200 import synth  # noqa: E402

for name, optname in (("C1_gba_accurate", "GBA_ACCURATE"), ("C1_kgba", "KGBA")):
    sc = synth.make_scene("C1")
    s = ol.ba_solve(sc, ol.ba_options(**getattr(ol, optname)))
    d = ol.summary_dict(s)
    d.pop("total_time_in_seconds")
    d["options"] = optname
    d["cam_t_head"] = sc.cam_t.ravel()[:30].tolist()
    d["pts_head"] = sc.pts.ravel()[:30].tolist()
    with open(os.path.join(HERE, f"ba_trace_{name}.json"), "w") as f:
        json.dump(d, f, indent=1)
    print(name, d["n_iterations_logged"], d["final_cost"])
