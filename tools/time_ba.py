"""BA timing / load breakdown on the C2 scene (development aid)."""
import os
import sys
import time

sys.path.insert(0, ".")
os.environ["XRB_BA_DEBUG"] = "1"
import bench  # noqa: E402
from xrsfm_b200 import ba  # noqa: E402

sc = bench.make_c2(1.0)
s = ba.BASolver()
for rep in range(2):
    w = sc.copy_state()
    t = time.perf_counter()
    s.load(w)
    print(f"load #{rep}: {(time.perf_counter() - t) * 1e3:.1f} ms", flush=True)
t = time.perf_counter()
r = s.run(max_iterations=10, fixed_iterations=1, function_tolerance=1e-5, parameter_tolerance=1e-6)
print(f"run 10 it: {(time.perf_counter() - t) * 1e3:.1f} ms", s.profile())
