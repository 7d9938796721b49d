"""Golden match lists produced by the REFERENCE'S OWN CUDA kernels (oracle/_ref/libxrref_match.so:
3rdparty/SiftGPU/ProgramCU.cu compiled verbatim for sm_100a + the harness oracle/ref_harness.cu).

Run on a GPU box:   python tests/golden/make_match_golden.py gpurun_out/match_ref_golden.npz
The committed copy (tests/golden/match_ref_golden.npz) pins the CPU oracle in the CPU-only
suite (tests/test_match_oracle.py::test_oracle_equals_reference_kernel_golden): inputs are
regenerated from the seeds below (their SHA-1 is stored next to the outputs)."""
import hashlib
import sys

import numpy as np

sys.path.insert(0, ".")
from tests import oracle_lib as ol  # noqa: E402
from xrsfm_b200 import synth  # noqa: E402

# (seed, n1, n2, distmax, ratiomax, mutual_best_match, max_match)
CASES = [
    (101, 512, 512, 0.7, 0.8, 1, 16384),     # the reference's constants (feature_processing.cc:118-154)
    (102, 300, 517, 0.7, 0.8, 1, 16384),     # ragged
    (103, 700, 333, 0.7, 0.8, 0, 16384),     # no mutual filter
    (104, 640, 640, 0.9, 0.95, 1, 16384),    # loose thresholds: many accepted, ties matter
    (105, 400, 450, 0.5, 0.6, 1, 16384),     # tight thresholds
    (106, 256, 256, 0.7, 0.8, 1, 40),        # truncation at max_match (SiftMatchCU.cpp:199-207)
    (107, 1, 64, 0.7, 0.8, 1, 16384),        # a single descriptor
    (108, 1024, 1024, 0.7, 0.8, 1, 16384),
]


def inputs(seed, n1, n2):
    imgs, _ = synth.make_images(2, max(n1, n2), seed=seed)
    return imgs[0][:n1].copy(), imgs[1][:n2].copy()


def main(out_path):
    if ol.load_ref() is None:
        raise SystemExit("oracle/_ref/libxrref_match.so is missing: build it with `make -C oracle ref`")
    out = {}
    for k, (seed, n1, n2, dmax, rmax, mbm, mm) in enumerate(CASES):
        a, b = inputs(seed, n1, n2)
        m, m12, m21 = ol.ref_match_pair(a, b, dmax, rmax, mbm, mm, want_m=True)
        out[f"case{k}_matches"] = m.astype(np.uint32)
        out[f"case{k}_m12"] = m12.astype(np.int32)
        out[f"case{k}_m21"] = m21.astype(np.int32)
        out[f"case{k}_sha1"] = np.frombuffer(hashlib.sha1(a.tobytes() + b.tobytes()).digest(), dtype=np.uint8)
        print(k, (seed, n1, n2), "matches", m.shape[0])
    np.savez_compressed(out_path, **out)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/match_ref_golden.npz")
