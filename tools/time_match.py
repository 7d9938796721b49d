"""Quick device timing of the matcher (development aid; bench.py is the contract)."""
import ctypes as C
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from tests import oracle_lib as ol  # noqa: E402
from xrsfm_b200 import _lib, matching, synth  # noqa: E402

n_img, n_feat = 64, 4096
imgs, _ = synth.make_images(n_img, n_feat, seed=1)
pairs = synth.sequential_pairs(n_img, window=8, n_retrieval=2, seed=0)
print("pairs", pairs.shape[0])
m = matching.SiftMatchGPU()
assert matching.CreateSiftGPUMatcher(m)
lib = _lib.lib()
block = torch.from_numpy(np.concatenate(imgs)).cuda()
offs = np.arange(n_img + 1, dtype=np.int64) * n_feat
_lib.check(lib.xrb_match_attach_device(m._h, n_img, offs.ctypes.data, block.data_ptr()), "attach")
pd = torch.from_numpy(pairs).cuda()
counts = torch.zeros(pairs.shape[0], dtype=torch.int32, device="cuda")
out = torch.zeros((pairs.shape[0], n_feat, 2), dtype=torch.int32, device="cuda")
st = torch.cuda.current_stream().cuda_stream
for variant, distmax in ((1, 0.7), (2, 0.7), (2, 1e-3), (3, 0.7), (3, 1e-3)):  # distmax 1e-3: no candidates -> pure MMA + TMEM read + filter
    if m.set_variant(variant) != variant:
        continue
    for it in range(3):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        _lib.check(lib.xrb_match_pairs_device(m._h, pairs.shape[0], pd.data_ptr(), distmax, 0.8, 1, 16384,
                                              counts.data_ptr(), out.data_ptr(), n_feat, st), "pairs")
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print(f"variant {variant} distmax {distmax}: {ms:.2f} ms for {pairs.shape[0]} pairs -> {pairs.shape[0] / ms * 1e3:.0f} pairs/s,"
              f" mean matches {counts.float().mean().item():.0f}")
if ol.load_ref() is not None:
    t = time.time()
    k = 20
    for p in range(k):
        ol.ref_match_pair(imgs[pairs[p, 0]], imgs[pairs[p, 1]])
    dt = time.time() - t
    print(f"reference CUDA kernels (blocking copies): {k / dt:.1f} pairs/s")
t = time.time()
for p in range(4):
    ol.match_pair(imgs[pairs[p, 0]], imgs[pairs[p, 1]])
print(f"CPU oracle: {4 / (time.time() - t):.2f} pairs/s")
