"""One launch of the fused matcher kernel for ncu (development aid)."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from xrsfm_b200 import _lib, matching, synth  # noqa: E402

variant = int(sys.argv[1]) if len(sys.argv) > 1 else 3
n_img, n_feat = 40, 4096
imgs, _ = synth.make_images(n_img, n_feat, seed=1)
pairs = synth.sequential_pairs(n_img, window=4, n_retrieval=0, seed=0)[:148]
m = matching.SiftMatchGPU(n_feat)
assert m.VerifyContextGL() == 1
m.set_variant(variant)
lib = _lib.lib()
block = torch.from_numpy(np.concatenate(imgs)).cuda()
offs = np.arange(n_img + 1, dtype=np.int64) * n_feat
_lib.check(lib.xrb_match_attach_device(m._h, n_img, offs.ctypes.data, block.data_ptr()), "attach")
pd = torch.from_numpy(pairs).cuda()
counts = torch.zeros(pairs.shape[0], dtype=torch.int32, device="cuda")
out = torch.zeros((pairs.shape[0], n_feat, 2), dtype=torch.int32, device="cuda")
st = torch.cuda.current_stream().cuda_stream
for _ in range(2):
    _lib.check(lib.xrb_match_pairs_device(m._h, pairs.shape[0], pd.data_ptr(), 0.7, 0.8, 1, 16384,
                                          counts.data_ptr(), out.data_ptr(), n_feat, st), "pairs")
torch.cuda.synchronize()
print("pairs", pairs.shape[0], "mean matches", counts.float().mean().item())
