"""The C-ABI library loads and exports every symbol include/xrsfm_b200.h declares; the
product refuses to run without a GPU instead of falling back to a CPU path."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "xrsfm_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(xrb_[a-z0-9_]+)\s*\(", src)) - {"xrb_allreduce_fn"})


def test_library_exports_every_declared_symbol():
    from xrsfm_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), "run __graft_entry__.build() first"
    h = C.CDLL(_lib.LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(h, name), f"{name} declared in include/xrsfm_b200.h but not exported"
    assert set(declared) == set(_lib.SIGNATURES), "ctypes table drifted from the header"
    assert _lib.lib().xrb_abi_version() == 2


def test_no_cpu_fallback_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from xrsfm_b200 import _lib, matching
    assert not _lib.lib().xrb_match_create(4096, 0)
    assert "no CUDA device" in _lib.last_error() or "CUDA" in _lib.last_error()
    m = matching.SiftMatchGPU()
    assert m.VerifyContextGL() == 0
    with pytest.raises(_lib.XrbError):
        m.SetDescriptors(0, 1, [[0] * 128])
    assert not _lib.lib().xrb_ba_create(0)
    # the batched entry points take no handle: they must refuse just as loudly
    from xrsfm_b200 import pnp, synth
    b = synth.make_pose_batch(2, seed=1)
    q0 = b["q"].copy()
    with pytest.raises(_lib.XrbError, match="CUDA"):
        pnp.refine_poses(b["offsets"], b["uv"], b["xyz"], b["intr"], b["intr_model"], b["q"], b["t"])
    assert (b["q"] == q0).all()


def test_product_never_imports_the_oracle():
    """Only tests/, bench.py and __graft_entry__.smoke() may touch oracle/."""
    pkg = os.path.join(ROOT, "xrsfm_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", "Makefile")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert "liboracle" not in txt and "oracle_lib" not in txt and "oracle/" not in txt, f
