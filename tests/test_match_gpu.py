"""GPU parity tests of path M: the sm_100a matcher (through the C ABI) vs the CPU oracle, vs
the reference's own CUDA kernels (oracle/_ref, compiled verbatim from ProgramCU.cu) and vs
size-independent properties at BASELINE.json's full per-pair size (4096 x 4096).

Bar: bit-exact match lists (same length, same (i, j) sequence)."""
import ctypes as C

import numpy as np
import pytest

from tests import oracle_lib as ol
from xrsfm_b200 import _lib, matching, synth

pytestmark = pytest.mark.gpu

VARIANTS = [1, 2, 3]


@pytest.fixture(scope="module")
def matcher():
    m = matching.SiftMatchGPU()
    assert matching.CreateSiftGPUMatcher(m), _lib.last_error()
    return m


def _variants(m):
    out = []
    for v in VARIANTS:
        if m.set_variant(v) == v:
            out.append(v)
    m.set_variant(0)
    return out


def test_cuda_acos_table_equals_oracle_libm_over_whole_domain():
    """float(acos(double(min(dot*2^-18, 1)))) for EVERY dot that can reach the thresholds:
    CUDA's double acos (what ProgramCU.cu:1830 runs) vs the oracle's libm."""
    n = 262144 + 64
    dev = np.zeros(n, dtype=np.float32)
    _lib.check(_lib.lib().xrb_match_debug_dist_table(dev.ctypes.data, n), "dist_table")
    h = ol.load()
    cpu = np.array([h.xro_dist_of_dot(i) for i in range(0, n, 1)], dtype=np.float32)
    diff = np.flatnonzero(dev != cpu)
    assert diff.size == 0, f"{diff.size} entries differ, first {diff[:5]}"


@pytest.mark.parametrize("n1,n2,seed", [(512, 512, 1), (300, 517, 2), (1, 40, 3), (40, 1, 4),
                                        (129, 127, 5), (1000, 64, 6), (4096, 4096, 7)])
def test_pair_equals_oracle(matcher, n1, n2, seed):
    imgs, _ = synth.make_images(2, max(n1, n2), seed=seed)
    d1, d2 = imgs[0][:n1], imgs[1][:n2]
    exp = ol.match_pair(d1, d2)
    for v in _variants(matcher):
        matcher.set_variant(v)
        got = matching.SiftMatch(d1, d2, matcher)
        assert got.shape == exp.shape, (v, got.shape, exp.shape)
        np.testing.assert_array_equal(got, exp)
    matcher.set_variant(0)


def test_reference_kernels_agree(matcher):
    """Oracle == the reference's own MultiplyDescriptor/RowMatch/ColMatch kernels == ours."""
    if ol.load_ref() is None:
        pytest.skip("oracle/_ref/libxrref_match.so not built")
    for n1, n2, seed in [(512, 512, 11), (777, 1025, 12), (4096, 4096, 13), (33, 2000, 14)]:
        imgs, _ = synth.make_images(2, max(n1, n2), seed=seed)
        d1, d2 = imgs[0][:n1], imgs[1][:n2]
        ref, r12, r21 = ol.ref_match_pair(d1, d2, want_m=True)
        exp, e12, e21 = ol.match_pair(d1, d2, want_m=True)
        np.testing.assert_array_equal(r12, e12)
        np.testing.assert_array_equal(r21, e21)
        np.testing.assert_array_equal(ref, exp)
        got = matching.SiftMatch(d1, d2, matcher)
        np.testing.assert_array_equal(got, ref)
    # loose thresholds + one-way + tie-break order (ratiomax > 1 accepts ties)
    rng = np.random.default_rng(3)
    d1 = synth.random_descriptors(300, rng)
    d2 = np.concatenate([d1[:50], d1[:50], synth.random_descriptors(200, rng)])
    q = (d1.astype(np.float32) * 0.8).astype(np.uint8)
    for distmax, ratiomax, mbm in [(1.0, 1.5, 0), (1.0, 1.5, 1), (0.9, 0.95, 1), (2.0, 1.0, 0)]:
        ref, r12, r21 = ol.ref_match_pair(q, d2, distmax, ratiomax, mbm, want_m=True)
        exp, e12, e21 = ol.match_pair(q, d2, distmax, ratiomax, mbm, want_m=True)
        np.testing.assert_array_equal(r12, e12)
        if mbm:
            np.testing.assert_array_equal(r21, e21)
        np.testing.assert_array_equal(ref, exp)
        for v in _variants(matcher):
            matcher.set_variant(v)
            matcher.SetDescriptors(0, q.shape[0], q)
            matcher.SetDescriptors(1, d2.shape[0], d2)
            n, got = matcher.GetSiftMatch(16384, distmax, ratiomax, mbm)
            assert n == ref.shape[0], (v, distmax, ratiomax, mbm)
            np.testing.assert_array_equal(got, ref)
    matcher.set_variant(0)


def test_edge_cases(matcher):
    imgs, _ = synth.make_images(2, 256, seed=21)
    for v in _variants(matcher):
        matcher.set_variant(v)
        # empty sets -> 0 (SiftMatchCU.cpp:179-180)
        matcher.SetDescriptors(0, 0, imgs[0][:0])
        matcher.SetDescriptors(1, 256, imgs[1])
        assert matcher.GetSiftMatch(16384)[0] == 0
        # truncation at max_match keeps the first matches in ascending order
        full = ol.match_pair(imgs[0], imgs[1])
        matcher.SetDescriptors(0, 256, imgs[0])
        n, got = matcher.GetSiftMatch(9)
        assert n == 9
        np.testing.assert_array_equal(got, full[:9])
        # ties for best are rejected
        d2 = np.concatenate([imgs[0][:5], imgs[0][:5], imgs[1][:100]])
        exp = ol.match_pair(imgs[0][:5], d2)
        matcher.SetDescriptors(0, 5, imgs[0][:5])
        matcher.SetDescriptors(1, d2.shape[0], d2)
        n, got = matcher.GetSiftMatch(16384)
        np.testing.assert_array_equal(got, exp)
        # all-zero descriptors, unfiltered thresholds
        z = np.zeros((40, 128), dtype=np.uint8)
        matcher.SetDescriptors(0, 40, z)
        matcher.SetDescriptors(1, 256, imgs[1])
        assert matcher.GetSiftMatch(16384, 3.0, 2.0, 1)[0] == 0
        # saturated descriptors: the largest representable dot
        s = np.full((3, 128), 255, dtype=np.uint8)
        exp = ol.match_pair(s, np.concatenate([s[:1], imgs[1][:64]]), 0.7, 0.8)
        matcher.SetDescriptors(0, 3, s)
        matcher.SetDescriptors(1, 65, np.concatenate([s[:1], imgs[1][:64]]))
        n, got = matcher.GetSiftMatch(16384)
        np.testing.assert_array_equal(got, exp)
    matcher.set_variant(0)


def test_num_clamped_to_max_features_and_id_skips_upload():
    m = matching.SiftMatchGPU(64)
    assert m.VerifyContextGL() == 1
    assert m.GetMaxSift() == 64
    imgs, _ = synth.make_images(2, 100, seed=22)
    exp = ol.match_pair(imgs[0][:64], imgs[1][:64])
    m.SetDescriptors(0, 100, imgs[0])
    m.SetDescriptors(1, 100, imgs[1])
    n, got = m.GetSiftMatch(16384)
    np.testing.assert_array_equal(got, exp)
    # same id -> descriptors are NOT re-uploaded (SiftMatchCU.cpp:110-111)
    m.SetDescriptors(0, 64, imgs[0], id=5)
    m.SetDescriptors(0, 64, imgs[1], id=5)
    n2, got2 = m.GetSiftMatch(16384)
    np.testing.assert_array_equal(got2, exp)


def test_batched_pairs_equal_per_pair_oracle(matcher):
    n_img = 12
    rng = np.random.default_rng(5)
    imgs, _ = synth.make_images(n_img, 700, seed=31)
    imgs = [im[: int(rng.integers(300, 701))] for im in imgs]  # ragged
    imgs[3] = imgs[3][:0]                                       # an empty image
    pairs = synth.sequential_pairs(n_img, window=4, n_retrieval=2, seed=1)
    for v in _variants(matcher):
        matcher.set_variant(v)
        matcher.upload_images(imgs)
        off, mm = matcher.match_pairs(pairs)
        assert off[0] == 0 and off.shape[0] == pairs.shape[0] + 1
        for p, (a, b) in enumerate(pairs):
            exp = ol.match_pair(imgs[a], imgs[b])
            np.testing.assert_array_equal(mm[off[p]: off[p + 1]], exp)
    matcher.set_variant(0)
    # capacity error reports sizes
    with pytest.raises(_lib.XrbError):
        matcher.match_pairs(pairs, capacity=3)


def test_pipelined_chunks_keep_pair_order_and_offsets(matcher):
    """xrb_match_pairs drains chunk c on a copy stream while chunk c + 1 is scored (chunks of 512
    pairs): > 3 chunks of ragged small images, every pair against the oracle, offsets monotone."""
    n_img = 40
    rng = np.random.default_rng(11)
    imgs, _ = synth.make_images(n_img, 160, seed=77)
    imgs = [im[: int(rng.integers(60, 161))] for im in imgs]
    a, b = np.meshgrid(np.arange(n_img), np.arange(n_img), indexing="ij")
    pairs = np.stack([a.ravel(), b.ravel()], axis=1).astype(np.int32)  # 1600 pairs incl. (i, i)
    matcher.set_variant(0)
    matcher.upload_images(imgs)
    off, mm = matcher.match_pairs(pairs)
    assert off[0] == 0 and np.all(np.diff(off) >= 0) and off[-1] == mm.shape[0]
    for p in rng.choice(pairs.shape[0], 200, replace=False).tolist() + [0, 511, 512, 1023, 1024, 1599]:
        i, j = pairs[p]
        np.testing.assert_array_equal(mm[off[p]: off[p + 1]], ol.match_pair(imgs[i], imgs[j]))
    # a second call reuses the double buffers
    off2, mm2 = matcher.match_pairs(pairs[:700])
    np.testing.assert_array_equal(off2, off[:701])
    np.testing.assert_array_equal(mm2, mm[: off[700]])


def test_ftr_file_to_hbm_matches_like_uploaded_images(tmp_path, matcher):
    """ftr.bin -> HBM (xrb_match_upload_ftr, pinned staging ring) -> batched matching -> fp.bin:
    same match lists as the in-memory upload, and the file the reference's reader expects."""
    from tests import io_ref
    from xrsfm_b200 import io_formats
    rng = np.random.default_rng(21)
    imgs, _ = synth.make_images(5, 300, seed=13)
    imgs = [im[: int(rng.integers(120, 301))] for im in imgs]
    imgs[2] = imgs[2][:0]
    frames = [dict(name=f"{i}.jpg", keypoints=rng.random((len(im), 4), dtype=np.float32), descs=im) for i, im in enumerate(imgs)]
    path = str(tmp_path / "ftr.bin")
    io_ref.save_features(path, frames)
    pairs = synth.sequential_pairs(5, window=2, n_retrieval=1, seed=3)
    matcher.set_variant(0)
    matcher.upload_images(imgs)
    off_a, mm_a = matcher.match_pairs(pairs)
    matcher.upload_ftr(path)
    off_b, mm_b = matcher.match_pairs(pairs)
    np.testing.assert_array_equal(off_a, off_b)
    np.testing.assert_array_equal(mm_a, mm_b)
    out = str(tmp_path / "fp.bin")
    io_formats.SaveFramePairs(out, pairs, off_b, mm_b)
    back = io_ref.read_frame_pairs(out)
    keep = [k for k, (a, b) in enumerate(pairs) if a != b]
    assert [(p["id1"], p["id2"]) for p in back] == [tuple(pairs[k]) for k in keep]
    for p, k in zip(back, keep):
        assert [[i, j] for (i, j, _) in p["matches"]] == mm_b[off_b[k]: off_b[k + 1]].tolist()
    with pytest.raises(_lib.XrbError, match="cannot open"):
        matcher.upload_ftr(str(tmp_path / "nope.bin"))


def test_device_pointers_must_be_device_memory():
    m = matching.SiftMatchGPU(64)
    assert m.VerifyContextGL() == 1
    host = np.zeros((64, 128), dtype=np.uint8)
    offs = np.array([0, 64], dtype=np.int64)
    rc = _lib.lib().xrb_match_attach_device(m._h, 1, offs.ctypes.data, host.ctypes.data)
    assert rc == -1 and "attach_device" in _lib.last_error()  # XRB_ERR_INVALID


def test_full_size_properties(matcher):
    """4096 x 4096 (config C3 per-pair size): properties that need no oracle —
    symmetry under swapping the images, planted correspondences recovered, indices unique."""
    imgs, ids = synth.make_images(2, 4096, seed=41)
    for v in _variants(matcher):
        matcher.set_variant(v)
        ab = matching.SiftMatch(imgs[0], imgs[1], matcher)
        ba = matching.SiftMatch(imgs[1], imgs[0], matcher)
        assert set(map(tuple, ab.tolist())) == set((j, i) for i, j in ba.tolist())
        assert len(set(ab[:, 0].tolist())) == ab.shape[0] == len(set(ab[:, 1].tolist()))
        lut = {v_: i for i, v_ in enumerate(ids[0])}
        planted = {(lut[v_], j) for j, v_ in enumerate(ids[1]) if v_ in lut}
        found = set(map(tuple, ab.tolist()))
        assert len(found & planted) >= 0.98 * len(planted)
        assert len(found - planted) <= 0.01 * len(planted) + 2
    matcher.set_variant(0)


def test_images_above_the_fused_limit_fall_back_to_the_global_state_path():
    """> 4096 descriptors per image: generation 3 hands over to generation 2 (same MMA pipeline,
    top-2 state in global memory); the reference allows up to 8192 features per image
    (feature_extraction.cc:24) and Allocate(16384)."""
    m = matching.SiftMatchGPU()
    assert matching.CreateSiftGPUMatcher(m)  # max_sift 16384 like the reference
    imgs, _ = synth.make_images(2, 5000, seed=51)
    d1, d2 = imgs[0], imgs[1][:4500]
    exp = ol.match_pair(d1, d2)
    got = matching.SiftMatch(d1, d2, m)
    np.testing.assert_array_equal(got, exp)
    m.upload_images([d1, d2, imgs[1][:100]])
    off, mm = m.match_pairs(np.array([[0, 1], [2, 0]], dtype=np.int32))
    np.testing.assert_array_equal(mm[off[0]: off[1]], exp)
    np.testing.assert_array_equal(mm[off[1]: off[2]], ol.match_pair(imgs[1][:100], d1))


def test_out_of_range_pair_index(matcher):
    """Host path: a pair outside the resident set is rejected loudly.  Device path (indices already in HBM, the
    host cannot look at them): the pair is empty — no out-of-bounds read."""
    import torch
    from xrsfm_b200 import _lib
    imgs, _ = synth.make_images(3, 300, seed=5)
    matcher.upload_images(imgs)
    with pytest.raises(_lib.XrbError, match="outside"):
        matcher.match_pairs(np.array([[0, 1], [0, 7]], dtype=np.int32))
    pairs = torch.tensor([[0, 1], [0, 7], [-1, 2], [1, 2]], dtype=torch.int32, device="cuda")
    counts = torch.zeros(4, dtype=torch.int32, device="cuda")
    out = torch.zeros((4, 300, 2), dtype=torch.int32, device="cuda")
    _lib.check(_lib.lib().xrb_match_pairs_device(matcher._h, 4, pairs.data_ptr(), 0.7, 0.8, 1, 16384, counts.data_ptr(),
                                                 out.data_ptr(), 300, None), "pairs_device")
    torch.cuda.synchronize()
    n = counts.cpu().numpy()
    assert n[1] == 0 and n[2] == 0
    for p, (a, b) in ((0, (0, 1)), (3, (1, 2))):
        e = ol.match_pair(imgs[a], imgs[b])
        assert np.array_equal(out[p, : n[p]].cpu().numpy().astype(np.uint32), e)
    # the default capacity of the host path follows the resident image sizes
    off2, mm2 = matcher.match_pairs(np.array([[0, 1], [1, 2]], dtype=np.int32))
    assert off2[-1] == n[0] + n[3]
