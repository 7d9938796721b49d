"""CPU tests of the matcher oracle (oracle/match_oracle.c) against an independent numpy
restatement (int32 matmul + vectorised top-2) and known-answer cases.  No GPU needed.

Reference semantics under test: ProgramCU.cu:1491-1578 (dot + column partials),
:1780-1837 (row top-2 + thresholds), :1852-1872 (column merge), SiftMatchCU.cpp:186-215
(mutual filter, ascending order, truncation)."""
import numpy as np
import pytest

from tests import oracle_lib as ol
from xrsfm_b200 import synth


def numpy_match(d1, d2, distmax=0.7, ratiomax=0.8, mbm=True, max_match=16384):
    """Second opinion; valid when no best value is tied (ties are rejected by the ratio test
    for ratiomax <= 1, so the argmax choice cannot matter)."""
    dot = d1.astype(np.int32) @ d2.astype(np.int32).T

    def side(m):
        srt = np.sort(m, axis=1)
        best = np.maximum(srt[:, -1], 0)
        second = np.maximum(srt[:, -2], 0) if m.shape[1] > 1 else np.zeros_like(best)
        arg = np.argmax(m, axis=1)
        x = np.minimum((best.astype(np.float32) * np.float32(2.0 ** -18)).astype(np.float64), 1.0)
        xn = np.minimum((second.astype(np.float32) * np.float32(2.0 ** -18)).astype(np.float64), 1.0)
        dist = np.arccos(x).astype(np.float32)
        distn = np.arccos(xn).astype(np.float32)
        ok = (dist < np.float32(distmax)) & (dist < distn * np.float32(ratiomax)) & (best > 0)
        return np.where(ok, arg, -1)

    m12 = side(dot)
    m21 = side(dot.T)
    out = []
    for i in range(d1.shape[0]):
        j = m12[i]
        if j >= 0 and (not mbm or m21[j] == i):
            out.append((i, j))
            if len(out) >= max_match:
                break
    return np.array(out, dtype=np.uint32).reshape(-1, 2), m12, m21


@pytest.mark.parametrize("n1,n2,seed", [(256, 256, 1), (300, 517, 2), (1, 40, 3), (40, 1, 4), (513, 64, 5)])
def test_oracle_equals_numpy_on_planted_pairs(n1, n2, seed):
    imgs, _ = synth.make_images(2, max(n1, n2), seed=seed)
    d1, d2 = imgs[0][:n1], imgs[1][:n2]
    exp, e12, e21 = numpy_match(d1, d2)
    got, m12, m21 = ol.match_pair(d1, d2, want_m=True)
    np.testing.assert_array_equal(m12, e12)
    np.testing.assert_array_equal(m21, e21)
    np.testing.assert_array_equal(got, exp)


def test_dot_matrix_is_exact_int32():
    h = ol.load()
    rng = np.random.default_rng(0)
    a = rng.integers(0, 256, size=(37, 128), dtype=np.uint8)
    b = rng.integers(0, 256, size=(53, 128), dtype=np.uint8)
    a[0] = 255
    b[0] = 255  # maximum possible dot 128*255^2 = 8 323 200
    out = np.zeros((37, 53), dtype=np.int32)
    h.xro_dot_matrix(37, a.ctypes.data, 53, b.ctypes.data, out.ctypes.data)
    np.testing.assert_array_equal(out, a.astype(np.int64) @ b.astype(np.int64).T)
    assert out[0, 0] == 8323200


def test_threshold_arithmetic_known_answers():
    h = ol.load()
    assert h.xro_dist_of_dot(262144) == 0.0            # cos = 1
    assert h.xro_dist_of_dot(8323200) == 0.0           # clamped by min(.,1.0)
    assert h.xro_dist_of_dot(0) == pytest.approx(np.float32(np.pi / 2))
    assert h.xro_dist_of_dot(131072) == pytest.approx(np.float32(np.arccos(0.5)))
    # distance test is strict: dist < distmax
    d = float(h.xro_dist_of_dot(200000))
    assert h.xro_accept(200000, 0, np.float32(d), 0.8) == 0
    assert h.xro_accept(200000, 0, float(np.nextafter(np.float32(d), np.float32(10))), 0.8) == 1
    # ratio test: ties are always rejected for ratiomax <= 1
    assert h.xro_accept(250000, 250000, 0.7, 0.8) == 0
    assert h.xro_accept(250000, 250000, 0.7, 1.0) == 0
    assert h.xro_accept(262144, 262144, 0.7, 0.8) == 0  # dist = distn = 0: 0 < 0 false


def test_ties_for_best_are_rejected_and_second_counts_equal_values():
    rng = np.random.default_rng(9)
    base = synth.random_descriptors(8, rng)
    d1 = base[:1].copy()
    d2 = np.concatenate([base[:1], base[:1], base[1:]])  # two identical best candidates
    got, m12, m21 = ol.match_pair(d1, d2, want_m=True)
    assert m12[0] == -1 and got.shape[0] == 0
    # a unique best passes
    got2 = ol.match_pair(d1, np.concatenate([base[:1], base[1:]]))
    np.testing.assert_array_equal(got2, [[0, 0]])


def test_tie_break_order_matches_reference_scans_when_ratio_above_one():
    """With ratiomax > 1 a tied best can be accepted; the reported index must follow the
    reference's scan order: each lane (j % 32) keeps its lowest j, then the 16/8/4/2/1 tree
    (ProgramCU.cu:1798-1826) prefers the lower tree position, i.e. lanes in bit-reversed
    order 0,16,8,24,4,20,...; columns prefer the lowest i (:1556-1570,1858-1864)."""
    rng = np.random.default_rng(10)
    pool = synth.random_descriptors(80, rng)
    # query = 0.8 x pool[0]: its dot with pool[0] stays below 2^18, so dist > 0 and a tie
    # (dist == distn) passes dist < distn * 1.5
    d1 = (pool[:1].astype(np.float32) * 0.8).astype(np.uint8)
    d2 = pool[10:80].copy()
    d2[2] = pool[0]    # lane 2  (bit-reversed rank 8)
    d2[52] = pool[0]   # lane 20 (bit-reversed rank 5)  <- beats lane 2 although 52 > 2
    d2[64] = pool[0]   # lane 0  (rank 0)               <- beats both
    _, m12, _ = ol.match_pair(d1, d2, distmax=1.0, ratiomax=1.5, mbm=0, want_m=True)
    assert m12[0] == 64
    d2[64] = pool[20]
    _, m12, _ = ol.match_pair(d1, d2, distmax=1.0, ratiomax=1.5, mbm=0, want_m=True)
    assert m12[0] == 52
    d2[20] = pool[0]   # same lane 20, lower j wins inside the lane
    _, m12, _ = ol.match_pair(d1, d2, distmax=1.0, ratiomax=1.5, mbm=0, want_m=True)
    assert m12[0] == 20
    # columns: lowest row index among ties
    e1 = pool[10:40].copy()
    e1[17] = pool[0]
    e1[9] = pool[0]
    _, _, m21 = ol.match_pair(e1, d1, distmax=1.0, ratiomax=1.5, mbm=1, want_m=True)
    assert m21[0] == 9


def test_empty_sets_and_truncation():
    imgs, _ = synth.make_images(2, 128, seed=3)
    assert ol.match_pair(imgs[0][:0], imgs[1]).shape == (0, 2)
    assert ol.match_pair(imgs[0], imgs[1][:0]).shape == (0, 2)
    full = ol.match_pair(imgs[0], imgs[1])
    assert full.shape[0] > 10
    cut = ol.match_pair(imgs[0], imgs[1], max_match=7)
    np.testing.assert_array_equal(cut, full[:7])         # SiftMatchCU.cpp:199 stops at max_match
    assert np.all(np.diff(full[:, 0].astype(np.int64)) > 0)  # ascending first index


def test_one_way_matching_ignores_columns():
    imgs, _ = synth.make_images(2, 200, seed=4)
    one = ol.match_pair(imgs[0], imgs[1], mbm=0)
    two = ol.match_pair(imgs[0], imgs[1], mbm=1)
    assert set(map(tuple, two.tolist())) <= set(map(tuple, one.tolist()))


def test_all_zero_descriptors_never_match():
    z = np.zeros((5, 128), dtype=np.uint8)
    imgs, _ = synth.make_images(1, 16, seed=5)
    _, m12, m21 = ol.match_pair(z, imgs[0], distmax=3.0, ratiomax=2.0, want_m=True)
    assert (m12 == -1).all() and (m21 == -1).all()  # best stays at its initial (0, idx -1)


def test_quantisation_rule():
    raw = np.array([[4.0] + [0.0] * 127, [1.0] * 128], dtype=np.float32)
    q = synth.quantize_descriptors(raw)
    assert q[0, 0] == 255 and q[0, 1:].sum() == 0           # 512*sqrt(1) clamps to 255
    assert (q[1] == round(512 * np.sqrt(1 / 128))).all()     # 45


def test_oracle_equals_reference_kernel_golden():
    """The committed outputs of the reference's own CUDA kernels (tests/golden/match_ref_golden.npz,
    generated on a B200 by tests/golden/make_match_golden.py from ProgramCU.cu compiled verbatim)
    pin the CPU oracle without a GPU: match lists and both index maps, 8 cases."""
    import hashlib
    import importlib.util
    import os
    here = os.path.dirname(os.path.abspath(__file__))
    spec = importlib.util.spec_from_file_location("make_match_golden", os.path.join(here, "golden", "make_match_golden.py"))
    mk = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mk)
    gold = np.load(os.path.join(here, "golden", "match_ref_golden.npz"))
    for k, (seed, n1, n2, dmax, rmax, mbm, mm) in enumerate(mk.CASES):
        a, b = mk.inputs(seed, n1, n2)
        sha = np.frombuffer(hashlib.sha1(a.tobytes() + b.tobytes()).digest(), dtype=np.uint8)
        np.testing.assert_array_equal(sha, gold[f"case{k}_sha1"], err_msg=f"case {k}: the synthetic inputs changed")
        m, m12, m21 = ol.match_pair(a, b, dmax, rmax, mbm, mm, want_m=True)
        np.testing.assert_array_equal(m, gold[f"case{k}_matches"], err_msg=f"case {k}")
        np.testing.assert_array_equal(m12, gold[f"case{k}_m12"], err_msg=f"case {k} m12")
        if mbm:  # the reference only runs the column pass for the mutual test (SiftMatchCU.cpp:190-197)
            np.testing.assert_array_equal(m21, gold[f"case{k}_m21"], err_msg=f"case {k} m21")
