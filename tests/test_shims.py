"""The C++ drop-in shims (xrsfm_b200/shim/*.h) compile against stand-ins of the reference
types and link against libxrsfm_b200.so (no GPU needed: the binary only constructs objects)."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shims_compile_link_and_run(tmp_path):
    exe = tmp_path / "shim_check"
    cmd = ["g++", "-std=c++17", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "mock", "shim_check.cpp"), "-o", str(exe),
           "-L", os.path.join(ROOT, "xrsfm_b200"), "-lxrsfm_b200",
           "-Wl,-rpath," + os.path.join(ROOT, "xrsfm_b200")]
    subprocess.check_call(cmd)
    out = subprocess.check_output([str(exe)], text=True)
    assert "flatten: 0 cams" in out and "VerifyContextGL=" in out
    assert "ftr roundtrip: ok" in out
    # LBA window: frames 0..3 (set order), 11 observations, gauge frame 0 fixed; point 0 free (seen by the
    # new frame, small angle), point 1 constant (angle > 5), point 2 constant (not seen by the new frame)
    assert "lba: cams=4 obs=11 pts=3 fixed_t=1000 fixed_pts=011" in out
    assert "lba2: fixed_t=0011" in out
    # PoseRefiner queues on the host; on this GPU-less box Run() fails loudly and leaves the pose alone
    import torch
    if not torch.cuda.is_available():
        assert "pose: queued=1 max_it=10 status_ok=0 untouched=1" in out
    else:
        assert "pose: queued=1 max_it=10 status_ok=1" in out


# ------------------------------------------------------------------------------------------
# The same boundary EXECUTED on the GPU from C++ (tests/mock/shim_gpu_check.cpp): class BASolver
# (ba_solver.h:14-30) GBA / KGBA / LBA over a mock Map, SiftMatchGPU per pair and the compiled
# FeatureMatching (feature_processing.cc:222-308), compared with the ctypes path and the oracle.
# ------------------------------------------------------------------------------------------
import struct

import numpy as np
import pytest


def _build_gpu_check(tmp_path):
    exe = tmp_path / "shim_gpu_check"
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-Werror", "-fopenmp", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "mock", "shim_gpu_check.cpp"), "-o", str(exe),
                           "-L", os.path.join(ROOT, "xrsfm_b200"), "-lxrsfm_b200",
                           "-Wl,-rpath," + os.path.join(ROOT, "xrsfm_b200")])
    return str(exe)


def _write_scene(path, sc):
    gauge = [int(i) for i in np.flatnonzero(sc.cam_t_fixed)][:2]
    with open(path, "wb") as f:
        f.write(struct.pack("<4i", sc.n_cams, sc.n_pts, gauge[0], gauge[1]))
        f.write(np.asarray(sc.intr[0, :4], dtype="<f8").tobytes())
        order = np.argsort(sc.obs_cam, kind="stable")
        start = np.searchsorted(sc.obs_cam[order], np.arange(sc.n_cams + 1))
        for c in range(sc.n_cams):
            f.write(np.asarray(sc.cam_q[c], dtype="<f8").tobytes())
            f.write(np.asarray(sc.cam_t[c], dtype="<f8").tobytes())
            idx = order[start[c]: start[c + 1]]
            f.write(struct.pack("<i", len(idx)))
            for o in idx:
                f.write(np.asarray(sc.obs_uv[o], dtype="<f8").tobytes())
                f.write(struct.pack("<i", int(sc.obs_pt[o])))
        f.write(np.asarray(sc.pts, dtype="<f8").tobytes())


def _read_result(path, sc):
    raw = np.fromfile(path, dtype=np.uint8)
    n = sc.n_cams * 7 + sc.n_pts * 3
    vals = raw[: n * 8].view("<f8")
    cams = vals[: sc.n_cams * 7].reshape(sc.n_cams, 7)
    pts = vals[sc.n_cams * 7:].reshape(sc.n_pts, 3)
    nw = int(raw[n * 8: n * 8 + 4].view("<i4")[0])
    window = raw[n * 8 + 4: n * 8 + 4 + 4 * nw].view("<i4").copy()
    return cams[:, :4].copy(), cams[:, 4:].copy(), pts.copy(), window


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["gba", "kgba"])
def test_cpp_basolver_on_gpu_equals_ctypes_and_oracle(tmp_path, mode):
    from tests import oracle_lib as ol
    from xrsfm_b200 import ba, synth
    exe = _build_gpu_check(tmp_path)
    sc = synth.make_sphere_scene(14, 900, 6, 123, behind_frac=0.0)
    _write_scene(tmp_path / "scene.bin", sc)
    subprocess.check_call([exe, "ba", str(tmp_path / "scene.bin"), str(tmp_path / "out.bin"), mode])
    q, t, X, _ = _read_result(tmp_path / "out.bin", sc)
    opts = ol.GBA_ACCURATE if mode == "gba" else ol.KGBA
    got = sc.copy_state()
    ba.BASolver().solve_scene(got, **opts)
    ref = sc.copy_state()
    ol.ba_solve(ref, ol.ba_options(**opts))
    scale = max(1.0, np.abs(ref.pts).max())
    # same library underneath; the C++ flattening numbers the points in first-seen order, so sums differ in the last bits
    assert np.abs(q - got.cam_q).max() < 1e-8 and np.abs(t - got.cam_t).max() < 1e-7 and np.abs(X - got.pts).max() < 1e-7 * scale
    assert np.abs(q - ref.cam_q).max() < 1e-5 and np.abs(t - ref.cam_t).max() < 1e-5 * scale
    assert np.abs(X - ref.pts).max() < 1e-5 * scale
    assert np.abs(X - sc.pts).max() > 1e-6  # it did move


@pytest.mark.gpu
def test_cpp_lba_window_on_gpu_equals_oracle(tmp_path):
    """BASolver::LBA (ba_solver.cc:523-591): the window comes from the map walks (mock stand-ins), the flat
    problem from FlattenLBA, the solve from the GPU; the oracle solves the same window problem."""
    from tests import oracle_lib as ol
    from xrsfm_b200 import synth
    exe = _build_gpu_check(tmp_path)
    sc = synth.make_sphere_scene(10, 400, 5, 321, behind_frac=0.0)
    _write_scene(tmp_path / "scene.bin", sc)
    subprocess.check_call([exe, "ba", str(tmp_path / "scene.bin"), str(tmp_path / "out.bin"), "lba"])
    q, t, X, window = _read_result(tmp_path / "out.bin", sc)
    frame_id = sc.n_cams - 1
    # the mock's window: covisibility counts over the tracks of the new frame, stable by frame id
    tracks_new = set(sc.obs_pt[sc.obs_cam == frame_id].tolist())
    cov = np.zeros(sc.n_cams, dtype=int)
    for c, p in zip(sc.obs_cam, sc.obs_pt):
        if p in tracks_new:
            cov[c] += 1
    ranked = sorted([c for c in range(sc.n_cams) if cov[c] > 0], key=lambda c: (-cov[c], c))
    ids1 = ranked[:4]
    ids2 = [frame_id] + [c for c in ranked[:5] if c != frame_id][:3]
    local = sorted(set(ids1) | set(ids2))
    assert local == window.tolist()
    gauge = [int(i) for i in np.flatnonzero(sc.cam_t_fixed)][:2]
    fixed_t = {g for g in gauge if g in local} or set(ids2[-2:])
    keep = np.isin(sc.obs_cam, local)
    cam_map = {c: i for i, c in enumerate(local)}
    pts_used = sorted(set(sc.obs_pt[keep].tolist()))
    pt_map = {p: i for i, p in enumerate(pts_used)}
    sub = sc.copy_state()
    sub["cam_q"] = np.ascontiguousarray(sc.cam_q[local])
    sub["cam_t"] = np.ascontiguousarray(sc.cam_t[local])
    sub["pts"] = np.ascontiguousarray(sc.pts[pts_used])
    sub["cam_intr"] = np.ascontiguousarray(sc.cam_intr[local])
    sub["obs_cam"] = np.array([cam_map[c] for c in sc.obs_cam[keep]], dtype=np.int32)
    sub["obs_pt"] = np.array([pt_map[p] for p in sc.obs_pt[keep]], dtype=np.int32)
    sub["obs_uv"] = np.ascontiguousarray(sc.obs_uv[keep])
    sub["cam_q_fixed"] = np.zeros(len(local), dtype=np.uint8)
    sub["cam_t_fixed"] = np.array([1 if c in fixed_t else 0 for c in local], dtype=np.uint8)
    sub["pt_fixed"] = np.array([0 if p in tracks_new else 1 for p in pts_used], dtype=np.uint8)  # ba_solver.cc:380-382
    sub.n_cams, sub.n_pts, sub.n_obs = len(local), len(pts_used), int(keep.sum())
    for k in ("n_cams", "n_pts", "n_obs"):
        sub[k] = getattr(sub, k)
    ol.ba_solve(sub, ol.ba_options(max_iterations=5, function_tolerance=1e-4, parameter_tolerance=1e-5))
    scale = max(1.0, np.abs(sub.pts).max())
    assert np.abs(q[local] - sub.cam_q).max() < 1e-5 and np.abs(t[local] - sub.cam_t).max() < 1e-5 * scale
    assert np.abs(X[pts_used] - sub.pts).max() < 1e-5 * scale
    outside = [c for c in range(sc.n_cams) if c not in local]
    np.testing.assert_array_equal(q[outside], sc.cam_q[outside])
    assert np.abs(q[local] - sc.cam_q[local]).max() > 0


@pytest.mark.gpu
def test_cpp_matcher_facade_and_feature_matching_on_gpu(tmp_path):
    from tests import oracle_lib as ol
    from xrsfm_b200 import synth
    exe = _build_gpu_check(tmp_path)
    imgs, _ = synth.make_images(5, 700, seed=31)
    pairs = [(0, 1), (1, 2), (0, 4), (2, 3), (3, 4), (0, 2)]
    with open(tmp_path / "desc.bin", "wb") as f:
        f.write(struct.pack("<2i", len(imgs), len(pairs)))
        for im in imgs:
            f.write(struct.pack("<i", im.shape[0]))
            f.write(np.ascontiguousarray(im, dtype=np.uint8).tobytes())
        for a, b in pairs:
            f.write(struct.pack("<2i", a, b))
    subprocess.check_call([exe, "match", str(tmp_path / "desc.bin"), str(tmp_path / "out.bin")])
    raw = np.fromfile(tmp_path / "out.bin", dtype="<i4")
    pos = 0
    expected = [ol.match_pair(imgs[a], imgs[b]) for a, b in pairs]
    for exp in expected:  # (a) per-pair façade: bit-exact lists
        n = int(raw[pos]); pos += 1
        got = raw[pos: pos + 2 * n].reshape(n, 2); pos += 2 * n
        assert got.shape == exp.shape and (got == exp).all()
    nk = int(raw[pos]); pos += 1  # (b) FeatureMatching: >= 15 matches, stand-in verification keeps the even ones
    kept = []
    for (a, b), exp in zip(pairs, expected):
        n = exp.shape[0]
        inl = (n + 1) // 2
        if n >= 15 and inl >= max(15, int(0.25 * n)):
            kept.append((a, b, exp[::2]))
    assert nk == len(kept)
    for a, b, exp in kept:
        id1, id2, n, inl = (int(v) for v in raw[pos: pos + 4]); pos += 4
        got = raw[pos: pos + 2 * n].reshape(n, 2); pos += 2 * n
        assert (id1, id2, n, inl) == (a, b, exp.shape[0], exp.shape[0])
        assert (got == exp).all()
    rc2, n_verified = (int(v) for v in raw[pos: pos + 2])  # (c) matching + batched GPU LO-RANSAC ran end to end
    assert rc2 == 0 and 0 <= n_verified <= len(kept)


@pytest.mark.gpu
def test_cpp_pose_refiner_on_gpu_equals_ctypes(tmp_path):
    """xrsfm_b200::PoseRefiner (shim/pnp_b200.h, the replacement for pnp.cc:38-71) over mock frames, from C++."""
    from xrsfm_b200 import pnp, synth
    exe = _build_gpu_check(tmp_path)
    b = synth.make_pose_batch(9, seed=41, max_pts=120)
    with open(tmp_path / "pose.bin", "wb") as f:
        f.write(struct.pack("<i", 9))
        for p in range(9):
            lo, hi = int(b["offsets"][p]), int(b["offsets"][p + 1])
            f.write(struct.pack("<2i", hi - lo, int(b["intr_model"][p])))
            f.write(b["intr"][p].tobytes() + b["q"][p].tobytes() + b["t"][p].tobytes())
            for k in range(lo, hi):
                f.write(b["uv"][k].tobytes() + b["xyz"][k].tobytes() + struct.pack("<B", int(b["inlier"][k])))
    subprocess.check_call([exe, "pose", str(tmp_path / "pose.bin"), str(tmp_path / "out.bin")])
    got = np.fromfile(tmp_path / "out.bin", dtype="<f8").reshape(9, 9)
    q, t = b["q"].copy(), b["t"].copy()
    sums = pnp.refine_poses(b["offsets"], b["uv"], b["xyz"], b["intr"], b["intr_model"], q, t, inlier_mask=b["inlier"])
    assert np.array_equal(got[:, :4], q) and np.array_equal(got[:, 4:7], t)
    assert np.array_equal(got[:, 7], sums["initial_cost"]) and np.array_equal(got[:, 8], sums["final_cost"])
    assert np.abs(q - b["q"]).max() > 0
