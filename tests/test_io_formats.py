"""Wire formats: the C readers/writers of csrc/io_formats.cu against an
independent Python reading of the reference's writers (tests/io_ref.py), byte for byte.
CPU only: the functions under test are host code of libxrsfm_b200.so."""
import ctypes as C
import os

import numpy as np
import pytest

from tests import io_ref
from xrsfm_b200 import _lib, io_formats


def _frames(rng, counts):
    out = []
    for i, n in enumerate(counts):
        out.append(dict(name=f"img_{i:04d}.jpg" if i != 1 else "", keypoints=rng.random((n, 4), dtype=np.float32) * 100,
                        descs=rng.integers(0, 256, (n, 128), dtype=np.uint8)))
    return out


def test_ftr_read_equals_reference_layout(tmp_path):
    rng = np.random.default_rng(1)
    frames = _frames(rng, [5, 0, 17, 1])              # an empty frame and an empty name
    path = str(tmp_path / "ftr.bin")
    io_ref.save_features(path, frames)
    got = io_formats.ReadFeatures(path)
    assert got["names"] == [f["name"] for f in frames]
    np.testing.assert_array_equal(got["row_offsets"], np.cumsum([0] + [len(f["descs"]) for f in frames]))
    np.testing.assert_array_equal(got["descs"], np.concatenate([f["descs"] for f in frames]))
    np.testing.assert_array_equal(got["keypoints"], np.concatenate([f["keypoints"] for f in frames]))
    # descriptors only (what the matcher needs)
    got2 = io_formats.ReadFeatures(path, with_keypoints=False)
    assert got2["keypoints"] is None
    np.testing.assert_array_equal(got2["descs"], got["descs"])


def test_ftr_write_is_byte_identical(tmp_path):
    rng = np.random.default_rng(2)
    frames = _frames(rng, [3, 9, 0])
    ref, ours = str(tmp_path / "ref.bin"), str(tmp_path / "ours.bin")
    io_ref.save_features(ref, frames)
    off = np.cumsum([0] + [len(f["descs"]) for f in frames])
    io_formats.SaveFeatures(ours, [f["name"] for f in frames], off, np.concatenate([f["descs"] for f in frames]),
                            np.concatenate([f["keypoints"] for f in frames]))
    assert open(ref, "rb").read() == open(ours, "rb").read()


def test_ftr_rejects_foreign_and_truncated_files(tmp_path):
    rng = np.random.default_rng(3)
    path = str(tmp_path / "ftr.bin")
    io_ref.save_features(path, _frames(rng, [4, 6]))
    data = open(path, "rb").read()
    cut = str(tmp_path / "cut.bin")
    open(cut, "wb").write(data[: len(data) - 100])
    with pytest.raises(_lib.XrbError, match="truncated"):
        io_formats.ReadFeatures(cut)
    junk = str(tmp_path / "junk.bin")
    open(junk, "wb").write(b"\xff" * 64)
    with pytest.raises(_lib.XrbError):
        io_formats.ReadFeatures(junk)
    with pytest.raises(_lib.XrbError, match="cannot open"):
        io_formats.ReadFeatures(str(tmp_path / "missing.bin"))
    # the reference's reader refuses frames above 1e6 points (io_feature.hpp:61)
    big = str(tmp_path / "big.bin")
    open(big, "wb").write(np.int32(1).tobytes() + b"a\0" + np.int32(1000001).tobytes())
    with pytest.raises(_lib.XrbError, match="1e6"):
        io_formats.ReadFeatures(big)


def _pairs(rng):
    out = []
    for (a, b, m) in [(0, 1, 4), (2, 2, 3), (1, 3, 0), (5, 4, 7)]:  # one self-pair, one empty pair
        matches = [(int(rng.integers(0, 100)), int(rng.integers(0, 100)), float(rng.random())) for _ in range(m)]
        mask = [int(v) for v in rng.integers(0, 2, m)]
        out.append(dict(id1=a, id2=b, matches=matches, E=rng.random((3, 3)), inlier_num=int(sum(mask)), inlier_mask=mask))
    return out


def test_fp_read_drops_self_pairs_like_the_reference(tmp_path):
    rng = np.random.default_rng(4)
    pairs = _pairs(rng)
    path = str(tmp_path / "fp.bin")
    io_ref.save_frame_pairs(path, pairs)
    got = io_formats.ReadFramePairs(path)
    exp = io_ref.read_frame_pairs(path)
    assert got["ids"].tolist() == [[p["id1"], p["id2"]] for p in exp] == [[0, 1], [1, 3], [5, 4]]
    for k, p in enumerate(exp):
        s, e = got["offsets"][k], got["offsets"][k + 1]
        assert got["matches"][s:e].tolist() == [[i, j] for (i, j, _) in p["matches"]]
        np.testing.assert_array_equal(got["distances"][s:e], [d for (_, _, d) in p["matches"]])
        np.testing.assert_array_equal(got["E"][k].reshape(3, 3).T, p["E"])
        assert got["inlier_num"][k] == p["inlier_num"]
        assert got["inlier_mask"][s:e].tolist() == p["inlier_mask"]


def test_fp_write_is_byte_identical_and_defaults(tmp_path):
    rng = np.random.default_rng(5)
    pairs = [p for p in _pairs(rng) if p["id1"] != p["id2"]]
    ref, ours = str(tmp_path / "ref.bin"), str(tmp_path / "ours.bin")
    io_ref.save_frame_pairs(ref, pairs)
    off = np.cumsum([0] + [len(p["matches"]) for p in pairs])
    mm = np.array([[i, j] for p in pairs for (i, j, _) in p["matches"]], dtype=np.uint32).reshape(-1, 2)
    dist = np.array([d for p in pairs for (_, _, d) in p["matches"]])
    E = np.stack([np.asarray(p["E"]).T.reshape(9) for p in pairs])
    io_formats.SaveFramePairs(ours, [[p["id1"], p["id2"]] for p in pairs], off, mm, dist, E,
                              [p["inlier_num"] for p in pairs], np.concatenate([p["inlier_mask"] for p in pairs]))
    assert open(ref, "rb").read() == open(ours, "rb").read()
    # defaults: distance 0 (Match's default), E zeros, all matches inliers
    io_formats.SaveFramePairs(ours, [[p["id1"], p["id2"]] for p in pairs], off, mm)
    back = io_ref.read_frame_pairs(ours)
    assert all(d == 0.0 for p in back for (_, _, d) in p["matches"])
    assert [p["inlier_num"] for p in back] == [len(p["matches"]) for p in pairs]
    assert all(set(p["inlier_mask"]) <= {1} for p in back)


def _model(rng):
    cameras = [dict(id=7, model=2, params=[718.856, 607.19, 185.2, 0.01]),
               dict(id=3, model=4, params=[500.0, 510.0, 320.0, 240.0, 0.1, -0.05, 0.001, 0.002]),
               dict(id=9, model=0, params=[400.0, 100.0, 80.0])]
    tracks = [dict(id=t, xyz=rng.random(3).tolist(), error=float(rng.random()), obs=[]) for t in (0, 1, 4, 9)]
    alive = {t["id"] for t in tracks}
    frames = []
    for k, fid in enumerate((10, 2, 5)):
        p2d = []
        for j in range(6):
            tid = [0, -1, 1, 4, 7, 9][(j + k) % 6]   # 7 is not in points3D.bin (an outlier track), -1 untracked
            p2d.append((float(rng.random() * 100), float(rng.random() * 100), tid))
            if tid in alive:
                next(t for t in tracks if t["id"] == tid)["obs"].append((fid, j))
        q = rng.standard_normal(4)
        q /= np.linalg.norm(q)
        frames.append(dict(id=fid, q_wxyz=q.tolist(), t=rng.random(3).tolist(), camera_id=[7, 3, 7][k],
                           name=f"f{fid}.png", p2d=p2d))
    return cameras, frames, tracks


def test_colmap_model_to_flat_ba_problem(tmp_path):
    rng = np.random.default_rng(6)
    cameras, frames, tracks = _model(rng)
    d = str(tmp_path) + "/"
    io_ref.write_model(d, cameras, frames, tracks)
    pr = io_formats.ReadColMapDataBinary(d)
    assert (pr.n_cams, pr.n_pts, pr.n_intr) == (3, 4, 3)
    assert pr.camera_ids.tolist() == [7, 3, 9] and pr.intr_model.tolist() == [2, 4, 0]
    for i, c in enumerate(cameras):
        np.testing.assert_array_equal(pr.intr[i, : len(c["params"])], c["params"])
        assert np.all(pr.intr[i, len(c["params"]):] == 0)
    assert pr.frame_ids.tolist() == [10, 2, 5] and pr.cam_intr.tolist() == [0, 1, 0]
    for i, fr in enumerate(frames):
        w, x, y, z = fr["q_wxyz"]
        np.testing.assert_array_equal(pr.cam_q[i], [x, y, z, w])     # Eigen coeffs order
        np.testing.assert_array_equal(pr.cam_t[i], fr["t"])
    assert pr.track_ids.tolist() == [0, 1, 4, 9]
    np.testing.assert_array_equal(pr.pts, [t["xyz"] for t in tracks])
    # observations: SetUp's walk (ba_solver.cc:336-352): frames in order, p2d in order, track present
    tix = {t["id"]: k for k, t in enumerate(tracks)}
    exp = [(i, tix[tid], x, y, j) for i, fr in enumerate(frames) for j, (x, y, tid) in enumerate(fr["p2d"]) if tid in tix]
    assert pr.n_obs == len(exp) == 12
    assert pr.obs_cam.tolist() == [e[0] for e in exp] and pr.obs_pt.tolist() == [e[1] for e in exp]
    np.testing.assert_array_equal(pr.obs_uv, [[e[2], e[3]] for e in exp])
    assert pr.obs_p2d.tolist() == [e[4] for e in exp]


def test_colmap_write_updated_roundtrip(tmp_path):
    rng = np.random.default_rng(7)
    cameras, frames, tracks = _model(rng)
    d_in, d_out, d_ref = str(tmp_path / "in") + "/", str(tmp_path / "out") + "/", str(tmp_path / "ref") + "/"
    for d in (d_in, d_out, d_ref):
        os.makedirs(d)
    io_ref.write_model(d_in, cameras, frames, tracks)
    pr = io_formats.ReadColMapDataBinary(d_in)
    # unchanged state -> identical files
    io_formats.WriteColMapDataBinary(d_in, d_out, pr)
    for name in ("cameras.bin", "images.bin", "points3D.bin"):
        assert open(d_in + name, "rb").read() == open(d_out + name, "rb").read(), name
    # a "BA result": new poses and points -> what the reference's writer would produce from them
    pr["cam_t"] += 0.5
    pr["pts"] *= 2.0
    pr["cam_q"][:] = pr["cam_q"][:, [1, 0, 3, 2]] * [1, -1, 1, -1]
    for i, fr in enumerate(frames):
        x, y, z, w = pr["cam_q"][i]
        fr["q_wxyz"], fr["t"] = [w, x, y, z], pr["cam_t"][i].tolist()
    for k, t in enumerate(tracks):
        t["xyz"] = pr["pts"][k].tolist()
    io_ref.write_model(d_ref, cameras, frames, tracks)
    io_formats.WriteColMapDataBinary(d_in, d_out, pr)
    for name in ("cameras.bin", "images.bin", "points3D.bin"):
        assert open(d_ref + name, "rb").read() == open(d_out + name, "rb").read(), name
    # the natural "write back after BA": same directory in and out, and an output directory that does not exist yet
    io_formats.WriteColMapDataBinary(d_in, d_in, pr)
    d_new = str(tmp_path / "made" / "deeper") + "/"
    io_formats.WriteColMapDataBinary(d_in, d_new, pr)
    for name in ("cameras.bin", "images.bin", "points3D.bin"):
        assert open(d_ref + name, "rb").read() == open(d_in + name, "rb").read(), name
        assert open(d_ref + name, "rb").read() == open(d_new + name, "rb").read(), name
    assert sorted(os.listdir(d_in)) == ["cameras.bin", "images.bin", "points3D.bin"]  # no temporaries left


def test_colmap_rejects_inconsistent_models(tmp_path):
    rng = np.random.default_rng(8)
    cameras, frames, tracks = _model(rng)
    d = str(tmp_path) + "/"
    frames[1]["camera_id"] = 99                         # camera missing from cameras.bin
    io_ref.write_model(d, cameras, frames, tracks)
    with pytest.raises(_lib.XrbError, match="camera 99"):
        io_formats.ReadColMapDataBinary(d)
    cameras[0]["model"] = 5                              # unknown model id
    cameras[0]["params"] = [1.0] * 4
    io_ref.CAM_PARAMS[5] = 4
    try:
        io_ref.write_model(d, cameras, frames, tracks)
    finally:
        del io_ref.CAM_PARAMS[5]
    with pytest.raises(_lib.XrbError, match="model"):
        io_formats.ReadColMapDataBinary(d)


def _scene_to_model(sc):
    """A synthetic BA scene as the three model files would hold it (frames list their 2-D points
    in the scene's observation order; a few untracked points are mixed in)."""
    cameras = [dict(id=100 + k, model=int(sc.intr_model[k]),
                    params=sc.intr[k, : io_ref.CAM_PARAMS[int(sc.intr_model[k])]].tolist()) for k in range(sc.n_intr)]
    p2d = [[] for _ in range(sc.n_cams)]
    obs_of_track = [[] for _ in range(sc.n_pts)]
    for o in range(sc.n_obs):
        c, p = int(sc.obs_cam[o]), int(sc.obs_pt[o])
        if len(p2d[c]) % 5 == 2:
            p2d[c].append((1.0, 2.0, -1))                      # an untracked keypoint
        obs_of_track[p].append((c + 50, len(p2d[c])))
        p2d[c].append((float(sc.obs_uv[o, 0]), float(sc.obs_uv[o, 1]), 1000 + p))
    frames = [dict(id=c + 50, q_wxyz=[sc.cam_q[c, 3], *sc.cam_q[c, :3]], t=sc.cam_t[c].tolist(),
                   camera_id=100 + int(sc.cam_intr[c]), name=f"{c}.png", p2d=p2d[c]) for c in range(sc.n_cams)]
    tracks = [dict(id=1000 + p, xyz=sc.pts[p].tolist(), error=0.0, obs=obs_of_track[p]) for p in range(sc.n_pts)]
    return cameras, frames, tracks


def test_model_files_feed_the_same_bundle_adjustment(tmp_path):
    """scene -> cameras/images/points3D.bin -> flat problem -> CPU oracle: same costs, same solution
    as the scene itself (the observation order differs; the problem must not)."""
    from tests import oracle_lib as ol
    from xrsfm_b200 import ba, synth
    sc = synth.make_scene("C1", scale=0.25)
    d = str(tmp_path) + "/"
    io_ref.write_model(d, *_scene_to_model(sc))
    pr = io_formats.ReadColMapDataBinary(d)
    assert (pr.n_cams, pr.n_pts, pr.n_obs, pr.n_intr) == (sc.n_cams, sc.n_pts, sc.n_obs, sc.n_intr)
    np.testing.assert_array_equal(pr.cam_q, sc.cam_q)
    np.testing.assert_array_equal(pr.pts, sc.pts)
    np.testing.assert_array_equal(pr.intr, sc.intr)
    # same multiset of observations
    a = np.lexsort((sc.obs_pt, sc.obs_cam))
    b = np.lexsort((pr.obs_pt, pr.obs_cam))
    np.testing.assert_array_equal(sc.obs_cam[a], pr.obs_cam[b])
    np.testing.assert_array_equal(sc.obs_pt[a], pr.obs_pt[b])
    np.testing.assert_array_equal(sc.obs_uv[a], pr.obs_uv[b])
    pr["cam_t_fixed"][:] = sc.cam_t_fixed                      # the gauge GBA fixes (ba_solver.cc:611-614)
    ba.make_problem(pr)                                          # dtypes / contiguity the GPU engine insists on
    opts = ol.ba_options(**ol.GBA_ACCURATE)
    assert ol.ba_cost(pr, opts) == pytest.approx(ol.ba_cost(sc, opts), rel=1e-12)
    s1, s2 = ol.ba_solve(sc, opts), ol.ba_solve(pr, opts)
    assert s1.termination_type == s2.termination_type and s1.num_lm_iterations == s2.num_lm_iterations
    assert s2.final_cost == pytest.approx(s1.final_cost, rel=1e-9)
    np.testing.assert_allclose(pr.cam_t, sc.cam_t, rtol=1e-7, atol=1e-9)
    np.testing.assert_allclose(pr.pts, sc.pts, rtol=1e-7, atol=1e-9)
    # and back to files: the updated model re-reads as the optimised state
    out = str(tmp_path / "out") + "/"
    os.makedirs(out)
    io_formats.WriteColMapDataBinary(d, out, pr)
    again = io_formats.ReadColMapDataBinary(out)
    np.testing.assert_array_equal(again.cam_q, pr.cam_q)
    np.testing.assert_array_equal(again.cam_t, pr.cam_t)
    np.testing.assert_array_equal(again.pts, pr.pts)
    np.testing.assert_array_equal(again.obs_uv, pr.obs_uv)


def test_random_roundtrips(tmp_path):
    """Randomised write -> read round trips of both match-side formats (sizes, empty frames,
    names with spaces, empty pairs)."""
    rng = np.random.default_rng(99)
    for it in range(25):
        n = int(rng.integers(0, 6))
        counts = rng.integers(0, 40, n)
        off = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
        descs = rng.integers(0, 256, (int(off[-1]), 128), dtype=np.uint8)
        kps = rng.random((int(off[-1]), 4), dtype=np.float32)
        names = ["".join(rng.choice(list("ab /._-0"), int(rng.integers(0, 12)))) for _ in range(n)]
        path = str(tmp_path / f"f{it}.bin")
        io_formats.SaveFeatures(path, names, off, descs, kps)
        got = io_formats.ReadFeatures(path)
        assert got["names"] == names
        np.testing.assert_array_equal(got["row_offsets"], off)
        np.testing.assert_array_equal(got["descs"], descs)
        np.testing.assert_array_equal(got["keypoints"], kps)

        P = int(rng.integers(0, 7))
        ids = np.array([[i, i + 1 + int(rng.integers(0, 3))] for i in range(P)], dtype=np.int32).reshape(P, 2)
        mcount = rng.integers(0, 30, P)
        moff = np.concatenate([[0], np.cumsum(mcount)]).astype(np.int64)
        T = int(moff[-1])
        mm = rng.integers(0, 5000, (T, 2)).astype(np.uint32)
        dist = rng.random(T)
        E = rng.random((P, 9))
        mask = rng.integers(0, 2, T).astype(np.int8)
        inl = np.array([mask[moff[k]: moff[k + 1]].sum() for k in range(P)], dtype=np.int32)
        path = str(tmp_path / f"p{it}.bin")
        io_formats.SaveFramePairs(path, ids, moff, mm, dist, E, inl, mask)
        back = io_formats.ReadFramePairs(path)
        np.testing.assert_array_equal(back["ids"], ids)
        np.testing.assert_array_equal(back["offsets"], moff)
        np.testing.assert_array_equal(back["matches"], mm.view(np.int32))
        np.testing.assert_array_equal(back["distances"], dist)
        np.testing.assert_array_equal(back["E"], E)
        np.testing.assert_array_equal(back["inlier_num"], inl)
        np.testing.assert_array_equal(back["inlier_mask"], mask)
