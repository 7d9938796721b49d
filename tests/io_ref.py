"""Independent restatement of the reference's file writers/readers, used only by tests.

Each function follows the reference statement by statement with `struct`, so that the C
implementation (csrc/io_formats.cu) is checked against a second reading of the same source:
  ftr.bin   SaveFeatures / ReadFeatures    src/utility/io_feature.hpp:76-100 / 37-74
  fp.bin    SaveFramePairs / ReadFramePairs  io_feature.hpp:131-147 / 102-129
  model     WriteCamerasBinary / WriteImagesBinary / WritePoints3DBinary  src/utility/io_ecim.cc:145-224
  primitives write_data / write_name       src/utility/io_base.hpp:39-43, 84-87
"""
import struct

import numpy as np

CAM_PARAMS = {0: 3, 1: 4, 2: 4, 3: 5, 4: 8}  # camera_model.hpp kNumParams


def save_features(path, frames):
    """frames: list of dict(name, keypoints[n,4] f32, descs[n,128] u8)."""
    with open(path, "wb") as f:
        f.write(struct.pack("<i", len(frames)))                       # write_data(file, num_frames)
        for fr in frames:
            f.write(fr["name"].encode() + b"\0")                       # write_name
            n = len(fr["keypoints"])
            f.write(struct.pack("<i", n))
            for k in range(n):                                         # pt.x, pt.y, size, angle
                f.write(struct.pack("<4f", *[float(v) for v in fr["keypoints"][k]]))
            f.write(np.ascontiguousarray(fr["descs"], dtype=np.uint8).tobytes())  # with_descs


def save_frame_pairs(path, pairs):
    """pairs: list of dict(id1, id2, matches[(i, j, dist)], E[3,3] (Eigen: column-major in memory),
    inlier_num, inlier_mask)."""
    with open(path, "wb") as f:
        f.write(struct.pack("<Q", len(pairs)))                         # size_t num_framepairs
        for fp in pairs:
            f.write(struct.pack("<ii", fp["id1"], fp["id2"]))
            f.write(struct.pack("<Q", len(fp["matches"])))
            for (i, j, d) in fp["matches"]:                            # Match{int,int,double}
                f.write(struct.pack("<iid", i, j, d))
            f.write(np.asarray(fp["E"], dtype=np.float64).T.tobytes())  # Eigen Matrix3d is column-major
            f.write(struct.pack("<i", fp["inlier_num"]))
            f.write(bytes(bytearray(int(b) & 0xFF for b in fp["inlier_mask"])))


def read_frame_pairs(path):
    out = []
    with open(path, "rb") as f:
        (n,) = struct.unpack("<Q", f.read(8))
        for _ in range(n):
            id1, id2 = struct.unpack("<ii", f.read(8))
            (m,) = struct.unpack("<Q", f.read(8))
            matches = [struct.unpack("<iid", f.read(16)) for _ in range(m)]
            E = np.frombuffer(f.read(72), dtype=np.float64).reshape(3, 3).T
            (inl,) = struct.unpack("<i", f.read(4))
            mask = list(f.read(m))
            out.append(dict(id1=id1, id2=id2, matches=matches, E=E, inlier_num=inl, inlier_mask=mask))
    return [p for p in out if p["id1"] != p["id2"]]                   # io_feature.hpp:120-126


def write_model(dir_path, cameras, frames, tracks):
    """cameras: list of dict(id, model, params); frames: list of dict(id, q_wxyz, t, camera_id, name,
    p2d[(x, y, track_id or -1)]) (registered ones only); tracks: list of dict(id, xyz, error,
    obs[(frame_id, p2d_id)]) (inliers only)."""
    with open(dir_path + "cameras.bin", "wb") as f:
        f.write(struct.pack("<Q", len(cameras)))
        for c in cameras:
            cx, cy = (c["params"][1], c["params"][2]) if c["model"] in (0, 2) else (c["params"][2], c["params"][3])
            f.write(struct.pack("<II", c["id"], c["model"]))
            f.write(struct.pack("<QQ", int(2 * cx), int(2 * cy)))      # uint64_t w = 2 * cx, h = 2 * cy
            assert len(c["params"]) == CAM_PARAMS[c["model"]]
            f.write(struct.pack(f"<{len(c['params'])}d", *c["params"]))
    with open(dir_path + "images.bin", "wb") as f:
        f.write(struct.pack("<Q", len(frames)))
        for fr in frames:
            f.write(struct.pack("<I", fr["id"]))
            f.write(struct.pack("<4d", *fr["q_wxyz"]))                 # q_vec(w, x, y, z)
            f.write(struct.pack("<3d", *fr["t"]))
            f.write(struct.pack("<I", fr["camera_id"]))
            f.write(fr["name"].encode() + b"\0")
            f.write(struct.pack("<Q", len(fr["p2d"])))
            for (x, y, tid) in fr["p2d"]:
                f.write(struct.pack("<ddQ", x, y, tid & 0xFFFFFFFFFFFFFFFF))  # -1 -> 2^64 - 1
    with open(dir_path + "points3D.bin", "wb") as f:
        f.write(struct.pack("<Q", len(tracks)))
        for tr in tracks:
            f.write(struct.pack("<Q", tr["id"]))
            f.write(struct.pack("<3d", *tr["xyz"]))
            f.write(bytes([0, 0, 0]))
            f.write(struct.pack("<d", tr["error"]))
            f.write(struct.pack("<Q", len(tr["obs"])))
            for (fid, pid) in tr["obs"]:
                f.write(struct.pack("<ii", fid, pid))
