"""Wire formats either side of the two hot paths, host mirror.

Names follow the reference's free functions (src/utility/io_feature.hpp, io_ecim.cc); the data
come back as the flat arrays the engines take (packed descriptor block + row offsets,
`Scene`-like dict for the BA), not as Frame / FramePair / Map objects.  All parsing is done by
libxrsfm_b200.so (csrc/io_formats.cu): these functions only allocate numpy arrays.
"""
import ctypes as C

import numpy as np

from . import _lib


def _p(a):
    return a.ctypes.data if a is not None else None


def _b(path):
    return str(path).encode()


# ---- ftr.bin ---------------------------------------------------------------------------------
def ReadFeatures(file_name, with_keypoints=True):
    """io_feature.hpp:37-74 -> dict(names, row_offsets[n+1], descs[total,128] u8, keypoints[total,4] f32)."""
    lib = _lib.lib()
    n, total, nb = C.c_int32(), C.c_int64(), C.c_int64()
    _lib.check(lib.xrb_ftr_scan(_b(file_name), C.byref(n), C.byref(total), C.byref(nb)), "xrb_ftr_scan")
    off = np.zeros(n.value + 1, dtype=np.int64)
    descs = np.zeros((total.value, 128), dtype=np.uint8)
    kps = np.zeros((total.value, 4), dtype=np.float32) if with_keypoints else None
    names = np.zeros(max(1, nb.value), dtype=np.uint8)
    noff = np.zeros(n.value + 1, dtype=np.int64)
    _lib.check(lib.xrb_ftr_read(_b(file_name), n.value, _p(off), _p(descs), _p(kps), _p(names), _p(noff)), "xrb_ftr_read")
    raw = names.tobytes()
    name_list = [raw[noff[i]: noff[i + 1] - 1].decode("utf-8", "replace") for i in range(n.value)]
    return {"names": name_list, "row_offsets": off, "descs": descs, "keypoints": kps}


def SaveFeatures(file_name, names, row_offsets, descs, keypoints=None):
    """io_feature.hpp:76-100 with with_descs = true (run_matching.cc:31)."""
    off = np.ascontiguousarray(row_offsets, dtype=np.int64)
    n = off.shape[0] - 1
    d = np.ascontiguousarray(descs, dtype=np.uint8).reshape(-1, 128)
    k = None if keypoints is None else np.ascontiguousarray(keypoints, dtype=np.float32).reshape(-1, 4)
    blob, noff = None, None
    if names is not None:
        enc = [s.encode() + b"\0" for s in names]
        assert len(enc) == n
        blob = np.frombuffer(b"".join(enc) or b"\0", dtype=np.uint8).copy()
        noff = np.concatenate([[0], np.cumsum([len(e) for e in enc])]).astype(np.int64)
    _lib.check(_lib.lib().xrb_ftr_write(_b(file_name), n, _p(off), _p(d), _p(k), _p(blob), _p(noff)), "xrb_ftr_write")


# ---- fp.bin ----------------------------------------------------------------------------------
def ReadFramePairs(file_name):
    """io_feature.hpp:102-129 (self-pairs dropped) -> dict(ids[P,2], offsets[P+1], matches[T,2],
    distances[T], E[P,3,3] (as stored: column-major), inlier_num[P], inlier_mask[T])."""
    lib = _lib.lib()
    npairs, total = C.c_int64(), C.c_int64()
    _lib.check(lib.xrb_fp_scan(_b(file_name), C.byref(npairs), C.byref(total)), "xrb_fp_scan")
    P, T = npairs.value, total.value
    ids = np.zeros((P, 2), dtype=np.int32)
    off = np.zeros(P + 1, dtype=np.int64)
    mm = np.zeros((T, 2), dtype=np.int32)
    dist = np.zeros(T, dtype=np.float64)
    E = np.zeros((P, 9), dtype=np.float64)
    inl = np.zeros(P, dtype=np.int32)
    mask = np.zeros(T, dtype=np.int8)
    _lib.check(lib.xrb_fp_read(_b(file_name), P, _p(ids), _p(off), _p(mm), _p(dist), _p(E), _p(inl), _p(mask)),
               "xrb_fp_read")
    return {"ids": ids, "offsets": off, "matches": mm, "distances": dist, "E": E, "inlier_num": inl,
            "inlier_mask": mask}


def SaveFramePairs(file_name, ids, offsets, matches, distances=None, E=None, inlier_num=None, inlier_mask=None):
    """io_feature.hpp:131-147; `matches` may be the uint32 lists of SiftMatchGPU.match_pairs."""
    ids = np.ascontiguousarray(ids, dtype=np.int32).reshape(-1, 2)
    off = np.ascontiguousarray(offsets, dtype=np.int64)
    mm = np.ascontiguousarray(matches).reshape(-1, 2)
    mm = mm.view(np.int32) if mm.dtype == np.uint32 else np.ascontiguousarray(mm, dtype=np.int32)
    d = None if distances is None else np.ascontiguousarray(distances, dtype=np.float64)
    e = None if E is None else np.ascontiguousarray(E, dtype=np.float64).reshape(-1, 9)
    n = None if inlier_num is None else np.ascontiguousarray(inlier_num, dtype=np.int32)
    k = None if inlier_mask is None else np.ascontiguousarray(inlier_mask, dtype=np.int8)
    _lib.check(_lib.lib().xrb_fp_write(_b(file_name), ids.shape[0], _p(ids), _p(off), _p(mm), _p(d), _p(e), _p(n), _p(k)),
               "xrb_fp_write")


# ---- COLMAP-style model ----------------------------------------------------------------------
class ColmapProblem(dict):
    """The flat BA problem of a model directory; keys follow xrsfm_b200.synth.Scene so that
    ba.BASolver.solve_scene / load accept it directly."""

    __getattr__ = dict.__getitem__

    def copy_state(self):
        out = ColmapProblem(self)
        for k in ("cam_q", "cam_t", "pts"):
            out[k] = self[k].copy()
        return out


def ReadColMapDataBinary(dir_path):
    """io_ecim.cc:84-87 -> ColmapProblem (cam_q in Eigen coefficient order x, y, z, w)."""
    lib = _lib.lib()
    sz = _lib.ColmapSizes()
    _lib.check(lib.xrb_colmap_scan(_b(dir_path), C.byref(sz)), "xrb_colmap_scan")
    C_, P_, O_, K_ = sz.n_frames, sz.n_points, sz.n_obs, sz.n_cameras
    pr = ColmapProblem(
        cam_q=np.zeros((C_, 4)), cam_t=np.zeros((C_, 3)), pts=np.zeros((P_, 3)),
        intr=np.zeros((K_, 8)), intr_model=np.zeros(K_, dtype=np.int32), cam_intr=np.zeros(C_, dtype=np.int32),
        obs_cam=np.zeros(O_, dtype=np.int32), obs_pt=np.zeros(O_, dtype=np.int32), obs_uv=np.zeros((O_, 2)),
        cam_q_fixed=np.zeros(C_, dtype=np.uint8), cam_t_fixed=np.zeros(C_, dtype=np.uint8),
        pt_fixed=np.zeros(P_, dtype=np.uint8),
        frame_ids=np.zeros(C_, dtype=np.int32), camera_ids=np.zeros(K_, dtype=np.int32),
        track_ids=np.zeros(P_, dtype=np.uint64), obs_p2d=np.zeros(O_, dtype=np.int32))
    pr.update(n_cams=C_, n_pts=P_, n_obs=O_, n_intr=K_)
    pr["_sizes"] = sz
    prob = _flat(pr)
    _lib.check(lib.xrb_colmap_read_problem(_b(dir_path), C.byref(sz), C.byref(prob), _p(pr["frame_ids"]),
                                           _p(pr["camera_ids"]), _p(pr["track_ids"]), _p(pr["obs_p2d"])),
               "xrb_colmap_read_problem")
    return pr


def WriteColMapDataBinary(dir_in, dir_out, pr):
    """The model of dir_in with the poses / points of `pr` (io_ecim.cc:226-232 after a BA)."""
    prob = _flat(pr)
    _lib.check(_lib.lib().xrb_colmap_write_updated(_b(dir_in), _b(dir_out), C.byref(pr["_sizes"]), C.byref(prob)),
               "xrb_colmap_write_updated")


def _flat(pr):
    from .ba import make_problem
    return make_problem(pr)
