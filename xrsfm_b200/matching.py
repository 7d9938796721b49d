"""Host mirror of the reference's descriptor-matching surface (path M).

Mirrors, name for name and argument for argument:

* ``SiftMatchGPU`` — 3rdparty/SiftGPU/SiftGPU.h:277-372 as exercised by
  src/feature/feature_processing.cc:53-88,118-154 (``SetLanguage``, ``VerifyContextGL``,
  ``Allocate``, ``SetDescriptors``, ``GetSiftMatch``, ``gpu_index``).
* ``SiftMatch`` (uint8 overload) — feature_processing.cc:118-154: hard-wires
  distance_th = 0.7, max_ratio = 0.8, mutual best match, max_match = 16384.
* ``feature_matching_descriptors`` — the descriptor half of ``FeatureMatching``
  (feature_processing.cc:222-255) on the batched C-ABI entry point; the OpenMP
  F-matrix RANSAC half (feature_processing.cc:256-296) stays with the caller.

Everything runs through ``libxrsfm_b200.so``; numpy arrays are only host buffers.
"""
import ctypes as C

import numpy as np

from . import _lib

MAX_MATCH = 16384      # feature_processing.cc:23  `const int max_match = 16384;`
DISTANCE_TH = 0.7      # feature_processing.cc:122
MAX_RATIO = 0.8        # feature_processing.cc:123


def _u8(desc):
    a = np.ascontiguousarray(desc, dtype=np.uint8)
    if a.ndim != 2 or a.shape[1] != 128:
        raise ValueError("descriptors must be [n, 128] uint8 (src/base/types.h:9-10)")
    return a


class SiftMatchGPU:
    """Same call surface as the reference façade; backed by the sm_100a engine."""

    SIFTMATCH_SAME_AS_SIFTGPU = 0
    SIFTMATCH_GLSL = 2
    SIFTMATCH_CUDA = 3
    SIFTMATCH_CUDA_DEVICE0 = 3

    def __init__(self, max_sift=4096):
        self.gpu_index = 0
        self._max_sift = max_sift
        self._device = 0
        self._h = None
        self._keep = [None, None]

    # -- context / allocation (feature_processing.cc:65-85) -----------------------------
    def SetLanguage(self, gpu_language):
        if gpu_language >= self.SIFTMATCH_CUDA_DEVICE0:
            self._device = gpu_language - self.SIFTMATCH_CUDA_DEVICE0

    def VerifyContextGL(self):
        """1 when the CUDA engine can be created on the selected device, else 0."""
        try:
            self._ensure()
            return 1
        except _lib.XrbError:
            return 0

    CreateContextGL = VerifyContextGL

    def Allocate(self, max_sift, mbm):
        self._max_sift = max_sift
        self._destroy()
        try:
            self._ensure()
            return True
        except _lib.XrbError:
            return False

    def SetMaxSift(self, max_sift):
        self.Allocate(max_sift, True)

    def GetMaxSift(self):
        self._ensure()
        return _lib.lib().xrb_match_max_features(self._h)

    def set_variant(self, variant):
        self._ensure()
        return _lib.lib().xrb_match_set_variant(self._h, variant)

    # -- matching -----------------------------------------------------------------------
    def SetDescriptors(self, index, num, descriptors, id=-1):
        self._ensure()
        d = _u8(descriptors)
        num = min(int(num), d.shape[0])
        index = 1 if index > 1 else (0 if index < 0 else index)
        self._keep[index] = d
        rc = _lib.lib().xrb_match_set_descriptors(self._h, index, num, d.ctypes.data, id)
        _lib.check(rc, "xrb_match_set_descriptors")

    def GetSiftMatch(self, max_match, distmax=0.7, ratiomax=0.8, mutual_best_match=1):
        """Returns (n, match_buffer[n,2] uint32); n == -1 on a CUDA error like the reference."""
        self._ensure()
        buf = np.zeros((max(int(max_match), 1), 2), dtype=np.uint32)
        n = _lib.lib().xrb_match_get(self._h, int(max_match), buf.ctypes.data, distmax, ratiomax,
                                     int(mutual_best_match))
        return n, buf[: max(n, 0)]

    # -- batched surface ----------------------------------------------------------------
    def upload_images(self, descs):
        """descs: list of [n_i,128] uint8 arrays -> resident in HBM."""
        self._ensure()
        arrs = [_u8(d) for d in descs]
        counts = np.array([a.shape[0] for a in arrs], dtype=np.int32)
        ptrs = (C.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])
        rc = _lib.lib().xrb_match_upload_images(self._h, len(arrs), counts.ctypes.data, ptrs)
        _lib.check(rc, "xrb_match_upload_images")
        self._counts = np.minimum(counts.astype(np.int64), self.GetMaxSift())

    def upload_packed(self, row_offsets, block):
        self._ensure()
        off = np.ascontiguousarray(row_offsets, dtype=np.int64)
        blk = np.ascontiguousarray(block, dtype=np.uint8)
        rc = _lib.lib().xrb_match_upload_packed(self._h, len(off) - 1, off.ctypes.data,
                                                blk.ctypes.data)
        _lib.check(rc, "xrb_match_upload_packed")
        self._counts = np.minimum(np.diff(off), self.GetMaxSift())

    def upload_ftr(self, file_name):
        """Descriptors of the reference's ftr.bin (io_feature.hpp:76-100) straight to HBM."""
        self._ensure()
        _lib.check(_lib.lib().xrb_match_upload_ftr(self._h, str(file_name).encode()), "xrb_match_upload_ftr")
        self._counts = None  # sizes live in the file: fall back to the allocation's worst case

    def match_pairs(self, pairs, distmax=DISTANCE_TH, ratiomax=MAX_RATIO, mutual_best_match=1,
                    max_match=MAX_MATCH, capacity=None):
        """pairs [P,2] int32 -> (offsets[P+1] int64, matches[total,2] uint32)."""
        self._ensure()
        pr = np.ascontiguousarray(pairs, dtype=np.int32).reshape(-1, 2)
        n = pr.shape[0]
        off = np.zeros(n + 1, dtype=np.int64)
        if capacity is None:
            # at most min(n1, n2, max_match) matches per pair: bound by the resident image sizes when they are
            # known (upload_images / upload_packed), not by the worst case of the allocation
            counts = getattr(self, "_counts", None)
            if counts is not None and n and pr.min() >= 0 and pr.max() < len(counts):
                capacity = int(np.minimum(np.minimum(counts[pr[:, 0]], counts[pr[:, 1]]), max_match).sum())
            else:
                capacity = n * min(max_match, self.GetMaxSift())
        out = np.zeros((max(capacity, 1), 2), dtype=np.uint32)
        rc = _lib.lib().xrb_match_pairs(self._h, n, pr.ctypes.data, distmax, ratiomax,
                                        int(mutual_best_match), int(max_match), off.ctypes.data,
                                        out.ctypes.data, capacity)
        _lib.check(rc, "xrb_match_pairs")
        return off, out[: off[n]]

    # -- plumbing -----------------------------------------------------------------------
    def _ensure(self):
        if self._h is None:
            h = _lib.lib().xrb_match_create(self._max_sift, self._device)
            if not h:
                raise _lib.XrbError("xrb_match_create failed: " + _lib.last_error())
            self._h = h
            self.gpu_index = self._device

    def _destroy(self):
        if self._h is not None:
            _lib.lib().xrb_match_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self._destroy()
        except Exception:
            pass


def CreateSiftGPUMatcher(sift_match_gpu):
    """feature_processing.cc:53-88: SiftMatchGPU(4096), CUDA device 0, Allocate(max_match, true)."""
    sift_match_gpu.__init__(4096)
    sift_match_gpu.SetLanguage(SiftMatchGPU.SIFTMATCH_CUDA_DEVICE0 + 0)
    if sift_match_gpu.VerifyContextGL() == 0:
        return False
    if not sift_match_gpu.Allocate(MAX_MATCH, True):
        return False
    sift_match_gpu.gpu_index = 0
    return True


def SiftMatch(descs1, descs2, sift_match_gpu, max_ratio=None):
    """feature_processing.cc:118-154 — returns [(id1, id2), ...] as an [n,2] int array.

    Like the reference, the max_ratio argument is ignored and overwritten with 0.8."""
    d1, d2 = _u8(descs1), _u8(descs2)
    sift_match_gpu.SetDescriptors(0, d1.shape[0], d1)
    sift_match_gpu.SetDescriptors(1, d2.shape[0], d2)
    n, buf = sift_match_gpu.GetSiftMatch(MAX_MATCH, DISTANCE_TH, MAX_RATIO, True)
    if n < 0:
        raise _lib.XrbError("GetSiftMatch returned -1: " + _lib.last_error())
    return buf.astype(np.int64)


def feature_matching_descriptors(frame_descs, candidate_pairs, matcher=None):
    """Descriptor stage of FeatureMatching (feature_processing.cc:239-255), batched.

    frame_descs: list of [n_i,128] uint8; candidate_pairs: iterable of (id1, id2).
    Returns one (id1, id2, matches[n,2]) per candidate pair, in candidate order — the
    reference appends every pair to frame_pairs here; the `min_num_matches = 15` skip and
    the F-matrix filter belong to the OpenMP stage that follows (:256-296)."""
    own = matcher is None
    if own:
        matcher = SiftMatchGPU()
        if not CreateSiftGPUMatcher(matcher):
            raise _lib.XrbError("CreateSiftGPUMatcher failed: " + _lib.last_error())
    matcher.upload_images(frame_descs)
    pairs = np.asarray(list(candidate_pairs), dtype=np.int32).reshape(-1, 2)
    off, m = matcher.match_pairs(pairs)
    return [(int(pairs[p, 0]), int(pairs[p, 1]), m[off[p]: off[p + 1]].astype(np.int64))
            for p in range(pairs.shape[0])]
