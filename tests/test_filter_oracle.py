"""Post-BA filtering restatement (oracle/ba_oracle.cpp: xro_filter_points3d, the checker of
xrb_ba_filter_points3d) against a plain numpy reading of track_processor.cc:253-349.  CPU only."""
import ctypes as C

import numpy as np

from tests import ba_numpy, oracle_lib as ol
from xrsfm_b200 import synth


def _rot(q, v):  # Eigen q*v for q = (x, y, z, w)
    u, w = q[:3], q[3]
    c = 2 * np.cross(u, v)
    return v + w * c + np.cross(u, c)


def _filter_np(sc, max_re, deg):
    thr = np.deg2rad(deg)
    keep = np.ones(sc.n_obs, dtype=np.uint8)
    outl = np.zeros(sc.n_pts, dtype=np.uint8)
    err = np.zeros(sc.n_pts)
    ang = np.zeros(sc.n_pts)
    n1 = n2 = 0
    for p in range(sc.n_pts):
        obs = [o for o in np.flatnonzero(sc.obs_pt == p)]
        obs.sort(key=lambda o: sc.obs_cam[o])
        if not obs:
            continue
        dele, res = [], 0.0
        for o in obs:
            c = sc.obs_cam[o]
            pc = _rot(sc.cam_q[c], sc.pts[p]) + sc.cam_t[c]
            k = sc.cam_intr[c]
            uv = ba_numpy.project(int(sc.intr_model[k]), sc.intr[k], np.array([pc[0] / pc[2], pc[1] / pc[2]]))
            re = np.linalg.norm(uv - sc.obs_uv[o])
            if re > max_re or pc[2] < 1e-3 or pc[2] > 1e3:
                dele.append(o)
            else:
                res += re
        if len(dele) >= len(obs) - 1:
            n1 += len(obs)
            outl[p] = 1
            keep[obs] = 0
            continue
        n1 += len(dele)
        keep[dele] = 0
        left = [o for o in obs if keep[o]]
        err[p] = res / len(left)
        ctr = []
        for o in left:
            c = sc.cam_q[sc.obs_cam[o]]
            qi = np.array([-c[0], -c[1], -c[2], c[3]]) / (c @ c)
            ctr.append(-_rot(qi, sc.cam_t[sc.obs_cam[o]]))
        mx, done = 0.0, False
        for i in range(len(ctr)):
            for j in range(i + 1, len(ctr)):
                b2 = ((ctr[i] - ctr[j]) ** 2).sum()
                r1, r2 = ((sc.pts[p] - ctr[i]) ** 2).sum(), ((sc.pts[p] - ctr[j]) ** 2).sum()
                den = 2 * np.sqrt(r1 * r2)
                a = 0.0
                if den != 0:
                    a = abs(np.arccos((r1 + r2 - b2) / den))
                    a = min(a, np.pi - a)
                if a > mx:
                    mx = a
                    if mx > thr:
                        done = True
                        break
            if done:
                break
        ang[p] = mx
        if mx < thr:
            outl[p] = 1
            n2 += 1
            keep[obs] = 0
    return keep, outl, err, ang, (n1, n2)


def test_filter_equals_numpy_reading():
    sc = synth.make_scene("C1", scale=0.1)            # outliers, behind-camera points and noise included
    # a point far away (tiny triangulation angle) and a point with one good observation only
    far = int(sc.obs_pt[0])
    sc.pts[far] *= 400.0
    h = ol.load()
    h.xro_filter_points3d.restype = C.c_int
    h.xro_filter_points3d.argtypes = [C.c_void_p, C.c_double, C.c_double] + [C.c_void_p] * 5
    prob = ol.ba_problem(sc)
    keep = np.zeros(sc.n_obs, dtype=np.uint8)
    outl = np.zeros(sc.n_pts, dtype=np.uint8)
    err = np.zeros(sc.n_pts)
    ang = np.zeros(sc.n_pts)
    cnt = np.zeros(2, dtype=np.int32)
    for max_re, deg in ((8.0, 2.0), (2.0, 0.5)):      # th_rpe_gba-like and a tight setting
        assert h.xro_filter_points3d(C.byref(prob), max_re, deg, keep.ctypes.data, outl.ctypes.data, err.ctypes.data,
                                     ang.ctypes.data, cnt.ctypes.data) == 0
        k2, o2, e2, a2, c2 = _filter_np(sc, max_re, deg)
        np.testing.assert_array_equal(keep, k2)
        np.testing.assert_array_equal(outl, o2)
        live = (o2 == 0) | (a2 > 0)
        np.testing.assert_allclose(err[o2 == 0], e2[o2 == 0], rtol=1e-10)
        np.testing.assert_allclose(ang[live], a2[live], rtol=1e-9, atol=1e-12)
        assert tuple(cnt) == c2
        assert outl[far] == 1                            # the far point goes, one way or the other
        assert 0 < outl.sum() < sc.n_pts and 0 < keep.sum() < sc.n_obs
