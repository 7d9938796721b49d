"""Deterministic synthetic inputs for both hot paths.

Descriptors follow the reference's input contract for path M:
``u8 = clamp(round(512 * sqrt(x / ||x||_1)), 0, 255)``
(L1RootNormalizeFeatureDescriptors src/feature/sift_extractor.cc:99-110 and
FeatureDescriptorsToUnsignedByte src/feature/sift_extractor.h:22-34).

BA scenes are flat SoA problems in the layout of ``xrb_ba_problem`` — what a ``BASolver``
shim produces from ``Map``: SIMPLE_RADIAL cameras
(camera_model.hpp:131-153, the model rec_kitti.cc:25 instantiates), Tcw poses with Eigen
quaternion coefficient order (x, y, z, w), gauge = translations of cameras 0 and 1 fixed
(ba_solver.cc:611-614).
"""
import numpy as np

SEED_BASE = 20260924

# ----------------------------------------------------------------------------------------
# Path M — descriptors and pair lists
# ----------------------------------------------------------------------------------------


def quantize_descriptors(raw):
    """float [n,128] >= 0  ->  uint8 per sift_extractor.cc:99-110 + sift_extractor.h:22-34."""
    raw = np.asarray(raw, dtype=np.float32)
    norm = np.abs(raw).sum(axis=1, keepdims=True)
    norm[norm == 0] = 1.0
    root = np.sqrt(raw / norm)
    scaled = np.round(512.0 * root)  # std::round: half away from zero; values >= 0
    return np.clip(scaled, 0, 255).astype(np.uint8)


def make_pool(n_pool, rng):
    """World descriptors: Gamma(0.6,1) bins with 40 % of the bins zeroed (SIFT-like sparsity)."""
    g = rng.gamma(0.6, 1.0, size=(n_pool, 128)).astype(np.float32)
    g[rng.random((n_pool, 128)) < 0.4] = 0.0
    g[:, 0] += 1e-3  # never an all-zero descriptor
    return g


def make_images(n_images, n_feat, seed, n_pool=None, window=None, noise=0.02):
    """List of [n_feat,128] uint8 arrays; neighbouring images share ~50 % of their features.

    Image i draws n_feat distinct world descriptors from a window of the pool centred at
    i * (n_pool / n_images), perturbs them in the L1-root domain with N(0, noise^2) and
    quantises."""
    rng = np.random.default_rng(seed)
    if window is None:
        window = 2 * n_feat
    if n_pool is None:
        n_pool = max(window + 1, 100 * n_images)
    pool = make_pool(n_pool, rng)
    norm = pool.sum(axis=1, keepdims=True)
    root = np.sqrt(pool / norm)
    images, ids = [], []
    for i in range(n_images):
        centre = int(i * (n_pool / n_images))
        idx = (centre - window // 2 + rng.choice(window, size=n_feat, replace=False)) % n_pool
        d = root[idx] + rng.normal(0.0, noise, size=(n_feat, 128)).astype(np.float32)
        d = np.maximum(d, 0.0)
        q = np.clip(np.round(512.0 * d), 0, 255).astype(np.uint8)
        images.append(q)
        ids.append(idx)
    return images, ids


def random_descriptors(n, rng):
    """Unstructured descriptors (no planted correspondences)."""
    return quantize_descriptors(make_pool(n, rng))


def sequential_pairs(n_images, window=19, n_retrieval=5, seed=0):
    """run_matching.cc:125-151 shape: (i, i+k) for k = 1..window plus pseudo-retrieval
    neighbours; unique, i < j."""
    rng = np.random.default_rng(seed)
    s = set()
    for i in range(n_images):
        for k in range(1, window + 1):
            if i + k < n_images:
                s.add((i, i + k))
        for j in rng.integers(0, n_images, size=n_retrieval):
            j = int(j)
            if j != i:
                s.add((min(i, j), max(i, j)))
    return np.array(sorted(s), dtype=np.int32).reshape(-1, 2)


# ----------------------------------------------------------------------------------------
# Path B — bundle-adjustment scenes
# ----------------------------------------------------------------------------------------

KITTI_SIMPLE_RADIAL = (718.856, 607.1928, 185.27157, 0.0)  # rec_kitti.cc:25


def quat_from_rotmat(R):
    """Rotation matrix -> (x, y, z, w), w >= 0."""
    t = np.trace(R)
    if t > 0:
        s = np.sqrt(t + 1.0) * 2
        w, x, y, z = 0.25 * s, (R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s
    elif R[0, 0] > R[1, 1] and R[0, 0] > R[2, 2]:
        s = np.sqrt(1.0 + R[0, 0] - R[1, 1] - R[2, 2]) * 2
        w, x, y, z = (R[2, 1] - R[1, 2]) / s, 0.25 * s, (R[0, 1] + R[1, 0]) / s, (R[0, 2] + R[2, 0]) / s
    elif R[1, 1] > R[2, 2]:
        s = np.sqrt(1.0 + R[1, 1] - R[0, 0] - R[2, 2]) * 2
        w, x, y, z = (R[0, 2] - R[2, 0]) / s, (R[0, 1] + R[1, 0]) / s, 0.25 * s, (R[1, 2] + R[2, 1]) / s
    else:
        s = np.sqrt(1.0 + R[2, 2] - R[0, 0] - R[1, 1]) * 2
        w, x, y, z = (R[1, 0] - R[0, 1]) / s, (R[0, 2] + R[2, 0]) / s, (R[1, 2] + R[2, 1]) / s, 0.25 * s
    q = np.array([x, y, z, w])
    if w < 0:
        q = -q
    return q / np.linalg.norm(q)


def rotmat_from_quat(q):
    x, y, z, w = q
    return np.array([
        [1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
        [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
        [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def quat_mul(a, b):
    """Hamilton product, (x,y,z,w) order."""
    ax, ay, az, aw = a
    bx, by, bz, bw = b
    return np.array([aw * bx + ax * bw + ay * bz - az * by,
                     aw * by - ax * bz + ay * bw + az * bx,
                     aw * bz + ax * by - ay * bx + az * bw,
                     aw * bw - ax * bx - ay * by - az * bz])


def look_at_pose(centre, target, up=np.array([0.0, 0.0, 1.0])):
    """Tcw (q, t) of a camera at `centre` whose +z axis points at `target`."""
    zc = target - centre
    zc = zc / np.linalg.norm(zc)
    xc = np.cross(zc, up)
    if np.linalg.norm(xc) < 1e-8:
        xc = np.cross(zc, np.array([0.0, 1.0, 0.0]))
    xc /= np.linalg.norm(xc)
    yc = np.cross(zc, xc)
    Rcw = np.stack([xc, yc, zc])  # rows = camera axes in world
    return quat_from_rotmat(Rcw), -Rcw @ centre


def project_simple_radial(Rcw, tcw, X, intr):
    f, cx, cy, k = intr
    pc = X @ Rcw.T + tcw
    z = pc[:, 2]
    with np.errstate(divide="ignore", invalid="ignore"):
        x, y = pc[:, 0] / z, pc[:, 1] / z
    r2 = x * x + y * y
    u = f * (x + x * k * r2) + cx
    v = f * (y + y * k * r2) + cy
    return np.stack([u, v], axis=1), z


class BAScene(dict):
    """dict with attribute access; arrays in xrb_ba_problem layout + ground truth."""
    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__

    def copy_state(self):
        s = BAScene(self)
        for k in ("cam_q", "cam_t", "pts"):
            s[k] = self[k].copy()
        return s


def _finish_scene(rng, Rs, qs, ts, centres, X, obs_cam, obs_pt, intr, width, height,
                  noise_px, outlier_frac, behind_frac, baseline):
    n_cams, n_pts = len(qs), X.shape[0]
    order = np.lexsort((obs_cam, obs_pt))  # caller order: by point then camera (arbitrary but fixed)
    obs_cam, obs_pt = obs_cam[order].astype(np.int32), obs_pt[order].astype(np.int32)
    n_obs = obs_cam.shape[0]
    uv = np.empty((n_obs, 2))
    by_cam = np.argsort(obs_cam, kind="stable")  # one pass instead of one mask per camera
    start = np.searchsorted(obs_cam[by_cam], np.arange(n_cams + 1))
    for c in range(n_cams):
        m = by_cam[start[c]: start[c + 1]]
        if m.size:
            uv[m], _ = project_simple_radial(Rs[c], ts[c], X[obs_pt[m]], intr)
    uv += rng.normal(0.0, noise_px, size=uv.shape)
    n_out = int(round(outlier_frac * n_obs))
    if n_out:
        oi = rng.choice(n_obs, size=n_out, replace=False)
        ang = rng.uniform(0, 2 * np.pi, size=n_out)
        mag = rng.uniform(20.0, 100.0, size=n_out)
        uv[oi] += np.stack([mag * np.cos(ang), mag * np.sin(ang)], axis=1)
    # initial state = ground truth perturbed
    q0 = np.empty((n_cams, 4))
    t0 = np.empty((n_cams, 3))
    for c in range(n_cams):
        if c < 2:  # gauge cameras keep their exact translation; rotation still perturbed
            dt = np.zeros(3)
        else:
            dt = rng.normal(0.0, 0.02 * baseline, size=3)
        w = rng.normal(0.0, np.deg2rad(0.5), size=3)
        half = 0.5 * w
        dq = np.array([half[0], half[1], half[2], 1.0])
        dq /= np.linalg.norm(dq)
        q0[c] = quat_mul(dq, qs[c])
        q0[c] /= np.linalg.norm(q0[c])
        t0[c] = ts[c] + dt
    depth = np.linalg.norm(X[obs_pt] - centres[obs_cam], axis=1)
    mean_depth = np.zeros(n_pts)
    np.add.at(mean_depth, obs_pt, depth)
    cnt = np.bincount(obs_pt, minlength=n_pts).astype(np.float64)
    mean_depth /= np.maximum(cnt, 1)
    X0 = X + rng.normal(0.0, 1.0, size=X.shape) * (0.01 * mean_depth)[:, None]
    n_behind = int(round(behind_frac * n_pts))
    if n_behind:
        first = np.searchsorted(obs_pt, np.arange(n_pts + 1))  # obs_pt is sorted (lexsort above)
        for p in rng.choice(n_pts, size=n_behind, replace=False):
            o = np.arange(first[p], first[p + 1])
            if o.size == 0:
                continue
            c = obs_cam[rng.choice(o)]
            X0[p] = centres[c] - 0.5 * (X[p] - centres[c])  # mirrored behind camera c
    intr_arr = np.zeros((1, 8))
    intr_arr[0, :4] = intr
    cam_t_fixed = np.zeros(n_cams, dtype=np.uint8)
    cam_t_fixed[:2] = 1
    return BAScene(
        n_cams=n_cams, n_pts=n_pts, n_obs=n_obs, n_intr=1,
        cam_q=np.ascontiguousarray(q0), cam_t=np.ascontiguousarray(t0),
        pts=np.ascontiguousarray(X0),
        intr=intr_arr, intr_model=np.array([2], dtype=np.int32),
        cam_intr=np.zeros(n_cams, dtype=np.int32),
        obs_cam=obs_cam, obs_pt=obs_pt, obs_uv=np.ascontiguousarray(uv),
        cam_q_fixed=np.zeros(n_cams, dtype=np.uint8), cam_t_fixed=cam_t_fixed,
        pt_fixed=np.zeros(n_pts, dtype=np.uint8),
        gt_q=np.array(qs), gt_t=np.array(ts), gt_pts=X, width=width, height=height)


def make_sphere_scene(n_cams, n_pts, obs_per_pt, seed, intr=KITTI_SIMPLE_RADIAL,
                      noise_px=0.5, outlier_frac=0.02, behind_frac=1e-4, radius=10.0,
                      half_extent=3.0, width=1241, height=376):
    """C1 / C2 "unordered" scene: cameras on a sphere looking at the origin, points uniform
    in a cube, each point observed by exactly obs_per_pt cameras that see it in frame."""
    rng = np.random.default_rng(seed)
    # Fibonacci sphere (deterministic, even coverage)
    i = np.arange(n_cams) + 0.5
    phi = np.arccos(1 - 2 * i / n_cams)
    theta = np.pi * (1 + 5 ** 0.5) * i
    centres = radius * np.stack([np.cos(theta) * np.sin(phi), np.sin(theta) * np.sin(phi),
                                 np.cos(phi)], axis=1)
    qs, ts, Rs = [], [], []
    for c in range(n_cams):
        q, t = look_at_pose(centres[c], np.zeros(3))
        qs.append(q), ts.append(t), Rs.append(rotmat_from_quat(q))
    X = rng.uniform(-half_extent, half_extent, size=(n_pts, 3))
    k = min(obs_per_pt, n_cams)
    obs_cam = np.empty((n_pts, k), dtype=np.int64)
    chunk = max(1, min(n_pts, 4_000_000 // max(n_cams, 1)))
    for p0 in range(0, n_pts, chunk):
        Xc = X[p0:p0 + chunk]
        score = np.empty((Xc.shape[0], n_cams))
        for c in range(n_cams):
            uvc, z = project_simple_radial(Rs[c], ts[c], Xc, intr)
            inside = (z > 0.1) & (uvc[:, 0] >= 0) & (uvc[:, 0] < width) & (uvc[:, 1] >= 0) & (uvc[:, 1] < height)
            # visible cameras get a random priority in [0,1); others are ranked by distance
            # from the principal point and come after every visible camera
            off = np.hypot(uvc[:, 0] - intr[1], uvc[:, 1] - intr[2])
            score[:, c] = np.where(inside, rng.random(Xc.shape[0]), 10.0 + off)
        obs_cam[p0:p0 + chunk] = np.argsort(score, axis=1)[:, :k]
    obs_pt = np.repeat(np.arange(n_pts), k)
    nn = np.sqrt(4 * np.pi * radius ** 2 / n_cams)
    return _finish_scene(rng, Rs, qs, ts, centres, X, obs_cam.reshape(-1), obs_pt, intr, width,
                         height, noise_px, outlier_frac, behind_frac, baseline=nn)


def make_sequential_scene(n_cams, n_pts, obs_per_pt, seed, intr=KITTI_SIMPLE_RADIAL,
                          noise_px=0.5, outlier_frac=0.02, behind_frac=1e-4, step=1.0,
                          corridor=15.0, max_range=40.0, width=1241, height=376):
    """C4 "sequential" scene (run_kitti_reconstruction sizing): forward-looking cameras 1 m
    apart on a smooth planar trajectory; each point is seen by ~obs_per_pt consecutive frames
    within max_range, which makes the reduced camera system block-banded."""
    rng = np.random.default_rng(seed)
    s = np.arange(n_cams) * step
    heading = 0.6 * np.sin(s / 120.0) + 0.3 * np.sin(s / 37.0)
    dx, dy = np.cos(heading) * step, np.sin(heading) * step
    centres = np.stack([np.cumsum(dx) - dx[0], np.cumsum(dy) - dy[0], np.full(n_cams, 1.6)], axis=1)
    qs, ts, Rs = [], [], []
    for c in range(n_cams):
        fwd = np.array([np.cos(heading[c]), np.sin(heading[c]), 0.0])
        q, t = look_at_pose(centres[c], centres[c] + fwd)
        qs.append(q), ts.append(t), Rs.append(rotmat_from_quat(q))
    k = min(obs_per_pt, n_cams)
    # anchor frame a: the first frame that sees the point; the point sits `ahead` metres in
    # front of frame a + k (so frames a .. a+k-1 all see it in front, within max_range)
    a = rng.integers(0, max(1, n_cams - k), size=n_pts)
    last = np.minimum(a + k - 1, n_cams - 1)
    ahead = rng.uniform(4.0, max_range - k * step - 2.0, size=n_pts)
    lateral = rng.uniform(-corridor, corridor, size=n_pts)
    # keep the bearing inside the horizontal field of view of frame `last`
    max_lat = 0.75 * ahead * (intr[1] / intr[0])
    lateral = np.clip(lateral, -max_lat, max_lat)
    hgt = rng.uniform(-1.4, 0.22 * ahead * (intr[2] / intr[0]) * 2, size=n_pts)
    fwd = np.stack([np.cos(heading[last]), np.sin(heading[last]), np.zeros(n_pts)], axis=1)
    left = np.stack([-fwd[:, 1], fwd[:, 0], np.zeros(n_pts)], axis=1)
    X = centres[last] + fwd * ahead[:, None] + left * lateral[:, None]
    X[:, 2] = 1.6 + np.clip(hgt, -1.5, 3.0)
    obs_cam = (a[:, None] + np.arange(k)[None, :]).clip(0, n_cams - 1)
    obs_pt = np.repeat(np.arange(n_pts), k)
    return _finish_scene(rng, Rs, qs, ts, centres, X, obs_cam.reshape(-1), obs_pt, intr, width,
                         height, noise_px, outlier_frac, behind_frac, baseline=step)


def make_clustered_scene(n_cams, n_pts, obs_per_pt, seed, cams_per_cluster=100, shared_frac=0.08,
                         intr=KITTI_SIMPLE_RADIAL, noise_px=0.5, outlier_frac=0.02, behind_frac=1e-4,
                         radius=10.0, half_extent=3.0, spacing=14.0, width=1241, height=376):
    """C5 "1DSfM-shaped" unordered scene (rec_1dsfm sizing, src/rec_1dsfm.cc:21-55): cameras in clusters (a
    landmark photographed from all around), dense covisibility inside a cluster, a fraction of the points
    also seen from the neighbouring cluster, and ONE CAMERA MODEL PER IMAGE (internet photos: every frame
    has its own intrinsics block, all constant in the reference's BA).  The reduced camera system is block
    sparse: dense diagonal blocks per cluster, couplings between neighbours, fill from the elimination."""
    rng = np.random.default_rng(seed)
    n_cl = max(1, int(round(n_cams / cams_per_cluster)))
    cpc = [n_cams // n_cl + (1 if i < n_cams % n_cl else 0) for i in range(n_cl)]
    ppc = [n_pts // n_cl + (1 if i < n_pts % n_cl else 0) for i in range(n_cl)]
    side = int(np.ceil(np.sqrt(n_cl)))
    qs, ts, Rs, centres, X_all, oc, op = [], [], [], [], [], [], []
    cam0, pt0 = 0, 0
    cl_cams, cl_centre = [], []
    for ci in range(n_cl):
        ctr = np.array([(ci % side) * spacing, (ci // side) * spacing, 0.0])
        nc_, np_ = cpc[ci], ppc[ci]
        i = np.arange(nc_) + 0.5
        phi = np.arccos(1 - 2 * i / nc_)
        theta = np.pi * (1 + 5 ** 0.5) * i
        cc = ctr + radius * np.stack([np.cos(theta) * np.sin(phi), np.sin(theta) * np.sin(phi), np.cos(phi)], axis=1)
        for c in range(nc_):
            q, t = look_at_pose(cc[c], ctr)
            qs.append(q), ts.append(t), Rs.append(rotmat_from_quat(q))
        centres.append(cc)
        X = ctr + rng.uniform(-half_extent, half_extent, size=(np_, 3))
        k = min(obs_per_pt, nc_)
        n_sh = int(round(shared_frac * np_)) if n_cl > 1 else 0
        # own cluster: k random cameras per point (every camera of the cluster sees the landmark volume)
        pick = np.argsort(rng.random((np_, nc_)), axis=1)[:, :k]
        oc.append((cam0 + pick).reshape(-1))
        op.append(np.repeat(pt0 + np.arange(np_), k))
        cl_cams.append((cam0, nc_)), cl_centre.append(ctr)
        X_all.append(X)
        if n_sh and ci > 0:  # the first n_sh points are also seen by 3 cameras of the previous cluster that face them
            pc0, pnc = cl_cams[ci - 1]
            prev_cc = centres[ci - 1]
            facing = np.argsort(np.linalg.norm(prev_cc - ctr, axis=1))[: max(3, pnc // 4)]  # the far side looks this way
            far = np.argsort(-np.linalg.norm(prev_cc - ctr, axis=1))[: max(3, pnc // 4)]
            cand = far  # cameras on the far side of the previous cluster look through their landmark towards this one
            sel = cand[np.argsort(rng.random((n_sh, len(cand))), axis=1)[:, :3]]
            oc.append((pc0 + sel).reshape(-1))
            op.append(np.repeat(pt0 + np.arange(n_sh), 3))
            del facing
        cam0 += nc_
        pt0 += np_
    centres = np.concatenate(centres)
    X = np.concatenate(X_all)
    obs_cam, obs_pt = np.concatenate(oc), np.concatenate(op)
    # keep only observations in front of their camera
    R = np.array(Rs)[obs_cam]
    z = np.einsum("ij,ij->i", R[:, 2, :], X[obs_pt]) + np.array(ts)[obs_cam][:, 2]
    keep = z > 0.5
    obs_cam, obs_pt = obs_cam[keep], obs_pt[keep]
    sc = _finish_scene(rng, Rs, qs, ts, centres, X, obs_cam, obs_pt, intr, width, height, noise_px, outlier_frac,
                       behind_frac, baseline=np.sqrt(4 * np.pi * radius ** 2 / max(cpc)))
    # one camera model per image
    n_c = len(qs)
    sc["intr"] = np.ascontiguousarray(np.repeat(sc["intr"], n_c, axis=0))
    sc["intr_model"] = np.full(n_c, 2, dtype=np.int32)
    sc["cam_intr"] = np.arange(n_c, dtype=np.int32)
    sc["n_intr"] = n_c
    return sc


SCENES = {
    # name: (builder, n_cams, n_pts, obs_per_pt, seed index)
    "C1": (make_sphere_scene, 20, 2_000, 10, 0),
    "C2": (make_sphere_scene, 500, 200_000, 10, 1),
    "C4": (make_sequential_scene, 2_700, 1_000_000, 10, 3),
    "C5": (make_clustered_scene, 5_000, 1_500_000, 8, 4),
}


def make_scene(name, scale=1.0):
    """BASELINE.json configs by name; `scale` shrinks cameras/points for tests."""
    fn, c, p, k, si = SCENES[name]
    return fn(max(4, int(round(c * scale))), max(16, int(round(p * scale))), k, SEED_BASE + si)


# ------------------------------------------------------------------------------------------
# Pose-refinement batches (src/geometry/pnp.cc:38-71): per frame, 2D-3D inlier correspondences
# and a pose as the P3P LORANSAC leaves it (close to the truth, not at the optimum).
# ------------------------------------------------------------------------------------------
POSE_INTRINSICS = {  # model id -> 8 padded parameters (camera_model.hpp:93-210)
    0: [718.856, 607.19, 185.22, 0, 0, 0, 0, 0],
    1: [718.856, 712.3, 607.19, 185.22, 0, 0, 0, 0],
    2: [718.856, 607.19, 185.22, -0.02, 0, 0, 0, 0],
    3: [718.856, 712.3, 607.19, 185.22, -0.02, 0, 0, 0],
    4: [718.856, 712.3, 607.19, 185.22, -0.03, 0.01, 1e-3, -5e-4],
}


def project_model(model, intr, x, y):
    """WorldToImage of camera_model.hpp:93-210 on arrays (ids 0/1 keep the reference's 2 f x + c)."""
    if model == 0:
        return intr[0] * 2 * x + intr[1], intr[0] * 2 * y + intr[2]
    if model == 1:
        return intr[0] * 2 * x + intr[2], intr[1] * 2 * y + intr[3]
    r2 = x * x + y * y
    if model == 2:
        return intr[0] * (x + x * intr[3] * r2) + intr[1], intr[0] * (y + y * intr[3] * r2) + intr[2]
    if model == 3:
        return intr[0] * (x + x * intr[4] * r2) + intr[2], intr[1] * (y + y * intr[4] * r2) + intr[3]
    k1, k2, p1, p2 = intr[4:8]
    rad = k1 * r2 + k2 * r2 * r2
    du = x * rad + 2 * p1 * x * y + p2 * (r2 + 2 * x * x)
    dv = y * rad + 2 * p2 * x * y + p1 * (r2 + 2 * y * y)
    return intr[0] * (x + du) + intr[2], intr[1] * (y + dv) + intr[3]


def make_pose_batch(n_poses, seed, min_pts=20, max_pts=400, noise_px=0.7, outlier_frac=0.05, rot_deg=1.0,
                    trans_frac=0.02, models=(2, 4, 3, 1, 0), with_mask=True, behind_frac=0.0):
    """-> dict(offsets, uv, xyz, inlier, intr, intr_model, q, t, gt_q, gt_t): n_poses independent frames.

    Each frame sees between min_pts and max_pts points 4..40 units in front of it; its starting pose is the truth
    rotated by ~rot_deg and shifted by ~trans_frac of the mean depth; a fraction of the measurements are gross
    outliers (the Huber loss has work to do) and, with `with_mask`, a few correspondences are masked out the way
    SolvePnP_colmap's inlier_mask would.  `behind_frac` puts some points behind the camera (the constant-residual
    branch of the cost functor)."""
    rng = np.random.default_rng(seed)
    counts = rng.integers(min_pts, max_pts + 1, size=n_poses)
    offsets = np.zeros(n_poses + 1, dtype=np.int64)
    np.cumsum(counts, out=offsets[1:])
    total = int(offsets[-1])
    uv = np.empty((total, 2))
    xyz = np.empty((total, 3))
    inlier = np.ones(total, dtype=np.uint8)
    intr = np.zeros((n_poses, 8))
    intr_model = np.zeros(n_poses, dtype=np.int32)
    q = np.empty((n_poses, 4))
    t = np.empty((n_poses, 3))
    gt_q = np.empty((n_poses, 4))
    gt_t = np.empty((n_poses, 3))
    for p in range(n_poses):
        lo, hi = int(offsets[p]), int(offsets[p + 1])
        n = hi - lo
        model = models[p % len(models)]
        intr_model[p] = model
        intr[p] = POSE_INTRINSICS[model]
        centre = rng.normal(0.0, 20.0, size=3)
        target = centre + rng.normal(0.0, 1.0, size=3)
        qg, tg = look_at_pose(centre, target)
        R = rotmat_from_quat(qg)
        depth = rng.uniform(4.0, 40.0, size=n)
        xn = rng.uniform(-0.7, 0.7, size=n)
        yn = rng.uniform(-0.22, 0.22, size=n)
        pc = np.stack([xn * depth, yn * depth, depth], axis=1)
        X = (pc - tg) @ R  # Rcw^T (pc - t)
        u, v = project_model(model, intr[p], xn, yn)
        m = np.stack([u, v], axis=1) + rng.normal(0.0, noise_px, size=(n, 2))
        n_out = int(round(outlier_frac * n))
        if n_out:
            oi = rng.choice(n, size=n_out, replace=False)
            ang = rng.uniform(0, 2 * np.pi, size=n_out)
            mag = rng.uniform(20.0, 100.0, size=n_out)
            m[oi] += np.stack([mag * np.cos(ang), mag * np.sin(ang)], axis=1)
        n_behind = int(round(behind_frac * n))
        if n_behind:
            bi = rng.choice(n, size=n_behind, replace=False)
            X[bi] = centre - 0.5 * (X[bi] - centre)
        if with_mask and n > min_pts:
            inlier[lo + rng.choice(n, size=max(1, n // 16), replace=False)] = 0
        uv[lo:hi], xyz[lo:hi] = m, X
        w = rng.normal(0.0, np.deg2rad(rot_deg), size=3) * 0.5
        dq = np.array([w[0], w[1], w[2], 1.0])
        dq /= np.linalg.norm(dq)
        q0 = quat_mul(dq, qg)
        q[p] = q0 / np.linalg.norm(q0)
        t[p] = tg + rng.normal(0.0, trans_frac * depth.mean(), size=3)
        gt_q[p], gt_t[p] = qg, tg
    return dict(offsets=offsets, uv=np.ascontiguousarray(uv), xyz=np.ascontiguousarray(xyz), inlier=inlier,
                intr=intr, intr_model=intr_model, q=np.ascontiguousarray(q), t=np.ascontiguousarray(t),
                gt_q=gt_q, gt_t=gt_t)


def pose_as_scene(batch, p):
    """Pose p of a make_pose_batch dict as a one-camera BAScene (points constant, masked correspondences dropped):
    the same problem in xrb_ba_problem form, for xrb_ba_solve and the BA oracle."""
    lo, hi = int(batch["offsets"][p]), int(batch["offsets"][p + 1])
    keep = np.flatnonzero(batch["inlier"][lo:hi]) + lo
    n = keep.shape[0]
    return BAScene(
        n_cams=1, n_pts=n, n_obs=n, n_intr=1,
        cam_q=batch["q"][p:p + 1].copy(), cam_t=batch["t"][p:p + 1].copy(),
        pts=np.ascontiguousarray(batch["xyz"][keep]),
        intr=batch["intr"][p:p + 1].copy(), intr_model=batch["intr_model"][p:p + 1].copy(),
        cam_intr=np.zeros(1, dtype=np.int32),
        obs_cam=np.zeros(n, dtype=np.int32), obs_pt=np.arange(n, dtype=np.int32),
        obs_uv=np.ascontiguousarray(batch["uv"][keep]),
        cam_q_fixed=np.zeros(1, dtype=np.uint8), cam_t_fixed=np.zeros(1, dtype=np.uint8),
        pt_fixed=np.ones(n, dtype=np.uint8))
