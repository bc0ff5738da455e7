"""Seeded synthetic worlds, point clouds and query sets.

`forest()` mimics the reference's simulator (`map_generator/random_forest`): uniformly placed axis-aligned
pillars, width U[0.5,1.2] m, height U[0.8,3.0] m (launch/simulator.xml:26-29, simulator2.xml:28-31 of the
reference), sampled at voxel centres, stored as float32 xyz with PCL PointXYZ stride 16.
Pure numpy; no GPU, no oracle.
"""
import numpy as np

CONFIGS = {
    # name: origin, size, resolution (SURVEY.md §8d / BASELINE.md §3)
    "C1": dict(origin=(-5.0, -5.0, -0.01), size=(50.0, 50.0, 10.0), resolution=0.2, seed=1, n_pillars=120),
    "C2": dict(origin=(-25.6, -25.6, -0.01), size=(51.2, 51.2, 12.8), resolution=0.1, seed=2, n_pillars=700),
    "C3": dict(origin=(-12.8, -12.8, -0.01), size=(25.6, 25.6, 6.4), resolution=0.1, seed=4, n_pillars=60),
    "C4": dict(origin=(-25.6, -25.6, -0.01), size=(51.2, 51.2, 12.8), resolution=0.1, seed=2, n_pillars=700),
}


def pack_cloud(xyz, point_step=16):
    """float32 (n,3) -> PointCloud2-style byte blob with the given stride (x,y,z at offsets 0,4,8)."""
    xyz = np.ascontiguousarray(xyz, dtype=np.float32)
    n = len(xyz)
    if point_step == 12:
        return xyz.copy()
    out = np.zeros((n, point_step // 4), dtype=np.float32)
    out[:, :3] = xyz
    return out


def forest(seed, origin, size, resolution, n_pillars, ground=False, n_points=None, keep_clear=()):
    """Returns float32 (n,3) points. keep_clear: list of (xyz, radius) around which no pillar is placed."""
    rng = np.random.default_rng(seed)
    o = np.asarray(origin, dtype=np.float64)
    s = np.asarray(size, dtype=np.float64)
    r = float(resolution)
    pts = []
    placed = 0
    guard = 0
    while placed < n_pillars and guard < 50 * n_pillars:
        guard += 1
        w = rng.uniform(0.5, 1.2)
        h = rng.uniform(0.8, 3.0)
        cx = rng.uniform(o[0] + 1.0, o[0] + s[0] - 1.0)
        cy = rng.uniform(o[1] + 1.0, o[1] + s[1] - 1.0)
        if any(np.hypot(cx - c[0], cy - c[1]) < rad + w for c, rad in keep_clear):
            continue
        placed += 1
        i0, i1 = int(np.floor((cx - w / 2 - o[0]) / r)), int(np.ceil((cx + w / 2 - o[0]) / r))
        j0, j1 = int(np.floor((cy - w / 2 - o[1]) / r)), int(np.ceil((cy + w / 2 - o[1]) / r))
        k1 = max(1, int(np.ceil(min(h, s[2] - 2 * r) / r)))
        ii, jj, kk = np.meshgrid(np.arange(i0, i1), np.arange(j0, j1), np.arange(0, k1), indexing="ij")
        p = np.stack([ii.ravel(), jj.ravel(), kk.ravel()], 1).astype(np.float64)
        pts.append((p + 0.5) * r + o)
    if ground:
        nx, ny = int(round(s[0] / r)), int(round(s[1] / r))
        ii, jj = np.meshgrid(np.arange(nx), np.arange(ny), indexing="ij")
        p = np.stack([ii.ravel(), jj.ravel(), np.zeros(ii.size)], 1).astype(np.float64)
        pts.append((p + 0.5) * r + o)
    xyz = np.concatenate(pts, 0) if pts else np.zeros((0, 3))
    if n_points is not None:
        if len(xyz) >= n_points:
            sel = rng.choice(len(xyz), n_points, replace=False)
            xyz = xyz[sel]
        else:
            extra = xyz[rng.integers(0, len(xyz), n_points - len(xyz))]
            extra = extra + rng.uniform(-0.5 * r, 0.5 * r, extra.shape)
            xyz = np.concatenate([xyz, extra], 0)
        xyz = xyz[rng.permutation(len(xyz))]
    return xyz.astype(np.float32)


def uniform_cloud(seed, origin, size, n_points, margin=0.0):
    rng = np.random.default_rng(seed)
    o = np.asarray(origin, dtype=np.float64) - margin
    s = np.asarray(size, dtype=np.float64) + 2 * margin
    return (o + rng.random((n_points, 3)) * s).astype(np.float32)


def config_cloud(name, n_points=None, ground=None):
    c = CONFIGS[name]
    ground = (name in ("C2", "C4")) if ground is None else ground
    return forest(c["seed"], c["origin"], c["size"], c["resolution"], c["n_pillars"], ground=ground, n_points=n_points,
                  keep_clear=[((0.0, 0.0, 0.0), 1.0), ((40.0, 40.0, 4.0), 1.0)] if name == "C1" else ())


def random_queries(seed, occ_fn, origin, size, n, min_dist=5.0, z_range=None, margin=0.5):
    """n (start, goal) pairs uniform over free inflated space; occ_fn(pos(n,3)) -> int8 {-1,0,1}."""
    rng = np.random.default_rng(seed)
    o = np.asarray(origin, dtype=np.float64)
    s = np.asarray(size, dtype=np.float64)
    lo = o + margin
    hi = o + s - margin
    if z_range is not None:
        lo[2], hi[2] = z_range
    starts, goals = [], []
    need = n
    while need > 0:
        m = max(1024, 3 * need)
        a = lo + rng.random((m, 3)) * (hi - lo)
        b = lo + rng.random((m, 3)) * (hi - lo)
        ok = (occ_fn(a) == 0) & (occ_fn(b) == 0) & (np.linalg.norm(a - b, axis=1) >= min_dist)
        a, b = a[ok][:need], b[ok][:need]
        starts.append(a)
        goals.append(b)
        need -= len(a)
    return np.concatenate(starts, 0), np.concatenate(goals, 0)


# ---- depth frames: pillars + ground rendered through a pinhole camera ----
def pillar_boxes(seed, origin, size, n_pillars):
    """(n,6) axis-aligned boxes [x0,y0,z0,x1,y1,z1], same distributions as forest()."""
    rng = np.random.default_rng(seed)
    o, s = np.asarray(origin, dtype=np.float64), np.asarray(size, dtype=np.float64)
    w = rng.uniform(0.5, 1.2, n_pillars)
    h = rng.uniform(0.8, 3.0, n_pillars)
    cx = rng.uniform(o[0] + 1.0, o[0] + s[0] - 1.0, n_pillars)
    cy = rng.uniform(o[1] + 1.0, o[1] + s[1] - 1.0, n_pillars)
    return np.stack([cx - w / 2, cy - w / 2, np.full(n_pillars, o[2]), cx + w / 2, cy + w / 2, o[2] + h], 1)


def look_at(yaw, pitch=0.0):
    """Camera-to-world rotation (row-major 3x3) of an optical frame (x right, y down, z forward) looking along `yaw`
    (rad, about world z) and `pitch` (rad, positive = down), and the matching quaternion (w, x, y, z)."""
    cyw, syw, cp, sp = np.cos(yaw), np.sin(yaw), np.cos(pitch), np.sin(pitch)
    fwd = np.array([cyw * cp, syw * cp, -sp])
    right = np.array([syw, -cyw, 0.0])
    down = np.cross(fwd, right)
    R = np.stack([right, down, fwd], 1)
    tr = np.trace(R)
    if tr > 0:
        k = np.sqrt(tr + 1.0) * 2
        q = np.array([k / 4, (R[2, 1] - R[1, 2]) / k, (R[0, 2] - R[2, 0]) / k, (R[1, 0] - R[0, 1]) / k])
    else:
        i = int(np.argmax(np.diag(R)))
        j, l = (i + 1) % 3, (i + 2) % 3
        k = np.sqrt(1.0 + R[i, i] - R[j, j] - R[l, l]) * 2
        q = np.zeros(4)
        q[0] = (R[l, j] - R[j, l]) / k
        q[1 + i] = k / 4
        q[1 + j] = (R[j, i] + R[i, j]) / k
        q[1 + l] = (R[l, i] + R[i, l]) / k
    return R, q / np.linalg.norm(q)


def render_depth(boxes, ground_z, cam_pos, R, rows, cols, fx, fy, cx, cy, scale=1000.0, max_depth=20.0, dropout=0.0, seed=0):
    """uint16 depth image (z-depth * scale, 0 = no return) of `boxes` + the plane z = ground_z."""
    v, u = np.mgrid[0:rows, 0:cols]
    d_cam = np.stack([(u - cx) / fx, (v - cy) / fy, np.ones_like(u, dtype=np.float64)], -1).reshape(-1, 3)
    d = d_cam @ np.asarray(R).T
    o = np.asarray(cam_pos, dtype=np.float64)
    t_best = np.full(len(d), np.inf)
    with np.errstate(divide="ignore", invalid="ignore"):
        tg = (ground_z - o[2]) / d[:, 2]
        t_best = np.where((tg > 0) & np.isfinite(tg), tg, t_best)
        inv = 1.0 / np.where(d == 0.0, 1e-300, d)
        reach = max_depth * float(np.linalg.norm(d_cam, axis=1).max())  # z-depth < max_depth => range < reach
        fwd = np.asarray(R)[:, 2]
        for b in np.asarray(boxes).reshape(-1, 6):
            if np.linalg.norm(np.maximum(np.maximum(b[:3] - o, o - b[3:]), 0.0)) > reach:
                continue  # cannot produce a return within max_depth
            corners = np.array([[b[i], b[1 + j], b[2 + k]] for i in (0, 3) for j in (0, 3) for k in (0, 3)])
            if ((corners - o) @ fwd).max() <= 0:
                continue  # entirely behind the camera
            t0 = (b[:3] - o) * inv
            t1 = (b[3:] - o) * inv
            lo, hi = np.minimum(t0, t1), np.maximum(t0, t1)
            tn = np.maximum(np.maximum(lo[:, 0], lo[:, 1]), lo[:, 2])
            tf = np.minimum(np.minimum(hi[:, 0], hi[:, 1]), hi[:, 2])
            hit = (tf >= tn) & (tf > 0)
            tn = np.maximum(tn, 0)
            t_best = np.where(hit & (tn < t_best), tn, t_best)
    depth = np.where(np.isfinite(t_best) & (t_best < max_depth), t_best, 0.0)
    img = np.clip(np.rint(depth * scale), 0, 65535).astype(np.uint16).reshape(rows, cols)
    if dropout > 0:
        img[np.random.default_rng(seed).random((rows, cols)) < dropout] = 0
    return img
