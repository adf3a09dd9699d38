"""Seeded synthetic lidar scan pairs.  numpy only; no dataset or network needed.

`scan_pair(seed, n_points, kind)` -> (cur, prev) float32 arrays `(N, F)` = [x, y, z, feat...]
(ONCE: F=4, intensity; Waymo: F=5, tanh(intensity), elongation).  `collate` prepends the batch
index the way pcdet's `collate_batch` does (pcdet/datasets/dataset.py:203-208).
"""
import numpy as np

ONCE = dict(range=[-74.88, -74.88, -5.0, 74.88, 74.88, 3.0], voxel=[0.32, 0.32, 8.0], beams=40,
            elev=(-25.0, 15.0), feats=4)
WAYMO = dict(range=[-75.2, -75.2, -2.0, 75.2, 75.2, 4.0], voxel=[0.32, 0.32, 6.0], beams=64,
             elev=(-17.6, 2.4), feats=5)
# The released ONCE model block needs a BEV grid divisible by 4 (two stride-2 sparse convs, then x1/x2/x4 transposed
# convs concatenated, SiamWCA_MAE.py:231-253); 75.2 m / 0.32 m = 470 is not, and no Waymo model YAML is released
# (README.md:21).  "waymo" = the Waymo-shaped scan on the nearest usable grid (472^2, +-75.52 m).
WAYMO_DATASET_RANGE = list(WAYMO["range"])
WAYMO = dict(WAYMO, range=[-75.52, -75.52, -2.0, 75.52, 75.52, 4.0])
SHAPES = {"once": ONCE, "waymo": WAYMO}


def grid_size(shape):
    r, v = np.asarray(shape["range"], np.float64), np.asarray(shape["voxel"], np.float64)
    return np.round((r[3:6] - r[0:3]) / v).astype(np.int64)  # data_processor.py:166-172


def _scene(rng, n_bins=720):
    """per-azimuth-bin obstacle range (inf = none): piecewise vertical walls in ~55% of bins."""
    rng_obst = np.full(n_bins, np.inf)
    i = 0
    while i < n_bins:
        w = int(rng.integers(3, 40))
        if rng.random() < 0.55:
            rng_obst[i:i + w] = np.exp(rng.uniform(np.log(4.0), np.log(110.0)))
        i += w
    height = rng.uniform(0.5, 4.0, n_bins)
    return rng_obst, height


def _cast(rng, shape, n_points, scene, sensor_h=1.8, ground_sigma=0.06):
    rng_obst, height = scene
    n_bins = rng_obst.shape[0]
    n = int(n_points * 1.6)  # over-sample, crop later
    beam = rng.integers(0, shape["beams"], n)
    elev = np.deg2rad(np.linspace(shape["elev"][0], shape["elev"][1], shape["beams"]))[beam]
    az = rng.uniform(0, 2 * np.pi, n)
    b = (az / (2 * np.pi) * n_bins).astype(np.int64) % n_bins
    # ground hit (downward beams) or max range
    with np.errstate(divide="ignore", invalid="ignore"):
        r_ground = np.where(elev < 0, sensor_h / np.tan(-elev), np.inf)
    r_obst = rng_obst[b]
    # obstacle is hit if the ray is below the wall top at that range
    z_at = sensor_h + r_obst * np.tan(elev)
    hit_obst = np.isfinite(r_obst) & (r_obst < r_ground) & (z_at < height[b]) & (z_at > 0)
    # uneven ground / beam divergence: rings are several pillars wide; obstacles have depth
    r_ground = r_ground * np.exp(rng.normal(0, ground_sigma, n))
    r = np.where(hit_obst, r_obst + rng.uniform(0, 1.5, n), r_ground)
    ok = np.isfinite(r) & (r < 120.0)
    r, az, elev = r[ok], az[ok], elev[ok]
    r = r + rng.normal(0, 0.02, r.shape)
    x, y = r * np.cos(az), r * np.sin(az)
    z = r * np.tan(elev)  # sensor frame: ground plane at z = -sensor_h
    return np.stack([x, y, z], 1)


def _finish(rng, shape, xyz, n_points):
    lo, hi = shape["range"][:3], shape["range"][3:]
    m = (xyz[:, 0] >= lo[0]) & (xyz[:, 0] <= hi[0]) & (xyz[:, 1] >= lo[1]) & (xyz[:, 1] <= hi[1])
    xyz = xyz[m]  # mask_points_by_range crops x, y only (common_utils.py:124-127)
    if xyz.shape[0] > n_points:
        xyz = xyz[rng.permutation(xyz.shape[0])[:n_points]]
    n = xyz.shape[0]
    # 0.5% just below and 0.5% above the z range: exercises the trunc-toward-zero rule
    k = max(1, n // 200)
    idx = rng.permutation(n)
    xyz[idx[:k], 2] = rng.uniform(lo[2] - 8.0, lo[2], k)
    xyz[idx[k:2 * k], 2] = rng.uniform(hi[2], hi[2] + 4.0, k)
    feats = [rng.uniform(0, 1, (n, 1))]
    if shape["feats"] == 5:
        feats[0] = np.tanh(feats[0] * 3.0)
        feats.append(rng.uniform(0, 1.5, (n, 1)))
    pts = np.concatenate([xyz] + feats, 1).astype(np.float32)
    return pts[rng.permutation(n)]  # shuffle_points (data_processor.py:92-102)


def scan_pair(seed, n_points=60000, kind="once"):
    shape = SHAPES[kind]
    rng = np.random.default_rng(seed)
    scene = _scene(rng)
    cur = _cast(rng, shape, n_points, scene)
    prev = _cast(rng, shape, n_points, scene)
    yaw = np.deg2rad(rng.uniform(-2, 2))
    t = rng.uniform(-2, 2, 2)
    c, s = np.cos(yaw), np.sin(yaw)
    prev[:, :2] = prev[:, :2] @ np.array([[c, s], [-s, c]]) + t
    prev += rng.normal(0, 0.02, prev.shape)
    return _finish(rng, shape, cur, n_points), _finish(rng, shape, prev, n_points)


def collate(frames):
    """list of (N_i, F) -> (sum N_i, 1+F) with the batch index in column 0."""
    out = []
    for b, p in enumerate(frames):
        out.append(np.concatenate([np.full((p.shape[0], 1), b, np.float32), p], 1))
    return np.concatenate(out, 0)


def batch(first_seed, batch_size, n_points=60000, kind="once"):
    """(points, points_prev) collated float32 arrays for scan indices first_seed .. +batch_size-1."""
    pairs = [scan_pair(first_seed + i, n_points, kind) for i in range(batch_size)]
    return collate([p[0] for p in pairs]), collate([p[1] for p in pairs])


def _quat_pose(rng, yaw_deg, trans):
    """ONCE pose vector [qx, qy, qz, qw, tx, ty, tz] (vehicle -> global; once_utils.py:12-13 reads it so) with small roll/pitch."""
    yaw, pitch, roll = np.deg2rad(yaw_deg), np.deg2rad(rng.uniform(-0.5, 0.5)), np.deg2rad(rng.uniform(-0.5, 0.5))
    cy, sy, cp, sp, cr, sr = np.cos(yaw / 2), np.sin(yaw / 2), np.cos(pitch / 2), np.sin(pitch / 2), np.cos(roll / 2), np.sin(roll / 2)
    q = [sr * cp * cy - cr * sp * sy, cr * sp * cy + sr * cp * sy, cr * cp * sy - sr * sp * cy, cr * cp * cy + sr * sp * sy]
    return np.array(q + list(trans), np.float64)


def _quat_matrix(q):
    x, y, z, w = q / np.linalg.norm(q)
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def raw_scan_pair(seed, n_points=60000, kind="once", static=False):
    """One RAW sample as the ONCE temporal dataset reads it before any processing (once_temporal_dataset.py:159-165):
    dict(points, points_prev (n_points, F) float32 in their OWN sensor frames, not cropped (returns out to 120 m), with
    ~1 % ego-vehicle returns inside |x|,|y| < 2 m; pose, pose_prev (7,) float64 [qx,qy,qz,qw,tx,ty,tz]).  static=True gives
    the all-zero poses ONCE stores for a parked vehicle (once_utils.py:5-6)."""
    shape = SHAPES[kind]
    rng = np.random.default_rng(seed + 77000)
    scene = _scene(rng)

    def frame():
        xyz = _cast(rng, shape, n_points, scene)
        xyz = xyz[rng.permutation(xyz.shape[0])[:n_points]]
        k = max(1, xyz.shape[0] // 100)
        xyz[:k] = np.stack([rng.uniform(-1.9, 1.9, k), rng.uniform(-1.9, 1.9, k), rng.uniform(-1.0, 0.5, k)], 1)
        return xyz[rng.permutation(xyz.shape[0])]

    cur, prev = frame(), frame()
    if static:
        pose = pose_prev = np.zeros(7)
    else:
        base_yaw, base_t = rng.uniform(-180, 180), np.append(rng.uniform(-500, 500, 2), rng.uniform(-2, 2))
        pose = _quat_pose(rng, base_yaw, base_t)
        pose_prev = _quat_pose(rng, base_yaw + rng.uniform(-2, 2), base_t + np.append(rng.uniform(-2, 2, 2), rng.uniform(-0.1, 0.1)))
        # the previous scan sees the same scene from its own pose: current frame -> global -> previous frame
        Rc, Rp = _quat_matrix(pose[:4]), _quat_matrix(pose_prev[:4])
        prev = (prev @ Rc.T + pose[4:] - pose_prev[4:]) @ Rp
    out = {}
    for key, xyz in (("points", cur), ("points_prev", prev)):
        feats = [rng.uniform(0, 1, (xyz.shape[0], 1))]
        if shape["feats"] == 5:
            feats = [np.tanh(feats[0] * 3.0), rng.uniform(0, 1.5, (xyz.shape[0], 1))]
        out[key] = np.concatenate([xyz] + feats, 1).astype(np.float32)
    out["pose"], out["pose_prev"] = pose, pose_prev
    out["frame_id"], out["frame_id_prev"] = str(seed), str(seed) + "p"
    return out
