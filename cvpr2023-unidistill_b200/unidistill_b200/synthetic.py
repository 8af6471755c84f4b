"""Synthetic nuScenes-shaped inputs (no dataset on the box): LiDAR sweeps, a 6-camera rig, GT boxes.

Shapes follow the reference's collate_fn contract (unidistill/data/multisensorfusion/
nuscenes_multimodal.py:418-495): points (B, N, 5) [x,y,z,intensity,dt], imgs (B,1,6,3,256,704),
mats_dict{sensor2ego_mats,intrin_mats,ida_mats,sensor2sensor_mats:(B,1,6,4,4), bda_mat:(B,4,4)},
gt_boxes (B,M,9) zero padded [x,y,z,dx,dy,dz,yaw,vx,vy], gt_labels (B,M).
Everything is generated on the host with numpy from a seed.
"""
import math

import numpy as np
import torch

# benchmark geometry: 0.1 m voxels -> [1024,1024,40] grid -> 128x128 BEV at stride 8
PC_RANGE_128 = [-51.2, -51.2, -5.0, 51.2, 51.2, 3.0]
VOXEL_SIZE_128 = [0.1, 0.1, 0.2]
GRID_SIZE_128 = [1024, 1024, 40]
# reference default (exps/.../_base_/base_nuscenes_cfg.py:1-5): 180x180 BEV
PC_RANGE_180 = [-54.0, -54.0, -5.0, 54.0, 54.0, 3.0]
VOXEL_SIZE_180 = [0.075, 0.075, 0.2]
GRID_SIZE_180 = [1440, 1440, 40]

# nuScenes class size priors (dx, dy, dz) in metres, 10 classes in CLASS_NAMES order
CLASS_NAMES = ["car", "truck", "construction_vehicle", "bus", "trailer", "barrier", "motorcycle", "bicycle",
               "pedestrian", "traffic_cone"]
_SIZE_PRIORS = np.array([
    [4.6, 1.95, 1.73], [6.9, 2.5, 2.8], [6.4, 2.8, 3.2], [11.0, 2.9, 3.5], [12.3, 2.9, 3.9],
    [0.5, 2.5, 1.0], [2.1, 0.8, 1.5], [1.7, 0.6, 1.3], [0.7, 0.7, 1.8], [0.4, 0.4, 1.1]], np.float32)


def lidar_sweep(rng, n_points, boxes=None, sensor_height=1.84, max_range=60.0):
    """One spinning-LiDAR sweep: 32 beams x azimuth steps, ground-plane hits, a share of box returns."""
    beams = 32
    az_steps = max(1, int(math.ceil(n_points / beams)))
    elev = np.deg2rad(np.linspace(-30.7, 10.7, beams, dtype=np.float64))
    az = np.linspace(0.0, 2 * np.pi, az_steps, endpoint=False)
    E, A = np.meshgrid(elev, az, indexing="ij")
    E, A = E.reshape(-1)[:n_points], A.reshape(-1)[:n_points]
    # range to the ground plane for downward beams, far returns otherwise
    with np.errstate(divide="ignore"):
        r_ground = np.where(E < -0.01, sensor_height / np.maximum(np.tan(-E), 1e-6), max_range)
    r = np.minimum(r_ground, max_range * (0.35 + 0.65 * rng.random(E.shape)))
    r = r + rng.normal(0.0, 0.02, r.shape)
    x = r * np.cos(E) * np.cos(A)
    y = r * np.cos(E) * np.sin(A)
    z = sensor_height - 1.84 + r * np.sin(E)
    pts = np.stack([x, y, z], 1)
    if boxes is not None and len(boxes) > 0:
        # 10 % of the returns land on object surfaces
        k = n_points // 10
        which = rng.integers(0, len(boxes), k)
        b = boxes[which]
        local = (rng.random((k, 3)) - 0.5) * b[:, 3:6]
        face = rng.integers(0, 3, k)
        sign = np.where(rng.random(k) < 0.5, -0.5, 0.5)
        local[np.arange(k), face] = sign * b[np.arange(k), 3 + face]
        c, s = np.cos(b[:, 6]), np.sin(b[:, 6])
        gx = b[:, 0] + c * local[:, 0] - s * local[:, 1]
        gy = b[:, 1] + s * local[:, 0] + c * local[:, 1]
        gz = b[:, 2] + local[:, 2]
        sel = rng.choice(n_points, k, replace=False)
        pts[sel] = np.stack([gx, gy, gz], 1)
    intensity = rng.random((n_points, 1))
    return np.concatenate([pts, intensity], 1).astype(np.float32)


def gt_boxes_sample(rng, n_valid=30, max_objs=40, pc_range=PC_RANGE_128):
    """(max_objs, 9) zero padded boxes + (max_objs,) 0-based labels; a few centres lie outside the map."""
    boxes = np.zeros((max_objs, 9), np.float32)
    labels = np.zeros((max_objs,), np.int64)
    n_valid = min(n_valid, max_objs)
    cls = rng.integers(0, 10, n_valid)
    size = _SIZE_PRIORS[cls] * (0.85 + 0.3 * rng.random((n_valid, 3))).astype(np.float32)
    span = pc_range[3] - 2.0
    xy = (rng.random((n_valid, 2)) * 2 - 1) * span
    if n_valid >= 4:
        xy[-1] = [pc_range[3] + 0.6, 3.0]      # centre outside the map (survey quirk 13)
        # not a multiple of half a BEV cell: a key-point landing exactly on a tap boundary makes the
        # relation-loss gradient (1/(norm + 1e-4)) decide on fp32 round-off
        xy[-2] = [-2.0, pc_range[1] - 0.3]
    z = -1.0 + 0.5 * size[:, 2] - 0.84 + rng.normal(0, 0.05, n_valid)
    yaw = (rng.random(n_valid) * 2 - 1) * np.pi
    vel = rng.normal(0, 2.0, (n_valid, 2)) * (cls[:, None] < 8)
    boxes[:n_valid, 0:2] = xy
    boxes[:n_valid, 2] = z
    boxes[:n_valid, 3:6] = size
    boxes[:n_valid, 6] = yaw
    boxes[:n_valid, 7:9] = vel
    labels[:n_valid] = cls
    return boxes, labels


def lidar_sweeps_raw(seed, points_per_sweep=40000, sweeps=10, boxes=None):
    """The raw material of one sample, as a dataset would hand it to the input transforms: `sweeps` clouds (N, 5)
    [x, y, z, intensity, 0] each in its own sweep frame (sweep 0 = key frame), the 4x4 float64 sweep -> key-frame matrices
    (ego motion v * 0.05 s per sweep: a translation) and the time lags in seconds."""
    rng = np.random.default_rng(seed)
    speed = rng.random() * 10.0
    heading = rng.random() * 2 * np.pi
    vel = np.array([np.cos(heading), np.sin(heading)]) * speed
    clouds, mats, lags = [], [], []
    for s in range(sweeps):
        p = lidar_sweep(rng, points_per_sweep, boxes)
        clouds.append(np.concatenate([p, np.zeros((p.shape[0], 1), np.float32)], 1))
        m = np.eye(4)
        m[0, 3], m[1, 3] = -vel[0] * 0.05 * s, -vel[1] * 0.05 * s
        mats.append(m)
        lags.append(0.05 * s)
    return clouds, mats, lags, vel


def lidar_sample(seed, points_per_sweep=40000, sweeps=10, pc_range=PC_RANGE_128, boxes=None):
    """Key frame + (sweeps-1) ego-motion-shifted sweeps, range filtered in x,y: (N, 5) float32."""
    raw, _, _, vel = lidar_sweeps_raw(seed, points_per_sweep, sweeps, boxes)
    clouds = []
    for s, c in enumerate(raw):
        p = c[:, :4].copy()
        p[:, 0] -= vel[0] * 0.05 * s
        p[:, 1] -= vel[1] * 0.05 * s
        dt = np.full((p.shape[0], 1), 0.05 * s, np.float32)
        clouds.append(np.concatenate([p, dt], 1))
    pts = np.concatenate(clouds, 0)
    keep = (pts[:, 0] > pc_range[0]) & (pts[:, 0] < pc_range[3]) & (pts[:, 1] > pc_range[1]) & \
           (pts[:, 1] < pc_range[4])
    return np.ascontiguousarray(pts[keep], dtype=np.float32)


def lidar_raw_batch(batch_size, seed=0, points_per_sweep=40000, sweeps=10, pc_range=PC_RANGE_128, n_boxes=30,
                    max_objs=40):
    """The same samples as lidar_batch(), before the input transforms: a list of per-sample dicts for
    unidistill_b200.input_side.pack_sweeps (key-frame points, sweep clouds, sweep -> key matrices, time lags) + GT."""
    samples, boxes, labels = [], [], []
    for b in range(batch_size):
        rng = np.random.default_rng(1000 + seed * 97 + b)
        bx, lb = gt_boxes_sample(rng, n_boxes, max_objs, pc_range)
        boxes.append(bx)
        labels.append(lb)
        raw, mats, lags, _ = lidar_sweeps_raw(seed * 97 + b, points_per_sweep, sweeps, bx[:n_boxes])
        samples.append(dict(points=raw[0], sweep_points=raw[1:], sweep_mats=mats[1:], sweep_time_lags=lags[1:]))
    return {"samples": samples, "gt_boxes": torch.from_numpy(np.stack(boxes)),
            "gt_labels": torch.from_numpy(np.stack(labels))}


def lidar_batch(batch_size, seed=0, points_per_sweep=40000, sweeps=10, pc_range=PC_RANGE_128, pad=True,
                n_boxes=30, max_objs=40):
    """Returns dict(points (B,Nmax,5) zero padded [or list], gt_boxes (B,M,9), gt_labels (B,M))."""
    clouds, boxes, labels = [], [], []
    for b in range(batch_size):
        rng = np.random.default_rng(1000 + seed * 97 + b)
        bx, lb = gt_boxes_sample(rng, n_boxes, max_objs, pc_range)
        boxes.append(bx)
        labels.append(lb)
        clouds.append(lidar_sample(seed * 97 + b, points_per_sweep, sweeps, pc_range, bx[:n_boxes]))
    out = {"gt_boxes": torch.from_numpy(np.stack(boxes)), "gt_labels": torch.from_numpy(np.stack(labels))}
    if pad:
        nmax = max(c.shape[0] for c in clouds)
        arr = np.zeros((batch_size, nmax, 5), np.float32)   # collate_fn pads with zeros (quirk 1)
        for b, c in enumerate(clouds):
            arr[b, : c.shape[0]] = c
        out["points"] = torch.from_numpy(arr)
    else:
        out["points"] = [torch.from_numpy(c) for c in clouds]
    return out


# ---- cameras -----------------------------------------------------------------------------------
def _rot_z(a):
    c, s = math.cos(a), math.sin(a)
    return np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]], np.float64)


def camera_rig(batch_size, seed=0, final_dim=(256, 704), random_bda=False):
    """6-camera ring (yaw 0, +-55, +-108, 180 deg; f = 1266 px; c = (800, 450); mounted 1.5 m).

    Returns mats_dict with (B,1,6,4,4) sensor2ego/intrin/ida/sensor2sensor and (B,4,4) bda."""
    rng = np.random.default_rng(seed + 7)
    yaws = np.deg2rad([0.0, 55.0, -55.0, 108.0, -108.0, 180.0])
    # camera frame (x right, y down, z forward) -> ego (x forward, y left, z up)
    cam2ego_axes = np.array([[0, 0, 1], [-1, 0, 0], [0, -1, 0]], np.float64)
    s2e = np.zeros((batch_size, 1, 6, 4, 4), np.float64)
    intr = np.zeros_like(s2e)
    ida = np.zeros_like(s2e)
    s2s = np.zeros_like(s2e)
    bda = np.zeros((batch_size, 4, 4), np.float64)
    fH, fW = final_dim
    for b in range(batch_size):
        for c, yaw in enumerate(yaws):
            m = np.eye(4)
            m[:3, :3] = _rot_z(yaw) @ cam2ego_axes
            m[:3, 3] = [1.5 * math.cos(yaw), 1.5 * math.sin(yaw), 1.5]
            s2e[b, 0, c] = m
            k = np.eye(4)
            k[0, 0] = k[1, 1] = 1266.0
            k[0, 2], k[1, 2] = 800.0, 450.0
            intr[b, 0, c] = k
            resize = 0.44 + (rng.random() - 0.5) * 0.02
            a = np.eye(4)
            a[0, 0] = a[1, 1] = resize
            a[0, 3] = -(1600 * resize - fW) / 2.0
            a[1, 3] = -(900 * resize - fH)
            ida[b, 0, c] = a
            s2s[b, 0, c] = np.eye(4)
        m = np.eye(4)
        if random_bda:
            m[:3, :3] = _rot_z((rng.random() - 0.5) * 0.7853) * (0.95 + 0.1 * rng.random())
            if rng.random() < 0.5:
                m[1, 1] *= -1
        bda[b] = m
    f32 = lambda x: torch.from_numpy(x.astype(np.float32))  # noqa: E731
    return {"sensor2ego_mats": f32(s2e), "intrin_mats": f32(intr), "ida_mats": f32(ida),
            "sensor2sensor_mats": f32(s2s), "bda_mat": f32(bda)}


def camera_batch(batch_size, seed=0, final_dim=(256, 704), random_bda=False):
    g = torch.Generator().manual_seed(seed)
    imgs = torch.randn((batch_size, 1, 6, 3, final_dim[0], final_dim[1]), generator=g)
    return {"imgs": imgs, "mats_dict": camera_rig(batch_size, seed, final_dim, random_bda)}
