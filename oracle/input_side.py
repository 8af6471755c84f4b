"""oracle/input_side.py -- TEST INFRASTRUCTURE (checker only).

numpy restatement of the reference's per-sample input transforms for the LiDAR points:
  collect_lidar_sweeps   <- CollectLidarSweeps.forward, unidistill/data/multisensorfusion/transforms3d.py:379-414
  mask_points_by_range   <- ObjectRangeFilter.mask_points_by_range, transforms3d.py:247-256
  collate_points         <- collate_fn.fill_batch_tensor, data/multisensorfusion/nuscenes_multimodal.py:440-470
Pinned by tests/golden/input_side.npz, produced by the reference's own classes (tests/golden/make_golden.py)."""
import numpy as np


def collect_lidar_sweeps(points, sweep_points, key_ego_to_global, key_lidar_to_ego, sweep_lidar_to_ego, timestamp,
                         sweep_timestamps):
    all_points = np.array(points, copy=True)
    if all_points.shape[-1] == 5:
        all_points[:, -1] = 0.0
    for frame, s2g, ts in zip(sweep_points, sweep_lidar_to_ego, sweep_timestamps):
        frame = np.array(frame, copy=True)
        hom = np.ones((frame.shape[0], 4))
        hom[:, :3] = frame[:, :3]
        on_key = (np.linalg.inv(key_lidar_to_ego) @ np.linalg.inv(key_ego_to_global) @ s2g @ key_lidar_to_ego @ hom.T).T
        frame[:, :3] = on_key[:, :3]
        if all_points.shape[-1] == 5:
            frame[:, -1] = (timestamp - ts) / 1e6
        all_points = np.concatenate([all_points, frame])
    return all_points


def mask_points_by_range(points, limit_range):
    lr = np.asarray(limit_range, np.float32)
    return (points[:, 0] >= lr[0]) & (points[:, 0] <= lr[3]) & (points[:, 1] >= lr[1]) & (points[:, 1] <= lr[4])


def collate_points(clouds):
    """zero padding to the longest cloud (what the GPU path never materialises)."""
    n = max(c.shape[0] for c in clouds)
    out = np.zeros((len(clouds), n, clouds[0].shape[1]), np.float32)
    for i, c in enumerate(clouds):
        out[i, : c.shape[0]] = c
    return out
