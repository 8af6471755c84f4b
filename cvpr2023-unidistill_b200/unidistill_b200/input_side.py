"""Input side of the LiDAR branch on the GPU (host side of csrc/input_side.cu).

One launch sequence per batch replaces, per sample, CollectLidarSweeps.forward (reference
unidistill/data/multisensorfusion/transforms3d.py:379-414: sweeps transformed into the key frame in float64, time lag into
the 5th column), ObjectRangeFilter.mask_points_by_range (transforms3d.py:247-256,271-276) and the zero-padding collate of the
points (data/multisensorfusion/nuscenes_multimodal.py:418-470).  The batch stays ragged -- `(flat rows, per-sample offsets)`,
the layout the voxelizer takes -- so no padded origin points are created, copied or voxelized.
"""
import numpy as np
import torch

from . import native


def sweep_to_key_matrix(key_lidar_to_ego, key_ego_to_global, sweep_lidar_to_ego):
    """inv(key_lidar_to_ego) @ inv(key_ego_to_global) @ sweep_ego_to_global @ key_lidar_to_ego in float64, exactly the product
    of transforms3d.py:395-401 (the reference names the sweep's matrix `sweep_lidar_to_ego` and uses it as ego->global)."""
    k_l2e = np.asarray(key_lidar_to_ego, np.float64)
    return np.linalg.inv(k_l2e) @ np.linalg.inv(np.asarray(key_ego_to_global, np.float64)) @ \
        np.asarray(sweep_lidar_to_ego, np.float64) @ k_l2e


def pack_sweeps(samples, pin=True):
    """samples: list (batch) of dict(points (N0,C) float32 array = key frame, sweep_points [ (Ni,C) arrays ],
    sweep_mats [ (4,4) float64 sweep->key matrices ], sweep_time_lags [ seconds ]).  -> dict of HOST tensors ready for one
    host->device copy each: points (n,C), seg_offsets (n_seg+1), sample_first_seg (B+1), mats (n_seg,3,4) f64, time_lag."""
    segs, mats, lags, first = [], [], [], [0]
    for s in samples:
        key = np.ascontiguousarray(s["points"], np.float32)
        segs.append(key)
        mats.append(np.eye(4)[:3])
        lags.append(0.0)                                        # key frame: time column forced to 0 (transforms3d.py:388-389)
        for pts, m, lag in zip(s.get("sweep_points", []), s.get("sweep_mats", []), s.get("sweep_time_lags", [])):
            segs.append(np.ascontiguousarray(pts, np.float32))
            mats.append(np.asarray(m, np.float64)[:3])
            lags.append(float(lag))
        first.append(len(segs))
    C = segs[0].shape[1]
    offs = np.zeros(len(segs) + 1, np.int64)
    np.cumsum([x.shape[0] for x in segs], out=offs[1:])
    out = {"points": torch.from_numpy(np.concatenate(segs, 0) if segs else np.zeros((0, C), np.float32)),
           "seg_offsets": torch.from_numpy(offs.astype(np.int32)),
           "sample_first_seg": torch.tensor(first, dtype=torch.int32),
           "mats": torch.from_numpy(np.stack(mats).astype(np.float64)),
           "time_lag": torch.tensor(lags, dtype=torch.float32)}
    return {k: (v.pin_memory() if pin and torch.cuda.is_available() else v) for k, v in out.items()}


def collect_sweeps(packed, point_cloud_range):
    """packed: the dict of pack_sweeps() on the device.  -> (flat (n_kept, C) f32, offsets (B+1) int32 DEVICE): the merged,
    x/y-range-filtered points of every sample, rows in the reference's order (key frame, then sweep by sweep)."""
    pts = packed["points"]
    native.require_cuda(pts, packed["seg_offsets"], packed["mats"])
    lib = native.load()
    n, C = int(pts.shape[0]), int(pts.shape[1])
    n_seg = int(packed["seg_offsets"].shape[0]) - 1
    B = int(packed["sample_first_seg"].shape[0]) - 1
    dev = pts.device
    out = torch.empty((max(n, 1), C), dtype=torch.float32, device=dev)
    offsets = torch.empty((B + 1,), dtype=torch.int32, device=dev)
    ws = native.WORKSPACE.get(native.query_bytes("ud_collect_sweeps_workspace_bytes", n), dev)
    r = point_cloud_range
    rng = native.host_floats([r[0], r[1], r[3], r[4]])
    rc = lib.ud_collect_sweeps(native.ptr(pts.contiguous()), n, C, native.ptr(packed["seg_offsets"]), n_seg,
                               native.ptr(packed["sample_first_seg"]), B, native.ptr(packed["mats"].contiguous()),
                               native.ptr(packed["time_lag"]), native.host_ptr(rng), native.ptr(out), native.ptr(offsets),
                               native.ptr(ws), ws.numel(), native.stream_ptr(dev))
    native.check(rc, "ud_collect_sweeps")
    return out, offsets


def split_ragged(flat, offsets):
    """(flat, DEVICE offsets) -> list of per-sample (N_i, C) views -- the `points` list LidarEncoder / Voxelization take.
    Reads the B+1 offsets back (one small host sync)."""
    offs = offsets.cpu().tolist()
    return [flat[offs[b]: offs[b + 1]] for b in range(len(offs) - 1)]
