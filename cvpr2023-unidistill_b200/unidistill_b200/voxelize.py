"""Hard voxelization + mean VFE on the GPU (host side of csrc/voxelize.cu).

Mirrors what reference unidistill/data/det3d/preprocess/voxelization.py:44-73 obtains from
spconv's PointToVoxel, for the whole batch in one launch sequence.
"""
import torch

from . import native


def hard_voxelize(points, voxel_size, point_cloud_range, grid_size, max_num_points, max_voxels,
                  want_voxels=True, want_mean=True):
    """points: list of (N_i, C) float32 CUDA tensors (or one tensor = batch of 1).

    Returns (voxels|None, coords int32 (M,4) [b,z,y,x], num_points int32 (M,), mean|None, counts list)
    where rows are ordered sample by sample and, inside a sample, by first appearance of the voxel.
    """
    if not isinstance(points, (list, tuple)):
        points = [points]
    native.require_cuda(*points)
    lib = native.load()
    batch_size = len(points)
    C = int(points[0].shape[1])
    offsets = [0]
    for p in points:
        if p.dim() != 2 or p.shape[1] != C or p.dtype != torch.float32:
            raise RuntimeError("hard_voxelize: every cloud must be a float32 (N, %d) tensor" % C)
        offsets.append(offsets[-1] + int(p.shape[0]))
    n_total = offsets[-1]
    device = points[0].device
    if batch_size == 1:
        flat = points[0].contiguous()
    else:
        flat = torch.cat([p.reshape(-1, C) for p in points], dim=0)
    cap = max(1, min(n_total, batch_size * int(max_voxels)))
    coords = torch.empty((cap, 4), dtype=torch.int32, device=device)
    num = torch.empty((cap,), dtype=torch.int32, device=device)
    voxels = torch.empty((cap, max_num_points, C), dtype=torch.float32, device=device) if want_voxels else None
    mean = torch.empty((cap, C), dtype=torch.float32, device=device) if want_mean else None
    counts = torch.empty((1 + batch_size,), dtype=torch.int32, device=device)
    ws_bytes = native.query_bytes("ud_voxelize_workspace_bytes", n_total)
    ws = native.WORKSPACE.get(ws_bytes, device)
    offs = native.host_ints(offsets)
    vs = native.host_floats(voxel_size)
    pr = native.host_floats(point_cloud_range)
    gs = native.host_ints(grid_size)
    if native.TIMER.enabled:
        native.TIMER.meta = {"n_points": n_total, "counts": counts, "voxels": want_voxels}
    rc = lib.ud_hard_voxelize(
        native.ptr(flat), native.host_ptr(offs), batch_size, C, native.host_ptr(vs), native.host_ptr(pr),
        native.host_ptr(gs), int(max_num_points), int(max_voxels), cap, native.ptr(voxels),
        native.ptr(coords), native.ptr(num), native.ptr(mean), native.ptr(counts), native.ptr(ws),
        ws.numel(), native.stream_ptr(device))
    native.check(rc, "ud_hard_voxelize")
    counts_host = counts.cpu().tolist()   # the one host sync of the voxelizer
    m = counts_host[0]
    return (
        voxels[:m] if voxels is not None else None,
        coords[:m],
        num[:m],
        mean[:m] if mean is not None else None,
        counts_host[1:],
    )
