"""PointPillarScatter (reference map_to_bev/pointpillar_scatter.py:5-40): pillar features -> BEV canvas.

Not instantiated by the BEVFusion experiments; provided with the same signature.  The nz == 1 pillar
grid is the D == 1 case of the scatter-to-BEV kernel, so the per-sample Python loop disappears."""
import torch.nn as nn

from unidistill_b200.sparse import scatter_to_dense


class PointPillarScatter(nn.Module):
    def __init__(self, num_bev_features, grid_size):
        super().__init__()
        self.num_bev_features = num_bev_features
        self.nx, self.ny, self.nz = [int(v) for v in grid_size]
        assert self.nz == 1

    def forward(self, pillar_features, coords):
        """pillar_features (M, C); coords (M, 4) [b, z, y, x] -> (B, C*nz, ny, nx)"""
        batch_size = int(coords[:, 0].max().item()) + 1
        volume = scatter_to_dense(pillar_features, coords.int().contiguous(), batch_size,
                                  [self.nz, self.ny, self.nx])
        return volume.view(batch_size, self.num_bev_features * self.nz, self.ny, self.nx)
