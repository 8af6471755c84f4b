"""LSSFPN -- lift-splat camera encoder with the reference's constructor / forward contract
(unidistill/layers/blocks_3d/mmdet3d/lss_fpn.py:85-368) on sm_100a kernels.

`voxel_pooling` / `VoxelPooling` keep the reference signature (geom_xyz int32, input_features f32,
voxel_num) and run the sorted segmented reduction of csrc/voxel_pool.cu.  `LSSFPN._forward_single_sweep`
uses the fused lift-splat: softmax(depth) (x) context is consumed inside the pooling kernel, so the
(B,6,112,16,44,256) volume of lss_fpn.py:289-310 is never written.
"""
import torch
from torch import nn

from unidistill_b200.camera import VoxelPooling, lift_splat, lss_geometry_index, voxel_pooling

try:  # the reference builds these through the OpenMMLab registries (lss_fpn.py:2-3,143-149)
    from mmdet.models import build_backbone
    from mmdet3d.models import build_neck
except ImportError:  # not installed on this image: torchvision-based stand-ins with the same key names
    from unidistill_b200.imgnets import build_backbone, build_neck

__all__ = ["LSSFPN", "VoxelPooling", "voxel_pooling"]


class LSSFPN(nn.Module):
    def __init__(self, x_bound, y_bound, z_bound, d_bound, final_dim, downsample_factor, output_channels,
                 img_backbone_conf, img_neck_conf, depth_net_conf, timestamp_net_conf=None):
        super().__init__()
        self.downsample_factor = downsample_factor
        self.d_bound = d_bound
        self.final_dim = final_dim
        self.timestamp_net_conf = timestamp_net_conf
        self.output_channels = output_channels
        bounds = [x_bound, y_bound, z_bound]
        self.register_buffer("voxel_size", torch.Tensor([b[2] for b in bounds]))
        self.register_buffer("voxel_coord", torch.Tensor([b[0] + b[2] / 2.0 for b in bounds]))
        self.register_buffer("voxel_num", torch.LongTensor([round((b[1] - b[0]) / b[2]) for b in bounds]))
        self._voxel_num_host = [int(round((b[1] - b[0]) / b[2])) for b in bounds]
        self.register_buffer("frustum", self.create_frustum())
        self.depth_channels = self.frustum.shape[0]
        self.img_backbone = build_backbone(img_backbone_conf)
        self.img_neck = build_neck(img_neck_conf)
        self.depth_net = self._configure_depth_net(depth_net_conf)
        self.timestamp_net = self._configure_timestamp_net()
        self.img_neck.init_weights()
        self.img_backbone.init_weights()
        self.fuse_lift_splat = True
        self.fuse_geometry = True

    def _configure_timestamp_net(self):
        if self.timestamp_net_conf is None:
            return None
        return nn.Linear(self.timestamp_net_conf["in_channels"], self.timestamp_net_conf["out_channels"])

    def _configure_depth_net(self, depth_net_conf):
        layers = []
        if depth_net_conf.get("num_res_layer", 0) == 0:
            layers.append(nn.Conv2d(depth_net_conf["in_channels"], self.depth_channels + self.output_channels,
                                    kernel_size=1))
        return nn.Sequential(*layers)

    def create_frustum(self):
        """(D, fH, fW, 4) image-plane grid [u, v, depth, 1] (reference :173-198)."""
        img_h, img_w = self.final_dim
        fh, fw = img_h // self.downsample_factor, img_w // self.downsample_factor
        depth = torch.arange(*self.d_bound, dtype=torch.float).view(-1, 1, 1).expand(-1, fh, fw)
        nd = depth.shape[0]
        us = torch.linspace(0, img_w - 1, fw, dtype=torch.float).view(1, 1, fw).expand(nd, fh, fw)
        vs = torch.linspace(0, img_h - 1, fh, dtype=torch.float).view(1, fh, 1).expand(nd, fh, fw)
        return torch.stack((us, vs, depth, torch.ones_like(depth)), -1)

    def get_geometry(self, sensor2ego_mat, intrin_mat, ida_mat, bda_mat):
        """Frustum -> ego coordinates, (B, ncam, D, H, W, 3) (reference :200-240; same op order so that
        the fp32 results -- and hence the truncated voxel indices -- agree with the reference)."""
        batch, ncam = sensor2ego_mat.shape[:2]
        pts = ida_mat.view(batch, ncam, 1, 1, 1, 4, 4).inverse().matmul(self.frustum.unsqueeze(-1))
        pts = torch.cat((pts[:, :, :, :, :, :2] * pts[:, :, :, :, :, 2:3], pts[:, :, :, :, :, 2:]), 5)
        cam2ego = sensor2ego_mat.matmul(torch.inverse(intrin_mat))
        pts = cam2ego.view(batch, ncam, 1, 1, 1, 4, 4).matmul(pts)
        if bda_mat is not None:
            bda = bda_mat.unsqueeze(1).repeat(1, ncam, 1, 1).view(batch, ncam, 1, 1, 1, 4, 4)
            pts = (bda @ pts).squeeze(-1)
        else:
            pts = pts.squeeze(-1)
        return pts[..., :3]

    def get_cam_feats(self, imgs):
        b, sweeps, cams, ch, h, w = imgs.shape
        feats = self.img_neck(self.img_backbone(imgs.flatten().view(b * sweeps * cams, ch, h, w)))[0]
        return feats.reshape(b, sweeps, cams, feats.shape[1], feats.shape[2], feats.shape[3])

    def _forward_depth_net(self, feat, *args, **kwargs):
        return self.depth_net(feat)

    def _forward_voxel_net(self, img_feat_with_depth):
        return img_feat_with_depth

    def _forward_single_sweep(self, sweep_index, sweep_imgs, mats_dict, is_return_depth=False):
        batch, _, ncam = sweep_imgs.shape[:3]
        src = self.get_cam_feats(sweep_imgs)[:, 0, ...]
        depth_feature = self._forward_depth_net(
            src.reshape(batch * ncam, src.shape[2], src.shape[3], src.shape[4]), mats_dict, sweep_index)
        depth = depth_feature[:, : self.depth_channels].softmax(1)
        context = depth_feature[:, self.depth_channels: self.depth_channels + self.output_channels]
        mats = (mats_dict["sensor2ego_mats"][:, sweep_index, ...], mats_dict["intrin_mats"][:, sweep_index, ...],
                mats_dict["ida_mats"][:, sweep_index, ...], mats_dict.get("bda_mat", None))
        if self.fuse_geometry and self.frustum.is_cuda:
            # get_geometry + the truncating index in one kernel (.int() truncates toward zero, as in the reference :311-313)
            geom_xyz = lss_geometry_index(self.frustum, *mats, self.voxel_coord, self.voxel_size)
        else:
            geom = self.get_geometry(*mats)
            geom_xyz = ((geom - (self.voxel_coord - self.voxel_size / 2.0)) / self.voxel_size).int()
        fused_ok = (self.fuse_lift_splat and type(self)._forward_voxel_net is LSSFPN._forward_voxel_net
                    and self.output_channels in (128, 256))
        if fused_ok:
            feature_map = lift_splat(depth, context, geom_xyz, self._voxel_num_host)
        else:
            vol = self._forward_voxel_net(depth.unsqueeze(1) * context.unsqueeze(2))
            vol = vol.reshape(batch, ncam, vol.shape[1], vol.shape[2], vol.shape[3], vol.shape[4])
            vol = vol.permute(0, 1, 3, 4, 5, 2)
            feature_map = voxel_pooling(geom_xyz, vol.contiguous(), self._voxel_num_host).contiguous()
        if is_return_depth:
            return feature_map, depth
        return feature_map

    def forward(self, sweep_imgs, mats_dict, timestamps=None, is_return_depth=False):
        """sweep_imgs (B, num_sweeps, num_cams, 3, H, W) -> BEV feature map (B, C*num_sweeps, ny, nx)."""
        num_sweeps = sweep_imgs.shape[1]
        key = self._forward_single_sweep(0, sweep_imgs[:, 0:1, ...], mats_dict, is_return_depth=is_return_depth)
        if num_sweeps == 1:
            return key
        maps = [key[0] if is_return_depth else key]
        for s in range(1, num_sweeps):
            with torch.no_grad():
                maps.append(self._forward_single_sweep(s, sweep_imgs[:, s: s + 1, ...], mats_dict,
                                                       is_return_depth=False))
        if is_return_depth:
            return torch.cat(maps, 1), key[1]
        return torch.cat(maps, 1)
