"""Host-side mirror of the reference's PointNet++ Python surface, on top of pvn3d_b200._ext.

Mirrors (same names, argument meaning, tensor layouts and state_dict keys):
  * pvn3d/lib/pointnet2_utils/pointnet2_utils.py   furthest_point_sample, gather_operation, three_nn,
        three_interpolate, grouping_operation, ball_query, QueryAndGroup, GroupAll
  * pvn3d/lib/pointnet2_utils/pointnet2_modules.py PointnetSAModuleMSG, PointnetSAModule,
        PointnetFPModule
  * pvn3d/lib/utils/etw_pytorch_utils/pytorch_utils.py:25-50   SharedMLP (1x1 Conv2d + BN2d + ReLU)
  * pvn3d/lib/pvn3d.py:46-154                      Pointnet2MSG (4 SA-MSG + 4 FP levels)
so a reference checkpoint's `pointnet2.*` entries load unchanged and the parity tests read like the
reference's call sites.  The reference modules themselves also run unmodified on this package's
`_ext` (see compat.install()); this mirror exists because /root/reference is not present on the GPU
box and because its forward can take the fused kernels (query_and_group, three_nn_interpolate).
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.autograd import Function

from . import _ext

# --------------------------------------------------------------------------------------------------
# autograd wrappers (pointnet2_utils.py:37-273)
# --------------------------------------------------------------------------------------------------


class FurthestPointSampling(Function):
    @staticmethod
    def forward(ctx, xyz, npoint):
        idx = _ext.furthest_point_sampling(xyz, npoint)
        ctx.mark_non_differentiable(idx)
        return idx

    @staticmethod
    def backward(ctx, a=None):
        return None, None


furthest_point_sample = FurthestPointSampling.apply


class GatherOperation(Function):
    @staticmethod
    def forward(ctx, features, idx):
        ctx.for_backwards = (idx, features.size(1), features.size(2))
        return _ext.gather_points(features, idx)

    @staticmethod
    def backward(ctx, grad_out):
        idx, _, n = ctx.for_backwards
        return _ext.gather_points_grad(grad_out.contiguous(), idx, n), None


gather_operation = GatherOperation.apply


class ThreeNN(Function):
    @staticmethod
    def forward(ctx, unknown, known):
        dist2, idx = _ext.three_nn(unknown, known)
        dist = torch.sqrt(dist2)  # the op returns SQUARED distances (pointnet2_utils.py:124-126)
        ctx.mark_non_differentiable(dist, idx)
        return dist, idx

    @staticmethod
    def backward(ctx, a=None, b=None):
        return None, None


three_nn = ThreeNN.apply


class ThreeInterpolate(Function):
    @staticmethod
    def forward(ctx, features, idx, weight):
        ctx.three_interpolate_for_backward = (idx, weight, features.size(2))
        return _ext.three_interpolate(features, idx, weight)

    @staticmethod
    def backward(ctx, grad_out):
        idx, weight, m = ctx.three_interpolate_for_backward
        return _ext.three_interpolate_grad(grad_out.contiguous(), idx, weight, m), None, None


three_interpolate = ThreeInterpolate.apply


class GroupingOperation(Function):
    @staticmethod
    def forward(ctx, features, idx):
        ctx.for_backwards = (idx, features.size(2))
        return _ext.group_points(features, idx)

    @staticmethod
    def backward(ctx, grad_out):
        idx, n = ctx.for_backwards
        return _ext.group_points_grad(grad_out.contiguous(), idx, n), None


grouping_operation = GroupingOperation.apply


class BallQuery(Function):
    @staticmethod
    def forward(ctx, radius, nsample, xyz, new_xyz):
        idx = _ext.ball_query(new_xyz, xyz, radius, nsample)  # centres first (ball_query.h:4-5)
        ctx.mark_non_differentiable(idx)
        return idx

    @staticmethod
    def backward(ctx, a=None):
        return None, None, None, None


ball_query = BallQuery.apply


class QueryAndGroup(nn.Module):
    """pointnet2_utils.py:276-330.  forward(xyz [B,N,3], new_xyz [B,M,3], features [B,C,N])
    -> [B, 3+C, M, nsample].  Without autograd the five reference launches collapse into the fused
    query_and_group kernel (bit-identical output)."""

    def __init__(self, radius, nsample, use_xyz=True):
        super().__init__()
        self.radius, self.nsample, self.use_xyz = radius, nsample, use_xyz

    def forward(self, xyz, new_xyz, features=None):
        needs_grad = torch.is_grad_enabled() and features is not None and features.requires_grad
        if self.use_xyz and not needs_grad and self.nsample <= 256:
            feat_pm = _ext.transpose_cn_to_nc(features) if features is not None else None
            out, _ = _ext.query_and_group(xyz, new_xyz, feat_pm, self.radius, self.nsample, want_idx=False)
            return out
        idx = ball_query(self.radius, self.nsample, xyz, new_xyz)
        grouped_xyz = grouping_operation(xyz.transpose(1, 2).contiguous(), idx)
        grouped_xyz = grouped_xyz - new_xyz.transpose(1, 2).unsqueeze(-1)
        if features is None:
            assert self.use_xyz, "Cannot have not features and not use xyz as a feature!"
            return grouped_xyz
        grouped_features = grouping_operation(features, idx)
        return torch.cat([grouped_xyz, grouped_features], dim=1) if self.use_xyz else grouped_features


class GroupAll(nn.Module):
    """pointnet2_utils.py:333-376"""

    def __init__(self, use_xyz=True):
        super().__init__()
        self.use_xyz = use_xyz

    def forward(self, xyz, new_xyz, features=None):
        grouped_xyz = xyz.transpose(1, 2).unsqueeze(2)
        if features is None:
            return grouped_xyz
        grouped_features = features.unsqueeze(2)
        return torch.cat([grouped_xyz, grouped_features], dim=1) if self.use_xyz else grouped_features


# --------------------------------------------------------------------------------------------------
# SharedMLP (pytorch_utils.py:25-50,80-134): state_dict keys layer{i}.conv.weight,
# layer{i}.normlayer.bn.{weight,bias,running_mean,running_var,num_batches_tracked}
# --------------------------------------------------------------------------------------------------


def _conv_bn_relu(c_in: int, c_out: int, bn: bool) -> nn.Sequential:
    blk = nn.Sequential()
    conv = nn.Conv2d(c_in, c_out, kernel_size=(1, 1), bias=not bn)  # bias = bias and (not bn) (:98)
    nn.init.kaiming_normal_(conv.weight)
    if conv.bias is not None:
        nn.init.constant_(conv.bias, 0)
    blk.add_module("conv", conv)
    if bn:
        norm = nn.Sequential()
        norm.add_module("bn", nn.BatchNorm2d(c_out))
        nn.init.constant_(norm.bn.weight, 1.0)
        nn.init.constant_(norm.bn.bias, 0)
        blk.add_module("normlayer", norm)
    blk.add_module("activation", nn.ReLU(inplace=True))
    return blk


class SharedMLP(nn.Sequential):
    def __init__(self, args: List[int], bn: bool = False):
        super().__init__()
        for i in range(len(args) - 1):
            self.add_module("layer{}".format(i), _conv_bn_relu(args[i], args[i + 1], bn))


# --------------------------------------------------------------------------------------------------
# set abstraction / feature propagation (pointnet2_modules.py:20-206)
# --------------------------------------------------------------------------------------------------


class _PointnetSAModuleBase(nn.Module):
    def __init__(self):
        super().__init__()
        self.npoint = None
        self.groupers = None
        self.mlps = None

    def forward(self, xyz: torch.Tensor, features: Optional[torch.Tensor] = None):
        """xyz [B,N,3], features [B,C,N] -> (new_xyz [B,npoint,3], new_features [B,sum C_out,npoint])"""
        new_xyz = None
        if self.npoint is not None:
            idx = furthest_point_sample(xyz, self.npoint)
            new_xyz = gather_operation(xyz.transpose(1, 2).contiguous(), idx).transpose(1, 2).contiguous()
        outs = []
        for grouper, mlp in zip(self.groupers, self.mlps):
            g = grouper(xyz, new_xyz, features)              # [B, C, npoint, nsample]
            g = mlp(g)                                       # [B, mlp[-1], npoint, nsample]
            g = F.max_pool2d(g, kernel_size=[1, g.size(3)])  # [B, mlp[-1], npoint, 1]
            outs.append(g.squeeze(-1))
        return new_xyz, torch.cat(outs, dim=1)


class PointnetSAModuleMSG(_PointnetSAModuleBase):
    def __init__(self, npoint, radii, nsamples, mlps, bn=True, use_xyz=True):
        super().__init__()
        assert len(radii) == len(nsamples) == len(mlps)
        self.npoint = npoint
        self.groupers = nn.ModuleList()
        self.mlps = nn.ModuleList()
        for radius, nsample, spec in zip(radii, nsamples, mlps):
            self.groupers.append(QueryAndGroup(radius, nsample, use_xyz=use_xyz) if npoint is not None
                                 else GroupAll(use_xyz))
            spec = list(spec)  # the reference mutates the caller's list (pointnet2_modules.py:108-110)
            if use_xyz:
                spec[0] += 3
            self.mlps.append(SharedMLP(spec, bn=bn))


class PointnetSAModule(PointnetSAModuleMSG):
    def __init__(self, mlp, npoint=None, radius=None, nsample=None, bn=True, use_xyz=True):
        super().__init__(mlps=[mlp], npoint=npoint, radii=[radius], nsamples=[nsample], bn=bn, use_xyz=use_xyz)


class PointnetFPModule(nn.Module):
    def __init__(self, mlp, bn=True):
        super().__init__()
        self.mlp = SharedMLP(mlp, bn=bn)

    def forward(self, unknown, known, unknow_feats, known_feats):
        """unknown [B,n,3], known [B,m,3], unknow_feats [B,C1,n], known_feats [B,C2,m] -> [B,mlp[-1],n]"""
        if known is not None:
            dist, idx = three_nn(unknown, known)
            dist_recip = 1.0 / (dist + 1e-8)
            norm = torch.sum(dist_recip, dim=2, keepdim=True)
            weight = dist_recip / norm
            interpolated = three_interpolate(known_feats, idx, weight)
        else:
            interpolated = known_feats.expand(*(list(known_feats.size()[0:2]) + [unknown.size(1)]))
        new_features = torch.cat([interpolated, unknow_feats], dim=1) if unknow_feats is not None else interpolated
        return self.mlp(new_features.unsqueeze(-1)).squeeze(-1)


# --------------------------------------------------------------------------------------------------
# Pointnet2MSG (pvn3d.py:46-154)
# --------------------------------------------------------------------------------------------------

#: (npoint, radii, nsamples, mlps-without-xyz) per SA level -- the layer spec of pvn3d.py:65-111
SA_SPEC = (
    (2048, (0.0175, 0.025), (16, 32), ((16, 16, 32), (32, 32, 64))),
    (1024, (0.025, 0.05), (16, 32), ((64, 64, 128), (64, 96, 128))),
    (512, (0.05, 0.1), (16, 32), ((128, 196, 256), (128, 196, 256))),
    (128, (0.1, 0.2), (16, 32), ((256, 256, 512), (256, 384, 512))),
)


class Pointnet2MSG(nn.Module):
    """PointNet++ multi-scale-grouping encoder/decoder: pointcloud [B,N,3+C] -> features [B,128,N]."""

    def __init__(self, input_channels=6, use_xyz=True):
        super().__init__()
        self.SA_modules = nn.ModuleList()
        c_in = input_channels
        skip = [input_channels]
        for npoint, radii, nsamples, mlps in SA_SPEC:
            self.SA_modules.append(PointnetSAModuleMSG(
                npoint=npoint, radii=list(radii), nsamples=list(nsamples),
                mlps=[[c_in] + list(m) for m in mlps], use_xyz=use_xyz))
            c_in = sum(m[-1] for m in mlps)
            skip.append(c_in)
        # pvn3d.py:113-118
        self.FP_modules = nn.ModuleList([
            PointnetFPModule(mlp=[256 + skip[0], 128, 128]),
            PointnetFPModule(mlp=[512 + skip[1], 256, 256]),
            PointnetFPModule(mlp=[512 + skip[2], 512, 512]),
            PointnetFPModule(mlp=[skip[4] + skip[3], 512, 512]),
        ])

    @staticmethod
    def _break_up_pc(pc):
        xyz = pc[..., 0:3].contiguous()
        features = pc[..., 3:].transpose(1, 2).contiguous() if pc.size(-1) > 3 else None
        return xyz, features

    def forward(self, pointcloud: torch.Tensor) -> torch.Tensor:
        xyz, features = self._break_up_pc(pointcloud)
        l_xyz, l_features = [xyz], [features]
        for sa in self.SA_modules:
            nx, nf = sa(l_xyz[-1], l_features[-1])
            l_xyz.append(nx)
            l_features.append(nf)
        for i in range(-1, -(len(self.FP_modules) + 1), -1):
            l_features[i - 1] = self.FP_modules[i](l_xyz[i - 1], l_xyz[i], l_features[i - 1], l_features[i])
        return l_features[0]
