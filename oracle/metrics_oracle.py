"""TEST INFRASTRUCTURE -- CPU restatement of the pose-error metrics of the reference's evaluation loop.

Follows pvn3d/lib/utils/basic_utils.py:617-635 (Basic_Utils.cal_add_cuda / cal_adds_cuda) op for op on
CPU float32 tensors: `torch.mm(p3ds, RT[:, :3].T) + RT[:, 3]` for both poses, then
  ADD   = mean_k |pred_k - gt_k|                                   (:622-623)
  ADD-S = mean_k min_j |pred_j - gt_k|  via the [N,N,3] repeat      (:629-635)
Pinned against the reference methods themselves (imported from /root/reference) by
tests/golden/make_golden_cpu.py: identical bits on every recorded case (tests/golden/metrics.npz).
"""
from __future__ import annotations

import torch


def cal_add(pred_RT: torch.Tensor, gt_RT: torch.Tensor, p3ds: torch.Tensor) -> torch.Tensor:
    pred_p3ds = torch.mm(p3ds, pred_RT[:, :3].transpose(1, 0)) + pred_RT[:, 3]
    gt_p3ds = torch.mm(p3ds, gt_RT[:, :3].transpose(1, 0)) + gt_RT[:, 3]
    dis = torch.norm(pred_p3ds - gt_p3ds, dim=1)
    return torch.mean(dis)


def cal_adds(pred_RT: torch.Tensor, gt_RT: torch.Tensor, p3ds: torch.Tensor) -> torch.Tensor:
    n, _ = p3ds.size()
    pd = torch.mm(p3ds, pred_RT[:, :3].transpose(1, 0)) + pred_RT[:, 3]
    pd = pd.view(1, n, 3).repeat(n, 1, 1)
    gt = torch.mm(p3ds, gt_RT[:, :3].transpose(1, 0)) + gt_RT[:, 3]
    gt = gt.view(n, 1, 3).repeat(1, n, 1)
    dis = torch.norm(pd - gt, dim=2)
    mdis = torch.min(dis, dim=1)[0]
    return torch.mean(mdis)
