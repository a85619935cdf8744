"""TEST INFRASTRUCTURE -- CPU restatement of cal_frame_poses / cal_frame_poses_lm.

Follows pvn3d/lib/utils/pvn3d_eval_utils.py:37-110 and :156-201 step by step on CPU float32
tensors (the reference hard-codes .cuda() at :46,58,64,103 and cannot run without a GPU).
Fixtures (mesh keypoints, radii) come in as arguments so the oracle does not depend on the product
package.  Returns the reference's outputs plus the intermediates the parity tests compare
(relabelled mask, voted keypoints).
Pinned against the real functions, executed on CPU through a `.cuda()` no-op patch, by
tests/golden/make_golden_cpu.py.
"""
from __future__ import annotations

import numpy as np
import torch

from .meanshift_oracle import MeanShiftOracle, best_fit_transform

RADIUS = 0.08  # :44,163


def cal_frame_poses(pcld, mask, ctr_of, pred_kp_of, use_ctr, n_cls, use_ctr_clus_flter,
                    mesh_kps_of, mesh_ctr_of, ycb_r_lst):
    """mesh_kps_of(cls_id) -> [8,3] f32, mesh_ctr_of(cls_id) -> [3] f32, ycb_r_lst: list of float64."""
    n_kps, n_pts, _ = pred_kp_of.size()
    pred_ctr = pcld - ctr_of[0]                                              # :41
    pred_kp = pcld.view(1, n_pts, 3).repeat(n_kps, 1, 1) - pred_kp_of       # :42
    cls_kps = torch.zeros(n_cls, n_kps + (1 if use_ctr else 0), 3)          # :45-48
    pred_cls_ids = np.unique(mask[mask > 0].contiguous().numpy())           # :50
    if use_ctr_clus_flter:                                                   # :51-73
        ctrs = []
        for cls_id in pred_cls_ids:
            ctr, _ = MeanShiftOracle(bandwidth=RADIUS).fit(pred_ctr[mask == cls_id, :])
            ctrs.append(ctr.numpy())
        ctrs = torch.from_numpy(np.array(ctrs).astype(np.float32))
        n_ctrs = ctrs.size(0)
        ctr_dis = torch.norm(pred_ctr.view(n_pts, 1, 3).repeat(1, n_ctrs, 1)
                             - ctrs.view(1, n_ctrs, 3).repeat(n_pts, 1, 1), dim=2)
        min_dis, min_idx = torch.min(ctr_dis, dim=1)
        msk_closest_ctr = torch.LongTensor(pred_cls_ids)[min_idx]
        new_msk = mask.clone()
        for cls_id in pred_cls_ids:
            if cls_id == 0:
                break
            min_msk = min_dis < ycb_r_lst[cls_id - 1] * 0.8
            update_msk = (mask > 0) & (msk_closest_ctr == cls_id) & min_msk
            new_msk[update_msk] = msk_closest_ctr[update_msk]
        mask = new_msk
    pred_pose_lst = []
    for cls_id in pred_cls_ids:                                              # :75-108
        if cls_id == 0:
            break
        cls_msk = mask == cls_id
        if cls_msk.sum() < 1:
            pred_pose_lst.append(np.identity(4)[:3, :])
            continue
        cls_voted_kps = pred_kp[:, cls_msk, :]
        ms = MeanShiftOracle(bandwidth=RADIUS)
        ctr, ctr_labels = ms.fit(pred_ctr[cls_msk, :])
        if ctr_labels.sum() < 1:
            ctr_labels[0] = 1
        if use_ctr:
            cls_kps[cls_id, n_kps, :] = ctr
        in_pred_kp = cls_voted_kps[:, ctr_labels, :] if use_ctr_clus_flter else cls_voted_kps
        for ikp, kps3d in enumerate(in_pred_kp):
            cls_kps[cls_id, ikp, :], _ = ms.fit(kps3d)
        mesh_kps = mesh_kps_of(int(cls_id))
        if use_ctr:
            mesh_kps = np.concatenate((mesh_kps, mesh_ctr_of(int(cls_id)).reshape(1, 3)), axis=0)
        mesh_kps = mesh_kps.astype(np.float32)
        pred_pose_lst.append(best_fit_transform(mesh_kps, cls_kps[cls_id].squeeze().contiguous().numpy()))
    return pred_cls_ids, pred_pose_lst, mask, cls_kps


def cal_frame_poses_lm(pcld, mask, ctr_of, pred_kp_of, use_ctr, n_cls, use_ctr_clus_flter,
                       mesh_kps, mesh_ctr):
    """:156-201 -- single class id 1; mesh_kps [8,3], mesh_ctr [3] of the LineMOD object."""
    n_kps, n_pts, _ = pred_kp_of.size()
    pred_ctr = pcld - ctr_of[0]
    pred_kp = pcld.view(1, n_pts, 3).repeat(n_kps, 1, 1) - pred_kp_of
    cls_kps = torch.zeros(n_cls, n_kps + 1, 3)
    cls_id = 1
    cls_msk = mask == cls_id
    if cls_msk.sum() < 1:
        return [np.identity(4)[:3, :]], cls_kps
    cls_voted_kps = pred_kp[:, cls_msk, :]
    ms = MeanShiftOracle(bandwidth=RADIUS)
    ctr, ctr_labels = ms.fit(pred_ctr[cls_msk, :])
    if ctr_labels.sum() < 1:
        ctr_labels[0] = 1
    cls_kps[cls_id, n_kps, :] = ctr
    in_pred_kp = cls_voted_kps[:, ctr_labels, :] if use_ctr_clus_flter else cls_voted_kps
    for ikp, kps3d in enumerate(in_pred_kp):
        cls_kps[cls_id, ikp, :], _ = ms.fit(kps3d)
    mk = np.concatenate((mesh_kps, mesh_ctr.reshape(1, 3)), axis=0).astype(np.float32)
    return [best_fit_transform(mk, cls_kps[cls_id].squeeze().contiguous().numpy())], cls_kps
