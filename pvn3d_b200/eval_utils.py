"""cal_frame_poses / cal_frame_poses_lm / best_fit_transform -- drop-ins for the reference's
post-network pose recovery (pvn3d/lib/utils/pvn3d_eval_utils.py:37-110,156-201 and
pvn3d/lib/utils/basic_utils.py:47-80), executed by csrc/poses.cu + csrc/meanshift.cu.

Reference-compatible entry points (same arguments, same return types):
    cal_frame_poses(pcld, mask, ctr_of, pred_kp_of, use_ctr, n_cls, use_ctr_clus_flter)
        -> (pred_cls_ids: np.ndarray, pred_pose_lst: list[np.ndarray(3,4) float64])
    cal_frame_poses_lm(pcld, mask, ctr_of, pred_kp_of, use_ctr, n_cls, use_ctr_clus_flter, obj_id)
        -> list[np.ndarray(3,4) float64]
    best_fit_transform(A, B) -> np.ndarray(3,4) float64
Batched, sync-free form used by the frame pipeline and bench.py:
    FramePoseSolver(...).solve(pcld[B,N,3], mask[B,N], ctr_of[B,N,3], kp_of[B,K,N,3])
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import numpy as np
import torch

from . import _lib, fixtures
from ._lib import check, ms_flags, ptr

RADIUS = 0.08  # clustering bandwidth hard-coded by the reference (pvn3d_eval_utils.py:44,163)


class FramePoseSolver:
    """Device-resident solver for a fixed (B, N, K, n_cls) problem shape; owns its workspace."""

    def __init__(self, batch: int, n_pts: int, n_kps: int, n_cls: int, mesh_kps: np.ndarray,
                 cls_radius: Optional[np.ndarray], use_ctr_clus_flter: bool, device="cuda",
                 bandwidth: float = RADIUS, max_iter: int = 300, early_exit: bool = False, mode: str = None):
        self.lib = _lib.load()
        self.dev = torch.device(device)
        if self.dev.type != "cuda":
            raise RuntimeError("FramePoseSolver: CUDA only -- no CPU fallback")
        self.b, self.n, self.k, self.n_cls = int(batch), int(n_pts), int(n_kps), int(n_cls)
        mesh_kps = np.ascontiguousarray(mesh_kps, dtype=np.float32)
        assert mesh_kps.shape == (self.n_cls, self.k + 1, 3), mesh_kps.shape
        self.mesh_kps = torch.from_numpy(mesh_kps).to(self.dev)
        self.use_filter = bool(use_ctr_clus_flter)
        self.cls_radius = None
        if self.use_filter:
            assert cls_radius is not None and len(cls_radius) == self.n_cls
            self.cls_radius = torch.from_numpy(np.ascontiguousarray(cls_radius, dtype=np.float32)).to(self.dev)
        self.bandwidth, self.max_iter = float(bandwidth), int(max_iter)
        self.flags = ms_flags(mode, early_exit)       # default: "certified" (see MeanShiftTorch)
        self.ws_bytes = int(self.lib.pvn3d_frame_poses_workspace_bytes(self.b, self.n, self.k, self.n_cls, self.max_iter))
        self._ws = torch.empty((self.ws_bytes + 256,), dtype=torch.uint8, device=self.dev)
        self._ws_ptr = (self._ws.data_ptr() + 255) // 256 * 256
        self.poses = torch.empty((self.b, self.n_cls, 3, 4), dtype=torch.float32, device=self.dev)
        self.present = torch.empty((self.b, self.n_cls), dtype=torch.uint8, device=self.dev)
        self.cls_kps = torch.empty((self.b, self.n_cls, self.k + 1, 3), dtype=torch.float32, device=self.dev)
        self.new_mask = torch.empty((self.b, self.n), dtype=torch.int32, device=self.dev)
        self._ms_off = int(self.lib.pvn3d_frame_poses_ms_workspace_offset(self.b, self.n, self.k, self.n_cls, self.max_iter))

    def certified_fits(self) -> int:
        """diagnostics (synchronises): how many of the last solve()'s centre + keypoint fits the witness
        kernel closed without sweeping all seeds (mode "certified")"""
        off = self._ws_ptr - self._ws.data_ptr() + self._ms_off
        return int(self._ws[off:off + 64].view(torch.int32)[_lib.PVN3D_MS_STAT_CERTIFIED].item())

    def solve(self, pcld: torch.Tensor, mask: torch.Tensor, ctr_of: torch.Tensor, kp_of: torch.Tensor):
        """All inputs on device, contiguous: pcld [B,N,3] f32, mask [B,N] i32, ctr_of [B,N,3] f32,
        kp_of [B,K,N,3] f32.  Returns (poses [B,n_cls,3,4], present [B,n_cls] u8, cls_kps, new_mask)
        -- views of solver-owned buffers, valid until the next solve() on the same stream."""
        b = pcld.size(0)
        assert b <= self.b and pcld.shape[1:] == (self.n, 3) and kp_of.shape[1:] == (self.k, self.n, 3)
        for t, dt in ((pcld, torch.float32), (mask, torch.int32), (ctr_of, torch.float32), (kp_of, torch.float32)):
            assert t.is_cuda and t.is_contiguous() and t.dtype == dt
        with torch.cuda.device(self.dev):
            rc = self.lib.pvn3d_frame_poses_batch(
                ptr(pcld), ptr(mask), ptr(ctr_of), ptr(kp_of), b, self.n, self.k, self.n_cls,
                ptr(self.mesh_kps), ptr(self.cls_radius), 1 if self.use_filter else 0,
                self.bandwidth, self.max_iter, self.flags, ptr(self.poses), ptr(self.present),
                ptr(self.cls_kps), ptr(self.new_mask), self._ws_ptr, self.ws_bytes,
                torch.cuda.current_stream(self.dev).cuda_stream)
        check(rc, "pvn3d_frame_poses_batch")
        return self.poses[:b], self.present[:b], self.cls_kps[:b], self.new_mask[:b]


_solver_cache = {}


def _solver(kind, n_pts, n_kps, n_cls, use_filter, dev, obj_id=None, early_exit=False, mode=None):
    key = (kind, n_pts, n_kps, n_cls, use_filter, str(dev), obj_id, ms_flags(mode, early_exit))
    s = _solver_cache.get(key)
    if s is None:
        if kind == "ycb":
            mesh = fixtures.mesh_kps_table_ycb()[:n_cls]
            rad = fixtures.radius_thresholds_ycb()[:n_cls] if use_filter else None
        else:
            mesh = np.zeros((n_cls, n_kps + 1, 3), np.float32)
            mesh[1] = fixtures.mesh_kps_table_lm(obj_id)[1]
            rad = np.full((n_cls,), np.inf, np.float32) if use_filter else None
        s = FramePoseSolver(1, n_pts, n_kps, n_cls, mesh, rad, use_filter, device=dev, early_exit=early_exit, mode=mode)
        _solver_cache[key] = s
    return s


def _prep(pcld, mask, ctr_of, pred_kp_of):
    if not pcld.is_cuda:
        raise RuntimeError("cal_frame_poses (pvn3d_b200): CUDA tensors only -- no CPU fallback")
    n_kps, n_pts, _ = pred_kp_of.size()
    return (pcld.reshape(1, n_pts, 3).contiguous().float(), ctr_of[0].reshape(1, n_pts, 3).contiguous().float(),
            pred_kp_of.reshape(1, n_kps, n_pts, 3).contiguous().float(), n_kps, n_pts)


def cal_frame_poses(pcld, mask, ctr_of, pred_kp_of, use_ctr, n_cls, use_ctr_clus_flter,
                    early_exit: bool = False, mode: str = None) -> Tuple[np.ndarray, List[np.ndarray]]:
    """pvn3d_eval_utils.py:37-110.  pcld [N,3], mask [N] (int64 class ids), ctr_of [1,N,3],
    pred_kp_of [K,N,3]; one device->host read at the end instead of >= 3 per class."""
    if not use_ctr:
        raise NotImplementedError("use_ctr=False is never exercised by the reference callers")
    p, c, kp, n_kps, n_pts = _prep(pcld, mask, ctr_of, pred_kp_of)
    m32 = mask.reshape(1, n_pts).to(torch.int32).contiguous()
    s = _solver("ycb", n_pts, n_kps, int(n_cls), bool(use_ctr_clus_flter), pcld.device, early_exit=early_exit,
                mode=mode)
    poses, present, _, _ = s.solve(p, m32, c, kp)
    poses_h = poses[0].double().cpu().numpy()
    present_h = present[0].cpu().numpy()
    pred_cls_ids = np.nonzero(present_h)[0].astype(np.int64)  # == np.unique(mask[mask>0]) (:50)
    return pred_cls_ids, [poses_h[c_id].copy() for c_id in pred_cls_ids]


def cal_frame_poses_lm(pcld, mask, ctr_of, pred_kp_of, use_ctr, n_cls, use_ctr_clus_flter, obj_id,
                       early_exit: bool = False, mode: str = None) -> List[np.ndarray]:
    """pvn3d_eval_utils.py:156-201: single class id 1, fixtures of LineMOD object `obj_id`."""
    if not use_ctr:
        raise NotImplementedError("use_ctr=False is never exercised by the reference callers")
    p, c, kp, n_kps, n_pts = _prep(pcld, mask, ctr_of, pred_kp_of)
    m32 = (mask.reshape(1, n_pts) == 1).to(torch.int32).contiguous()  # cls_msk = mask == 1 (:169-170)
    # the lm variant has no centre-cluster relabel pass; the flag only gates the inlier filter of
    # the keypoint votes (:182-185), which the solver's flag also controls -> give it an infinite
    # radius so the relabel pass is the identity.
    s = _solver("lm", n_pts, n_kps, max(int(n_cls), 2), bool(use_ctr_clus_flter), pcld.device, obj_id=obj_id,
                early_exit=early_exit, mode=mode)
    poses, _, _, _ = s.solve(p, m32, c, kp)
    return [poses[0, 1].double().cpu().numpy()]


def best_fit_transform(A, B) -> np.ndarray:
    """basic_utils.py:47-80: least-squares rigid transform mapping A [P,3] onto B [P,3] -> 3x4 float64.
    Accepts numpy arrays or tensors; the SVD runs on the GPU (float64)."""
    lib = _lib.load()
    a = torch.as_tensor(np.asarray(A, dtype=np.float32) if not torch.is_tensor(A) else A.float())
    b = torch.as_tensor(np.asarray(B, dtype=np.float32) if not torch.is_tensor(B) else B.float())
    assert a.shape == b.shape and a.dim() == 2 and a.size(1) == 3
    a = a.cuda().contiguous()
    b = b.cuda().contiguous()
    rt = torch.empty((1, 3, 4), dtype=torch.float32, device=a.device)
    with torch.cuda.device(a.device):
        rc = lib.pvn3d_best_fit_transform_batch(ptr(a), ptr(b), 0, 1, int(a.size(0)), ptr(rt),
                                                torch.cuda.current_stream(a.device).cuda_stream)
    check(rc, "pvn3d_best_fit_transform_batch")
    return rt[0].double().cpu().numpy()


# --------------------------------------------------------------------------------------------------
# callers either side of the path: segmentation argmax, ADD / ADD-S
# --------------------------------------------------------------------------------------------------
def seg_argmax(logits: torch.Tensor) -> torch.Tensor:
    """`_, classes_rgbd = torch.max(pred_rgbd_seg, -1)` (demo.py:108) as one kernel: logits [..., n_cls] f32
    -> int32 class ids [...] -- the `mask` FramePoseSolver.solve takes (first maximal index)."""
    if not logits.is_cuda:
        raise RuntimeError("seg_argmax (pvn3d_b200): CUDA tensors only -- no CPU fallback")
    lib = _lib.load()
    x = logits.contiguous().float()
    n_cls = x.size(-1)
    out = torch.empty(x.shape[:-1], dtype=torch.int32, device=x.device)
    with torch.cuda.device(x.device):
        rc = lib.pvn3d_seg_argmax(ptr(x), x.numel() // n_cls, n_cls, ptr(out), torch.cuda.current_stream(x.device).cuda_stream)
    check(rc, "pvn3d_seg_argmax")
    return out


def pose_add_adds(pred_RT: torch.Tensor, gt_RT: torch.Tensor, p3ds: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """ADD and ADD-S of F pose pairs over one mesh in two launches: pred_RT, gt_RT [F,3,4] (or [3,4]),
    p3ds [P,3] -> (add [F], adds [F]).  basic_utils.py:617-635."""
    if not p3ds.is_cuda:
        raise RuntimeError("pose_add_adds (pvn3d_b200): CUDA tensors only -- no CPU fallback")
    lib = _lib.load()
    dev = p3ds.device
    pr = pred_RT.reshape(-1, 3, 4).to(dev, torch.float32).contiguous()
    gt = gt_RT.reshape(-1, 3, 4).to(dev, torch.float32).contiguous()
    assert pr.shape == gt.shape
    pts = p3ds.to(torch.float32).contiguous()
    f, p = pr.size(0), pts.size(0)
    add = torch.empty((f,), dtype=torch.float32, device=dev)
    adds = torch.empty((f,), dtype=torch.float32, device=dev)
    ws_bytes = int(lib.pvn3d_pose_add_adds_workspace_bytes(f, p))
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        rc = lib.pvn3d_pose_add_adds(ptr(pr), ptr(gt), f, ptr(pts), p, ptr(add), ptr(adds), ptr(ws), ws_bytes,
                                     torch.cuda.current_stream(dev).cuda_stream)
    check(rc, "pvn3d_pose_add_adds")
    ws.record_stream(torch.cuda.current_stream(dev))
    return add, adds


def cal_add_cuda(pred_RT, gt_RT, p3ds) -> torch.Tensor:
    """Basic_Utils.cal_add_cuda (basic_utils.py:617-623): 0-dim tensor"""
    return pose_add_adds(pred_RT, gt_RT, p3ds)[0][0]


def cal_adds_cuda(pred_RT, gt_RT, p3ds) -> torch.Tensor:
    """Basic_Utils.cal_adds_cuda (basic_utils.py:625-635): 0-dim tensor"""
    return pose_add_adds(pred_RT, gt_RT, p3ds)[1][0]
