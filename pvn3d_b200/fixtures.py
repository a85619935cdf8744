"""Object-frame keypoint / radius constants of the reference, packed by tools/pack_fixtures.py.

Host-side mirror of Basic_Utils.get_kps / get_ctr (reference pvn3d/lib/utils/basic_utils.py:541-595)
and Config.ycb_r_lst / ycb_cls_lst / lm_obj_dict (pvn3d/common.py:80-81,92-106).
"""
from __future__ import annotations

import functools
import os

import numpy as np

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "fixtures")

N_KEYPOINTS = 8          # common.py:44
N_SAMPLE_POINTS = 12288  # common.py:43
YCB_N_CLASSES = 22       # 21 objects + background (common.py:50-51)


@functools.lru_cache(maxsize=None)
def _ycb():
    z = np.load(os.path.join(_DIR, "ycb.npz"))
    return {k: z[k] for k in z.files}


@functools.lru_cache(maxsize=None)
def _lm():
    z = np.load(os.path.join(_DIR, "linemod.npz"))
    return {k: z[k] for k in z.files}


def ycb_cls_lst():
    return [str(c) for c in _ycb()["classes"]]


def ycb_r_lst():
    """list of float64 radii, as list(np.loadtxt(radius.txt)) (common.py:80)"""
    return [float(r) for r in _ycb()["radius"]]


def lm_obj_dict():
    d = _lm()
    return {str(n): int(i) for n, i in zip(d["names"], d["obj_ids"])}


def _resolve(cls, ds_type):
    if ds_type == "ycb":
        names = ycb_cls_lst()
        if isinstance(cls, (int, np.integer)):
            return int(cls) - 1            # class ids are 1-based (basic_utils.py:548-549)
        return names.index(cls)
    d = _lm()
    if isinstance(cls, (int, np.integer)):
        return int(np.where(d["obj_ids"] == int(cls))[0][0])
    return [str(n) for n in d["names"]].index(cls)


def get_kps(cls, kp_type="farthest", ds_type="ycb"):
    """8 object-frame FPS keypoints, float32 [8,3] (basic_utils.py:541-571)."""
    if kp_type != "farthest":
        raise ValueError("only the 'farthest' (8-keypoint) set is packed")
    src = _ycb() if ds_type == "ycb" else _lm()
    return src["farthest"][_resolve(cls, ds_type)].copy()


def get_ctr(cls, ds_type="ycb"):
    """object centre = mean of the 8 bbox corners, float32 [3] (basic_utils.py:573-595)."""
    src = _ycb() if ds_type == "ycb" else _lm()
    return src["corners"][_resolve(cls, ds_type)].mean(0)


def mesh_kps_table_ycb() -> np.ndarray:
    """[22, 9, 3] float32: row c = 8 keypoints + centre (LAST) of class id c; row 0 unused."""
    t = np.zeros((YCB_N_CLASSES, N_KEYPOINTS + 1, 3), np.float32)
    for c in range(1, YCB_N_CLASSES):
        t[c, :N_KEYPOINTS] = get_kps(c, ds_type="ycb")
        t[c, N_KEYPOINTS] = get_ctr(c, ds_type="ycb")
    return t


def mesh_kps_table_lm(obj_id) -> np.ndarray:
    """[2, 9, 3] float32: row 1 = fixtures of LineMOD object `obj_id` (cal_frame_poses_lm uses class id 1)."""
    t = np.zeros((2, N_KEYPOINTS + 1, 3), np.float32)
    t[1, :N_KEYPOINTS] = get_kps(obj_id, ds_type="linemod")
    t[1, N_KEYPOINTS] = get_ctr(obj_id, ds_type="linemod")
    return t


def radius_thresholds_ycb() -> np.ndarray:
    """[22] float32: float32(float64(r) * 0.8) per class id -- the value the reference's
    `min_dis < config.ycb_r_lst[cls_id-1] * 0.8` compares a float32 tensor against
    (pvn3d_eval_utils.py:69)."""
    t = np.zeros((YCB_N_CLASSES,), np.float32)
    for c, r in enumerate(ycb_r_lst(), start=1):
        t[c] = np.float32(r * 0.8)
    return t
