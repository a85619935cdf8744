"""cal_frame_poses / cal_frame_poses_lm / best_fit_transform against the reference's recorded
outputs (tests/golden/poses_*.npz, bft_cases.npz -- produced by the reference's own functions).

Bar: predicted class ids and the relabelled instance mask bit-exact; voted keypoints and poses
within 1e-4 relative (rotation: Frobenius norm of dR; translation: |dt| / |t|).
"""
import os

import numpy as np
import pytest
import torch

from pvn3d_b200 import eval_utils, fixtures
from pvn3d_b200.eval_utils import FramePoseSolver

pytestmark = pytest.mark.gpu
REL_TOL = 1e-4


def _pose_close(p, q):
    dr = np.linalg.norm(p[:, :3] - q[:, :3])
    dt = np.linalg.norm(p[:, 3] - q[:, 3]) / max(np.linalg.norm(q[:, 3]), 1e-9)
    return dr <= REL_TOL * np.sqrt(3) and dt <= REL_TOL, (dr, dt)


@pytest.mark.parametrize("case", [0, 1, 2])
def test_cal_frame_poses_ycb(cuda_dev, golden_dir, case):
    z = np.load(os.path.join(golden_dir, "poses_ycb.npz"))
    j = case
    args = [torch.from_numpy(z[f"c{j}_{k}"]).to(cuda_dev) for k in ("pcld", "mask", "ctr_of", "kp_of")]
    ids, poses = eval_utils.cal_frame_poses(*args, True, 22, True)
    assert ids.dtype == np.int64 and np.array_equal(ids, z[f"c{j}_ids"])
    assert len(poses) == len(ids) and all(p.shape == (3, 4) and p.dtype == np.float64 for p in poses)
    for p, q in zip(poses, z[f"c{j}_poses"]):
        ok, err = _pose_close(p, q)
        assert ok, err
    # intermediates through the batched solver: relabelled mask exact, voted keypoints within tolerance
    n = args[0].shape[0]
    s = FramePoseSolver(1, n, 8, 22, fixtures.mesh_kps_table_ycb(), fixtures.radius_thresholds_ycb(), True,
                        device=cuda_dev)
    _, present, cls_kps, new_mask = s.solve(args[0][None].contiguous(), args[1][None].to(torch.int32).contiguous(),
                                            args[2][0][None].contiguous(), args[3][None].contiguous())
    assert np.array_equal(new_mask[0].cpu().numpy(), z[f"c{j}_new_mask"].astype(np.int32)), "instance labels must be bit-exact"
    want = z[f"c{j}_cls_kps"]
    got = cls_kps[0].cpu().numpy()
    for c in ids:
        assert np.linalg.norm(got[c] - want[c]) <= REL_TOL * np.linalg.norm(want[c])
    assert np.array_equal(np.nonzero(present[0].cpu().numpy())[0], ids)


def test_cal_frame_poses_lm(cuda_dev, golden_dir):
    z = np.load(os.path.join(golden_dir, "poses_lm.npz"))
    for j in range(int(z["n_cases"])):
        args = [torch.from_numpy(z[f"c{j}_{k}"]).to(cuda_dev) for k in ("pcld", "mask", "ctr_of", "kp_of")]
        poses = eval_utils.cal_frame_poses_lm(*args, True, 2, bool(z[f"c{j}_flt"]), int(z[f"c{j}_obj_id"]))
        assert len(poses) == 1
        ok, err = _pose_close(poses[0], z[f"c{j}_pose"])
        assert ok, err


def test_empty_class_and_background_only(cuda_dev, golden_dir):
    z = np.load(os.path.join(golden_dir, "poses_ycb.npz"))
    args = [torch.from_numpy(z[f"c0_{k}"]).to(cuda_dev) for k in ("pcld", "mask", "ctr_of", "kp_of")]
    ids, poses = eval_utils.cal_frame_poses(args[0], torch.zeros_like(args[1]), args[2], args[3], True, 22, True)
    assert len(ids) == 0 and poses == []
    lm = eval_utils.cal_frame_poses_lm(args[0], torch.zeros_like(args[1]), args[2], args[3], True, 2, False, 1)
    assert np.array_equal(lm[0], np.identity(4)[:3, :])       # pvn3d_eval_utils.py:171-172


def test_batched_solver_equals_per_frame(cuda_dev, golden_dir):
    z = np.load(os.path.join(golden_dir, "poses_ycb.npz"))
    # cases 0 and 2 share N = 2048 -> one batch of two different frames
    pc = torch.from_numpy(np.stack([z["c0_pcld"], z["c2_pcld"]])).to(cuda_dev)
    mk = torch.from_numpy(np.stack([z["c0_mask"], z["c2_mask"]]).astype(np.int32)).to(cuda_dev)
    co = torch.from_numpy(np.stack([z["c0_ctr_of"][0], z["c2_ctr_of"][0]])).to(cuda_dev)
    ko = torch.from_numpy(np.stack([z["c0_kp_of"], z["c2_kp_of"]])).to(cuda_dev)
    s = FramePoseSolver(2, 2048, 8, 22, fixtures.mesh_kps_table_ycb(), fixtures.radius_thresholds_ycb(), True,
                        device=cuda_dev)
    poses, present, _, _ = s.solve(pc, mk, co, ko)
    poses = poses.double().cpu().numpy()
    for bi, j in enumerate((0, 2)):
        ids = z[f"c{j}_ids"]
        assert np.array_equal(np.nonzero(present[bi].cpu().numpy())[0], ids)
        for c, q in zip(ids, z[f"c{j}_poses"]):
            ok, err = _pose_close(poses[bi, c], q)
            assert ok, err


def test_best_fit_transform(cuda_dev, golden_dir):
    z = np.load(os.path.join(golden_dir, "bft_cases.npz"))
    for a, b, tref in zip(z["A"], z["B"], z["T"]):
        got = eval_utils.best_fit_transform(a, b)
        assert got.shape == (3, 4) and got.dtype == np.float64
        ok, err = _pose_close(got, tref)
        assert ok, err
        assert abs(np.linalg.det(got[:, :3]) - 1.0) < 1e-5
