"""Parity AT THE BENCHMARK'S SIZES against outputs recorded from the reference itself
(tests/golden/*_big.npz, provenance: tests/golden/make_golden_big.py).

  * MeanShiftTorch.fit at n = 3900 (bandwidth 0.02 / 0.04 / 0.16), n = 6000 / 9000 (the multi-tile
    sweep), and the unstable-midpoint case -- meanshift_pytorch.py:24-51;
  * cal_frame_poses_lm on the first frame of the bench batch (12288 points, n_c = 3348 votes, the 9
    fits bench.py times) -- pvn3d_eval_utils.py:156-201;
  * cal_frame_poses on a 12288-point YCB frame with a mislabelled slab -- pvn3d_eval_utils.py:37-110;
  * Pointnet2MSG.forward at N = 12288 -- pvn3d.py:126-154.

Bar (BASELINE.json north_star): labels / inlier counts / relabelled instance mask / class ids
bit-exact; voted centres and poses within 1e-4 RELATIVE; iteration counts of the modes that apply
the reference's global stop rule within +-1 of the reference's.
"""
import os

import numpy as np
import pytest
import torch

from pvn3d_b200 import eval_utils, fixtures, mlp, testing
from pvn3d_b200.eval_utils import FramePoseSolver
from pvn3d_b200.meanshift import MeanShiftTorch

pytestmark = pytest.mark.gpu
REL_TOL = 1e-4
MODES = ["certified", "early_exit", "strict", "no_freeze"]
MS_BIG = ["bw002_n3900", "bw004_n3900", "bw016_n3900", "n6000", "n9000", "sym2"]


def _ms(bw, mode):
    return MeanShiftTorch(bandwidth=bw, mode=mode)


def _rel(a, b):
    return float(np.linalg.norm(np.asarray(a, np.float64) - np.asarray(b, np.float64)) / max(np.linalg.norm(b), 1e-12))


def _pose_close(p, q):
    dr = np.linalg.norm(p[:, :3] - q[:, :3])
    dt = np.linalg.norm(p[:, 3] - q[:, 3]) / max(np.linalg.norm(q[:, 3]), 1e-9)
    return dr <= REL_TOL * np.sqrt(3) and dt <= REL_TOL, (dr, dt)


def _check_iters(mode, got, want):
    if mode in ("strict", "no_freeze"):
        assert abs(got - want) <= 1, (mode, got, want)     # same global stop rule (a borderline shift may move it by one)
    else:
        assert got <= want + 1, (mode, got, want)          # stops at the stationary returned seed, never later


@pytest.mark.parametrize("name", MS_BIG)
@pytest.mark.parametrize("mode", MODES)
def test_meanshift_full_size(cuda_dev, golden_dir, name, mode):
    z = np.load(os.path.join(golden_dir, "ms_big.npz"))
    A = torch.from_numpy(z[f"{name}_A"]).to(cuda_dev)
    n = A.size(0)
    ms = _ms(float(z[f"{name}_bw"]), mode)
    ctr, labels = ms.fit(A)
    want_labels = np.unpackbits(z[f"{name}_labels"])[:n].astype(bool)
    assert np.array_equal(labels.cpu().numpy(), want_labels), "labels must be bit-exact"
    assert _rel(ctr.cpu().numpy(), z[f"{name}_ctr"]) <= REL_TOL
    _check_iters(mode, int(ms.last_iters[0].item()), int(z[f"{name}_iters"]))


@pytest.mark.parametrize("flt", [False, True], ids=["raw", "flt"])
@pytest.mark.parametrize("mode", MODES)
def test_bench_frame_linemod(cuda_dev, golden_dir, flt, mode):
    """the first frame of bench.py's batch: 9 fits at n_c = 3348 (raw = what bench.py times)"""
    z = np.load(os.path.join(golden_dir, "poses_lm_big.npz"))
    tag = "flt" if flt else "raw"
    pcld, ctr_of, kp_of = (torch.from_numpy(z[k]).to(cuda_dev) for k in ("pcld", "ctr_of", "kp_of"))
    mask = torch.from_numpy(z["mask"].astype(np.int64)).to(cuda_dev)
    obj_id = int(z["obj_id"])
    poses = eval_utils.cal_frame_poses_lm(pcld, mask, ctr_of, kp_of, True, 2, flt, obj_id, mode=mode)
    ok, err = _pose_close(poses[0], z[f"{tag}_pose"])
    assert ok, err
    # every fit the reference executed (centre fit first, then the 8 keypoints, :175-189)
    sel = mask == 1
    assert int(sel.sum()) == int(z["n_c"])
    ms = _ms(0.08, mode)
    ctr, ctr_labels = ms.fit((pcld - ctr_of[0])[sel])
    want_labels = np.unpackbits(z[f"{tag}_ctr_labels"])[:int(z["n_c"])].astype(bool)
    assert np.array_equal(ctr_labels.cpu().numpy(), want_labels), "centre-cluster labels must be bit-exact"
    assert int(ctr_labels.sum()) == int(z[f"{tag}_fit_n_in"][0])
    assert _rel(ctr.cpu().numpy(), z[f"{tag}_fit_ctr"][0]) <= REL_TOL
    _check_iters(mode, int(ms.last_iters[0].item()), int(z[f"{tag}_fit_iters"][0]))
    clouds = []
    for k in range(kp_of.size(0)):
        v = (pcld - kp_of[k])[sel]
        clouds.append(v[ctr_labels] if flt else v)
    ctrs, _ = ms.fit_many(clouds)
    iters = ms.last_iters.cpu().numpy().astype(int)
    for k in range(len(clouds)):
        assert clouds[k].size(0) == int(z[f"{tag}_fit_n"][1 + k])
        assert _rel(ctrs[k].cpu().numpy(), z[f"{tag}_fit_ctr"][1 + k]) <= REL_TOL, k
        _check_iters(mode, int(iters[k]), int(z[f"{tag}_fit_iters"][1 + k]))


@pytest.mark.parametrize("mode", MODES)
def test_bench_frame_ycb(cuda_dev, golden_dir, mode):
    z = np.load(os.path.join(golden_dir, "poses_ycb_big.npz"))
    pcld, ctr_of, kp_of = (torch.from_numpy(z[k]).to(cuda_dev) for k in ("pcld", "ctr_of", "kp_of"))
    mask = torch.from_numpy(z["mask"].astype(np.int64)).to(cuda_dev)
    ids, poses = eval_utils.cal_frame_poses(pcld, mask, ctr_of, kp_of, True, 22, True, mode=mode)
    assert np.array_equal(ids, z["ids"])
    for p, q in zip(poses, z["poses"]):
        ok, err = _pose_close(p, q)
        assert ok, err
    n = pcld.size(0)
    s = FramePoseSolver(1, n, 8, 22, fixtures.mesh_kps_table_ycb(), fixtures.radius_thresholds_ycb(), True,
                        device=cuda_dev, mode=mode)
    _, present, cls_kps, new_mask = s.solve(pcld[None].contiguous(), mask[None].to(torch.int32).contiguous(),
                                            ctr_of[0][None].contiguous(), kp_of[None].contiguous())
    assert np.array_equal(new_mask[0].cpu().numpy(), z["new_mask"].astype(np.int32)), "instance labels must be bit-exact"
    assert int((z["new_mask"] != z["mask"]).sum()) > 200, "the relabel pass must have had work"
    got, want = cls_kps[0].cpu().numpy(), z["cls_kps"]
    for c in ids:
        for k in range(9):
            assert _rel(got[c, k], want[c, k]) <= REL_TOL, (c, k)
    assert np.array_equal(np.nonzero(present[0].cpu().numpy())[0], ids)


def test_pointnet2msg_full_size(cuda_dev, golden_dir):
    """fused engine vs the reference module's fp32 CPU features at N = 12288 (TF32 operand class, as
    tests/test_mlp_gpu.py::test_fused_pointnet2msg_matches_reference_features at N = 4096)"""
    z = np.load(os.path.join(golden_dir, "pn2msg_big.npz"))
    model = testing.seeded_pointnet2msg(0, 1)
    eng = mlp.FusedPointnet2MSG(model, cuda_dev)
    x = torch.from_numpy(z["cld_rgb_nrm"])[None].to(cuda_dev)
    y = eng(x)
    assert y.shape == (1, 128, 12288)
    cols = torch.from_numpy(z["cols"]).long().to(cuda_dev)
    got = y[0][:, cols].cpu().numpy()
    scale = float(z["feat_abs_mean"])
    err = np.abs(got - z["feats"])
    assert err.mean() <= 3e-3 * scale and err.max() <= 5e-2 * scale, (err.mean() / scale, err.max() / scale)
    # fp32 module graph on the library's `_ext` ops: identical indices, only summation order differs
    prev = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    try:
        with torch.no_grad():
            y32 = model.to(cuda_dev)(x)[0][:, cols].cpu().numpy()
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = prev
    assert np.abs(y32 - z["feats"]).max() <= 2e-3 * max(scale, 1.0)
    assert np.abs(y32 - z["feats"]).mean() <= 1e-4 * max(scale, 1.0)
