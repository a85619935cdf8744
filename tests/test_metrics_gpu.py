"""Callers either side of the hot path on device: segmentation argmax
(demo.py:108) and the ADD / ADD-S pose errors (basic_utils.py:617-635) against values recorded from
the reference's own methods (tests/golden/metrics.npz)."""
import os

import numpy as np
import pytest
import torch

from pvn3d_b200 import eval_utils, fixtures, synth
from pvn3d_b200.eval_utils import FramePoseSolver

pytestmark = pytest.mark.gpu


def test_seg_argmax_is_torch_max(cuda_dev):
    g = torch.Generator().manual_seed(3)
    for shape in [(2, 12288, 22), (1, 4096, 2), (7, 1)]:
        x = torch.randn(*shape, generator=g).to(cuda_dev)
        x[..., 0][x[..., 0] > 1.0] = 5.0               # ties between classes -> first index wins
        if shape[-1] > 3:
            x[..., 3][x[..., 0] == 5.0] = 5.0
        got = eval_utils.seg_argmax(x)
        want = torch.max(x.cpu(), -1)[1]               # CPU torch: first maximal index
        assert got.dtype == torch.int32 and got.shape == x.shape[:-1]
        assert torch.equal(got.cpu().long(), want)


def test_poses_from_logits_equal_poses_from_labels(cuda_dev):
    """argmax-seg -> class compaction -> votes on device: the pipeline input is the network output"""
    n = 4096
    f = synth.make_frame("ycb", n_points=n, seed=5)
    logits = torch.full((1, n, 22), -4.0)
    logits[0, torch.arange(n), torch.from_numpy(f.labels)] = 3.0
    logits += torch.rand(1, n, 22, generator=torch.Generator().manual_seed(1))     # < 1: the arg-max is unchanged
    mask = eval_utils.seg_argmax(logits.to(cuda_dev))
    assert np.array_equal(mask[0].cpu().numpy(), f.labels.astype(np.int32))
    s = FramePoseSolver(1, n, 8, 22, fixtures.mesh_kps_table_ycb(), fixtures.radius_thresholds_ycb(), True, device=cuda_dev)
    args = [torch.from_numpy(x).to(cuda_dev) for x in (f.pcld[None], f.ctr_of, f.kp_of[None])]
    p1 = s.solve(args[0], mask, args[1], args[2])[0].clone()
    p2 = s.solve(args[0], torch.from_numpy(f.labels.astype(np.int32))[None].to(cuda_dev), args[1], args[2])[0]
    assert torch.equal(p1, p2)


def test_add_adds_match_reference(cuda_dev, golden_dir):
    z = np.load(os.path.join(golden_dir, "metrics.npz"))
    for use_sym in (0, 1):
        sel = np.nonzero(z["which"] == use_sym)[0]
        mesh = torch.from_numpy(z["sym"] if use_sym else z["mesh"]).to(cuda_dev)
        add, adds = eval_utils.pose_add_adds(torch.from_numpy(z["pred"][sel]), torch.from_numpy(z["gt"][sel]), mesh)
        add, adds = add.cpu().numpy(), adds.cpu().numpy()
        for k, i in enumerate(sel):
            # fp32 sums in a different order than torch.mm / torch.mean: 1e-5 relative (+ 1e-7 m absolute at zero)
            assert abs(add[k] - z["add"][i]) <= 1e-5 * z["add"][i] + 1e-7, (i, add[k], z["add"][i])
            assert abs(adds[k] - z["adds"][i]) <= 1e-5 * z["adds"][i] + 1e-7, (i, adds[k], z["adds"][i])
    # reference-shaped entry points (0-dim tensors), ADD-S <= ADD
    a = eval_utils.cal_add_cuda(torch.from_numpy(z["pred"][2]), torch.from_numpy(z["gt"][2]), torch.from_numpy(z["mesh"]).to(cuda_dev))
    s = eval_utils.cal_adds_cuda(torch.from_numpy(z["pred"][2]), torch.from_numpy(z["gt"][2]), torch.from_numpy(z["mesh"]).to(cuda_dev))
    assert a.dim() == 0 and s.dim() == 0 and float(s) <= float(a)


def test_adds_large_mesh_properties(cuda_dev):
    """multi-tile mesh (P > 2048): ADD-S of identical poses is 0, of a pure translation d is <= |d| and
    reproducible bit for bit"""
    rng = np.random.default_rng(0)
    mesh = torch.from_numpy(rng.uniform(-0.1, 0.1, (7001, 3)).astype(np.float32)).to(cuda_dev)
    G = torch.eye(4)[:3][None].clone()
    P = G.clone()
    add, adds = eval_utils.pose_add_adds(P, G, mesh)
    assert float(add[0]) == 0.0 and float(adds[0]) == 0.0
    P[0, :, 3] = torch.tensor([0.003, -0.004, 0.0])
    add, adds = eval_utils.pose_add_adds(P, G, mesh)
    assert abs(float(add[0]) - 0.005) < 1e-6 and 0.0 < float(adds[0]) <= 0.005 + 1e-7
    add2, adds2 = eval_utils.pose_add_adds(P, G, mesh)
    assert torch.equal(add, add2) and torch.equal(adds, adds2)
