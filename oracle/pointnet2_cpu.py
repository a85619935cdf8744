"""TEST INFRASTRUCTURE / CPU BASELINE -- Pointnet2MSG.forward on CPU.

The reference's orchestration (pvn3d/lib/pvn3d.py:126-154, pointnet2_modules.py:27-71,162-206,
pointnet2_utils.py:293-330) restated as plain functions over numpy/torch-CPU: the nine `_ext` ops are
served by the C oracle (oracle/pn2.py), the shared MLPs by torch CPU conv2d + batch_norm (eval) + relu
with the weights of a `state_dict` in the reference's key layout.  Used by bench.py's cpu_baseline and
`--impl reference` legs (timed) and by tests as a second opinion on hot path A.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from . import pn2

SA_SPEC = ((2048, (0.0175, 0.025), (16, 32)), (1024, (0.025, 0.05), (16, 32)),
           (512, (0.05, 0.1), (16, 32)), (128, (0.1, 0.2), (16, 32)))     # pvn3d.py:65-111


def _shared_mlp(x: torch.Tensor, sd, prefix: str) -> torch.Tensor:
    i = 0
    while f"{prefix}.layer{i}.conv.weight" in sd:
        p = f"{prefix}.layer{i}"
        x = F.conv2d(x, sd[f"{p}.conv.weight"])
        x = F.batch_norm(x, sd[f"{p}.normlayer.bn.running_mean"], sd[f"{p}.normlayer.bn.running_var"],
                         sd[f"{p}.normlayer.bn.weight"], sd[f"{p}.normlayer.bn.bias"], False, 0.0, 1e-5)
        x = F.relu(x)
        i += 1
    return x


@torch.no_grad()
def forward(pointcloud: np.ndarray, sd, threads=None) -> np.ndarray:
    """pointcloud [B,N,9] f32 -> features [B,128,N] f32"""
    sd = {k: v.detach().cpu() for k, v in sd.items()}
    xyz = np.ascontiguousarray(pointcloud[..., :3])
    feats = np.ascontiguousarray(pointcloud[..., 3:].transpose(0, 2, 1))
    l_xyz, l_feat = [xyz], [feats]
    for li, (npoint, radii, nsamples) in enumerate(SA_SPEC):
        x, f = l_xyz[-1], l_feat[-1]
        idx = pn2.furthest_point_sampling(x, npoint, threads)
        new_xyz = np.take_along_axis(x, idx[..., None].astype(np.int64).repeat(3, -1), 1)
        outs = []
        for si, (r, ns) in enumerate(zip(radii, nsamples)):
            g, _ = pn2.query_and_group(x, new_xyz, f, float(np.float32(r)), ns, threads)
            h = _shared_mlp(torch.from_numpy(g), sd, f"SA_modules.{li}.mlps.{si}")
            outs.append(F.max_pool2d(h, kernel_size=[1, h.size(3)]).squeeze(-1))
        l_xyz.append(new_xyz)
        l_feat.append(torch.cat(outs, dim=1).numpy())
    for i in range(-1, -5, -1):
        unknown, known = l_xyz[i - 1], l_xyz[i]
        d2, idx = pn2.three_nn(unknown, known, threads)
        dist = torch.sqrt(torch.from_numpy(d2))
        dist_recip = 1.0 / (dist + 1e-8)
        weight = dist_recip / torch.sum(dist_recip, dim=2, keepdim=True)
        interp = pn2.three_interpolate(l_feat[i], idx, weight.numpy())
        cat = np.concatenate([interp, l_feat[i - 1]], axis=1)
        h = _shared_mlp(torch.from_numpy(cat).unsqueeze(-1), sd, f"FP_modules.{4 + i}.mlp")
        l_feat[i - 1] = h.squeeze(-1).numpy()
    return l_feat[0]
