"""Record outputs of the UNMODIFIED reference op library (oracle/_ref/_ext.so) on the GPU box.

    gpurun -- python tests/golden/make_golden_gpu.py        # writes gpurun_out/pn2_ref.npz

The file is then committed as tests/golden/pn2_ref.npz; tests/test_oracle_cpu.py checks the C oracle
against it on CPU (this is what pins oracle/pn2_oracle.c to the reference), and the -m gpu tests check
the sm_100a kernels against it.  Inputs are regenerated from seeds (pvn3d_b200.synth), only the small
reference OUTPUTS are stored.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import load_ref_ext  # noqa: E402
from pvn3d_b200 import synth  # noqa: E402

SPEC = dict(n=4096, seed=77, fps_m=(1024, 256), bq=((0.03, 16), (0.06, 32)), c_group=8, c_interp=16)


def inputs():
    f = synth.make_frame("ycb", n_points=SPEC["n"], seed=SPEC["seed"])
    xyz = f.pcld[None].copy()
    xyz[0, 4000:4090] = xyz[0, 100:190]          # wrap-style duplicates -> exact FPS / 3-NN ties
    rng = np.random.default_rng(SPEC["seed"])
    feats = rng.normal(size=(1, SPEC["c_group"], SPEC["n"])).astype(np.float32)
    return xyz, feats


def main():
    ref = load_ref_ext()
    assert ref is not None, "oracle/_ref/_ext.so missing"
    dev = torch.device("cuda:0")
    xyz, feats = inputs()
    X = torch.from_numpy(xyz).to(dev)
    out = {}
    lvl = X
    for i, m in enumerate(SPEC["fps_m"]):
        idx = ref.furthest_point_sampling(lvl, m)
        out[f"fps{i}"] = idx.cpu().numpy()
        nxt = ref.gather_points(lvl.transpose(1, 2).contiguous(), idx).transpose(1, 2).contiguous()
        out[f"new_xyz{i}"] = nxt.cpu().numpy()
        if i == 0:
            for j, (r, ns) in enumerate(SPEC["bq"]):
                bq = ref.ball_query(nxt, lvl, r, ns)
                out[f"bq{j}"] = bq.cpu().numpy()
                if j == 0:
                    out["group0"] = ref.group_points(torch.from_numpy(feats).to(dev), bq).cpu().numpy()
            d2, nn = ref.three_nn(lvl, nxt)
            out["nn_d2"], out["nn_idx"] = d2.cpu().numpy(), nn.cpu().numpy()
            rng = np.random.default_rng(1)
            pf = rng.normal(size=(1, SPEC["c_interp"], m)).astype(np.float32)
            w = rng.uniform(size=(1, SPEC["n"], 3)).astype(np.float32)
            w /= w.sum(-1, keepdims=True)
            out["interp"] = ref.three_interpolate(torch.from_numpy(pf).to(dev), nn, torch.from_numpy(w).to(dev)).cpu().numpy()
        lvl = nxt
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    np.savez_compressed(os.path.join(ROOT, "gpurun_out", "pn2_ref.npz"), **out)
    print("wrote gpurun_out/pn2_ref.npz", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
