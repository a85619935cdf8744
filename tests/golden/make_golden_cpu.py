"""Generate the CPU golden vectors from the REFERENCE ITSELF (run in the build container only).

    python tests/golden/make_golden_cpu.py

Imports the unmodified reference Python from /root/reference/pvn3d (with the import shims of
pvn3d_b200.compat) and records, on seeded synthetic inputs:
  ms_cases.npz      MeanShiftTorch(0.08).fit                  (meanshift_pytorch.py:18-51)
  bft_cases.npz     best_fit_transform                         (basic_utils.py:47-80)
  poses_ycb.npz     cal_frame_poses   (use_ctr_clus_flter=True) (pvn3d_eval_utils.py:37-110)
  poses_lm.npz      cal_frame_poses_lm                         (pvn3d_eval_utils.py:156-201)
  pn2msg.npz        Pointnet2MSG.forward                       (pvn3d.py:126-154) on CPU, the nine
                    `_ext` ops served by oracle/ext_cpu.py
cal_frame_poses* hard-code .cuda(); they run here by making Tensor.cuda / Module.cuda a no-op.
While recording, every oracle restatement is checked bit-for-bit against the reference call it
restates; a mismatch aborts.  The .npz files are committed; this script is their provenance.
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = "/root/reference/pvn3d"
OUT = os.path.dirname(os.path.abspath(__file__))

from pvn3d_b200 import compat, fixtures, synth, testing  # noqa: E402
from oracle import ext_cpu, frame_poses_oracle, meanshift_oracle, metrics_oracle  # noqa: E402

compat.install_import_shims()
sys.modules["lib.pointnet2_utils._ext"] = ext_cpu
sys.path.insert(0, REF)
torch.Tensor.cuda = lambda self, *a, **k: self          # run the reference's .cuda() code on CPU
torch.nn.Module.cuda = lambda self, *a, **k: self

from lib.utils import meanshift_pytorch as ref_ms       # noqa: E402
from lib.utils import basic_utils as ref_bu             # noqa: E402
from lib.utils import pvn3d_eval_utils as ref_eval      # noqa: E402
from lib import pvn3d as ref_pvn3d                       # noqa: E402


def vote_cloud(rng, n, sigma, outlier, centre=(0.1, -0.05, 0.8), clusters=1):
    pts = []
    per = n // clusters
    for c in range(clusters):
        ctr = np.array(centre) + c * np.array([0.25, 0.1, 0.05])
        pts.append(ctr + rng.normal(0, sigma, size=(per, 3)))
    p = np.concatenate(pts, 0)
    n_out = int(outlier * len(p))
    if n_out:
        idx = rng.choice(len(p), n_out, replace=False)
        p[idx] = rng.uniform([-0.5, -0.4, 0.6], [0.5, 0.4, 1.2], size=(n_out, 3))
    return p.astype(np.float32)


def golden_meanshift():
    cases = {}
    specs = [  # (name, n, sigma, outlier, clusters, bandwidth)
        ("tight", 300, 0.005, 0.0, 1, 0.08), ("outl10", 512, 0.005, 0.10, 1, 0.08),
        ("outl30", 700, 0.02, 0.30, 1, 0.08), ("two", 400, 0.01, 0.05, 2, 0.08),
        ("wide", 256, 0.03, 0.10, 1, 0.08), ("bw002", 300, 0.005, 0.10, 1, 0.02),
        ("bw016", 300, 0.02, 0.10, 1, 0.16), ("single", 1, 0.0, 0.0, 1, 0.08),
        ("pair_far", 2, 0.0, 0.0, 1, 0.08), ("n1200", 1200, 0.006, 0.12, 1, 0.08),
    ]
    for i, (name, n, sigma, outl, ncl, bw) in enumerate(specs):
        rng = np.random.default_rng(100 + i)
        if name == "pair_far":
            A = np.array([[0.0, 0.0, 0.8], [0.5, 0.0, 0.8]], np.float32)
        elif name == "single":
            A = np.array([[0.1, 0.2, 0.7]], np.float32)
        else:
            A = vote_cloud(rng, n, sigma, outl, clusters=ncl)
        At = torch.from_numpy(A)
        ctr, labels = ref_ms.MeanShiftTorch(bandwidth=bw).fit(At)
        orc = meanshift_oracle.MeanShiftOracle(bandwidth=bw)
        octr, olab = orc.fit(At)
        assert torch.equal(ctr, octr) and torch.equal(labels, olab), f"oracle != reference on {name}"
        cases[f"{name}_A"] = A
        cases[f"{name}_bw"] = np.float64(bw)
        cases[f"{name}_ctr"] = ctr.numpy()
        cases[f"{name}_labels"] = labels.numpy()
        cases[f"{name}_iters"] = np.int32(orc.n_iter)
        print(f"  meanshift {name}: n={len(A)} iters={orc.n_iter} inliers={int(labels.sum())}")
    np.savez_compressed(os.path.join(OUT, "ms_cases.npz"), **cases)


def golden_best_fit():
    rng = np.random.default_rng(7)
    A, B, T = [], [], []
    for i in range(16):
        a = rng.uniform(-0.1, 0.1, size=(9, 3)).astype(np.float32)
        R = synth._haar_rotation(rng)
        if i == 3:
            R = R @ np.diag([1, 1, -1.0])        # forces the reflection branch
        b = (a @ R.T + rng.uniform(-0.5, 0.5, 3) + rng.normal(0, 1e-3, size=(9, 3))).astype(np.float32)
        t_ref = ref_bu.best_fit_transform(a, b)
        assert np.array_equal(t_ref, meanshift_oracle.best_fit_transform(a, b))
        A.append(a); B.append(b); T.append(t_ref)
    np.savez_compressed(os.path.join(OUT, "bft_cases.npz"), A=np.stack(A), B=np.stack(B), T=np.stack(T))


def _frame_tensors(f):
    return (torch.from_numpy(f.pcld), torch.from_numpy(f.labels), torch.from_numpy(f.ctr_of), torch.from_numpy(f.kp_of))


def golden_poses():
    bs = ref_eval.bs_utils
    out = {}
    for j, (seed, n_pts, n_inst) in enumerate([(11, 2048, 3), (12, 3072, 5), (13, 2048, 2)]):
        f = synth.make_frame("ycb", n_points=n_pts, seed=seed, n_instances=n_inst)
        pcld, mask, ctr_of, kp_of = _frame_tensors(f)
        if j == 2:   # mislabel a slab of one object as another class: exercises the relabel pass
            c0, c1 = int(f.cls_ids[0]), int(f.cls_ids[1])
            idx = torch.nonzero(mask == c0).flatten()[:60]
            mask = mask.clone(); mask[idx] = c1
        t0 = time.time()
        ids, poses = ref_eval.cal_frame_poses(pcld, mask, ctr_of, kp_of, True, 22, True)
        oids, oposes, omask, okps = frame_poses_oracle.cal_frame_poses(
            pcld, mask, ctr_of, kp_of, True, 22, True,
            lambda c: bs.get_kps(ref_eval.cls_lst[c - 1]), lambda c: bs.get_ctr(ref_eval.cls_lst[c - 1]),
            ref_eval.config.ycb_r_lst)
        assert np.array_equal(ids, oids) and all(np.array_equal(a, b) for a, b in zip(poses, oposes))
        print(f"  cal_frame_poses case {j}: classes {ids.tolist()} ({time.time() - t0:.1f}s), relabelled "
              f"{int((omask != mask).sum())} pts")
        out.update({f"c{j}_pcld": f.pcld, f"c{j}_mask": mask.numpy(), f"c{j}_ctr_of": f.ctr_of, f"c{j}_kp_of": f.kp_of,
                    f"c{j}_ids": ids, f"c{j}_poses": np.stack(poses), f"c{j}_new_mask": omask.numpy(),
                    f"c{j}_cls_kps": okps.numpy()})
    out["n_cases"] = np.int32(3)
    np.savez_compressed(os.path.join(OUT, "poses_ycb.npz"), **out)

    out = {}
    bs_lm = ref_eval.bs_utils_lm
    for j, (seed, n_pts, flt) in enumerate([(21, 2048, False), (22, 3072, True)]):
        f = synth.make_frame("linemod", n_points=n_pts, seed=seed, obj_frac=0.12)
        pcld, mask, ctr_of, kp_of = _frame_tensors(f)
        poses = ref_eval.cal_frame_poses_lm(pcld, mask, ctr_of, kp_of, True, 2, flt, f.obj_id)
        oposes, okps = frame_poses_oracle.cal_frame_poses_lm(
            pcld, mask, ctr_of, kp_of, True, 2, flt,
            bs_lm.get_kps(f.obj_id, ds_type="linemod"), bs_lm.get_ctr(f.obj_id, ds_type="linemod"))
        assert np.array_equal(poses[0], oposes[0])
        print(f"  cal_frame_poses_lm case {j}: obj {f.obj_id}, n_c={int((mask == 1).sum())}")
        out.update({f"c{j}_pcld": f.pcld, f"c{j}_mask": mask.numpy(), f"c{j}_ctr_of": f.ctr_of, f"c{j}_kp_of": f.kp_of,
                    f"c{j}_obj_id": np.int32(f.obj_id), f"c{j}_flt": np.bool_(flt), f"c{j}_pose": poses[0],
                    f"c{j}_cls_kps": okps.numpy()})
    out["n_cases"] = np.int32(2)
    np.savez_compressed(os.path.join(OUT, "poses_lm.npz"), **out)
    # fixtures must agree with what the reference loads from its txt files
    for c in range(1, 22):
        assert np.array_equal(bs.get_kps(ref_eval.cls_lst[c - 1]), fixtures.get_kps(c))
        assert np.array_equal(bs.get_ctr(ref_eval.cls_lst[c - 1]), fixtures.get_ctr(c))
    assert ref_eval.config.ycb_r_lst == fixtures.ycb_r_lst()


def golden_pn2msg():
    torch.manual_seed(0)
    ref_model = ref_pvn3d.Pointnet2MSG(input_channels=6)
    testing.randomize_bn_(ref_model, 1)
    ref_model.eval()
    mine = testing.seeded_pointnet2msg(0, 1)
    sd_ref, sd_mine = ref_model.state_dict(), mine.state_dict()
    assert list(sd_ref.keys()) == list(sd_mine.keys()), "state_dict keys differ from the reference"
    for k in sd_ref:
        assert torch.equal(sd_ref[k], sd_mine[k]), f"parameter {k} differs"
    print(f"  Pointnet2MSG mirror: {len(sd_ref)} state_dict entries identical to the reference")
    f = synth.make_frame("linemod", n_points=4096, seed=31)
    x = torch.from_numpy(f.cld_rgb_nrm)[None]
    t0 = time.time()
    with torch.no_grad():
        y = ref_model(x)                                   # [1,128,4096]
    print(f"  reference Pointnet2MSG forward on CPU: {time.time() - t0:.1f}s, out {tuple(y.shape)}")
    cols = np.sort(np.random.default_rng(5).choice(4096, 768, replace=False))
    np.savez_compressed(os.path.join(OUT, "pn2msg.npz"), cld_rgb_nrm=f.cld_rgb_nrm, cols=cols.astype(np.int32),
                        feats=y[0][:, cols].numpy(), feat_mean=np.float64(y.double().mean()),
                        feat_abs_mean=np.float64(y.double().abs().mean()))


def golden_metrics():
    """Basic_Utils.cal_add_cuda / cal_adds_cuda (basic_utils.py:617-635) on CPU float32 tensors: a random
    mesh of 1500 points, pose pairs from identical to 5 cm / 20 degrees apart, and a symmetric object
    (ADD large, ADD-S ~ 0)."""
    bs = ref_eval.bs_utils
    rng = np.random.default_rng(41)
    mesh = rng.uniform(-0.08, 0.08, size=(1500, 3)).astype(np.float32)
    sym = np.concatenate([mesh[:750], -mesh[:750]], 0)            # point-symmetric: R = -I maps it onto itself
    pred, gt, which, add, adds = [], [], [], [], []
    for i in range(6):
        Rg = synth._haar_rotation(rng)
        tg = rng.uniform([-0.2, -0.2, 0.6], [0.2, 0.2, 1.0])
        ang = [0.0, 0.01, 0.05, 0.35, 0.0, 0.1][i]
        ax = rng.normal(size=3); ax /= np.linalg.norm(ax)
        K = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
        dR = np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * K @ K
        Rp = dR @ Rg
        tp = tg + rng.normal(0, [0.0, 0.002, 0.01, 0.05, 0.0, 0.003][i], 3)
        use_sym = i == 4
        if use_sym:
            Rp = -Rg                                                # improper but isometric: exercises ADD >> ADD-S
        m = sym if use_sym else mesh
        P = np.concatenate([Rp, tp[:, None]], 1).astype(np.float32)
        G = np.concatenate([Rg, tg[:, None]], 1).astype(np.float32)
        a = bs.cal_add_cuda(torch.from_numpy(P), torch.from_numpy(G), torch.from_numpy(m))
        s = bs.cal_adds_cuda(torch.from_numpy(P), torch.from_numpy(G), torch.from_numpy(m))
        assert torch.equal(a, metrics_oracle.cal_add(torch.from_numpy(P), torch.from_numpy(G), torch.from_numpy(m)))
        assert torch.equal(s, metrics_oracle.cal_adds(torch.from_numpy(P), torch.from_numpy(G), torch.from_numpy(m)))
        pred.append(P); gt.append(G); which.append(int(use_sym)); add.append(float(a)); adds.append(float(s))
        print(f"  metrics case {i}: ADD {float(a):.6f}  ADD-S {float(s):.6f}")
    np.savez_compressed(os.path.join(OUT, "metrics.npz"), mesh=mesh, sym=sym, pred=np.stack(pred), gt=np.stack(gt),
                        which=np.array(which, np.int32), add=np.array(add, np.float32), adds=np.array(adds, np.float32))


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count())
    which = sys.argv[1:] or ["ms", "bft", "poses", "pn2msg", "metrics"]
    if "ms" in which:
        golden_meanshift()
    if "bft" in which:
        golden_best_fit()
    if "poses" in which:
        golden_poses()
    if "pn2msg" in which:
        golden_pn2msg()
    if "metrics" in which:
        golden_metrics()
    print("done")
