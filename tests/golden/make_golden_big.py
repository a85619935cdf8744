"""Golden vectors AT THE BENCHMARK'S SIZES, recorded from the REFERENCE ITSELF (build container only).

    python tests/golden/make_golden_big.py [ms] [lm] [ycb] [pn2msg]

Round-1 goldens stopped at n = 1200 votes / 3072-point frames; the benchmark runs 12288-point
frames with n_c ~ 3348 votes per fit and 98-237 sweeps per fit.  This script records, with the
unmodified reference Python imported from /root/reference/pvn3d (same shims as make_golden_cpu.py):

  ms_big.npz      MeanShiftTorch.fit (meanshift_pytorch.py:24-51) on
                    * bandwidth 0.02 / 0.04 / 0.16 at n = 3900,
                    * n = 6000 and n = 9000 (the multi-tile sweep of csrc/meanshift.cu, n > 4096),
                    * two mirror-symmetric clusters with a seed on the unstable midpoint (sym2);
  poses_lm_big.npz   cal_frame_poses_lm (pvn3d_eval_utils.py:156-201) on the FIRST FRAME OF THE
                     BENCH batch (synth seed 2000, 12288 points, n_c = 3348), unfiltered (what bench.py
                     times) and filtered keypoint votes;
  poses_ycb_big.npz  cal_frame_poses (pvn3d_eval_utils.py:37-110) on the first frame of the YCB bench
                     batch (synth seed 3000, 12288 points, 5 instances) with a slab of one object
                     mislabelled as its neighbour, so the relabel pass (:58-72) has work;
  pn2msg_big.npz     Pointnet2MSG.forward (pvn3d.py:126-154) at N = 12288 on the bench frame.

Every MeanShiftTorch.fit the reference executes inside cal_frame_poses* is RECORDED (centre, label
count, iteration count = number of gaussian_kernel calls) by a subclass that only wraps the untouched
reference `fit`; the oracle restatements are checked bit-for-bit against the reference outputs while
recording.  The frame inputs are stored (kp_of/ctr_of are zero off-object and compress well).
"""
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden_cpu as g  # noqa: E402  (installs the shims, imports the reference modules)

from pvn3d_b200 import synth, testing  # noqa: E402
from oracle import frame_poses_oracle, meanshift_oracle  # noqa: E402

ref_ms, ref_eval, ref_pvn3d = g.ref_ms, g.ref_eval, g.ref_pvn3d
OUT = g.OUT

_calls = [0]
_orig_kernel = ref_ms.gaussian_kernel


def _counting_kernel(distance, bandwidth):
    _calls[0] += 1
    return _orig_kernel(distance, bandwidth)


ref_ms.gaussian_kernel = _counting_kernel          # one call per sweep (meanshift_pytorch.py:36)


class RecordingMeanShift(ref_ms.MeanShiftTorch):
    """the reference class, its fit() untouched; logs what went in and what came out"""
    log = []

    def fit(self, A):
        c0 = _calls[0]
        t0 = time.time()
        ctr, labels = super().fit(A)
        RecordingMeanShift.log.append(dict(n=int(A.shape[0]), iters=_calls[0] - c0, ctr=ctr.numpy().copy(),
                                           n_in=int(labels.sum()), labels=labels.numpy().copy()))
        print(f"      fit n={A.shape[0]} iters={_calls[0] - c0} ({time.time() - t0:.0f}s)", flush=True)
        return ctr, labels


ref_eval.MeanShiftTorch = RecordingMeanShift


def _take_log():
    log, RecordingMeanShift.log = RecordingMeanShift.log, []
    return log


def golden_meanshift_big():
    cases = {}
    specs = [  # (name, n, sigma, outlier, clusters, bandwidth)
        ("bw002_n3900", 3900, 0.005, 0.10, 1, 0.02), ("bw004_n3900", 3900, 0.005, 0.10, 1, 0.04),
        ("bw016_n3900", 3900, 0.02, 0.10, 1, 0.16), ("n6000", 6000, 0.006, 0.12, 1, 0.08),
        ("n9000", 9000, 0.005, 0.05, 1, 0.08), ("sym2", 0, 0.0, 0.0, 2, 0.08),
    ]
    for i, (name, n, sigma, outl, ncl, bw) in enumerate(specs):
        rng = np.random.default_rng(500 + i)
        if name == "sym2":
            # two mirror-image clusters (x -> -x) 0.16 m apart + one seed on the unstable midpoint
            half = (np.array([0.08, 0.0, 0.8]) + rng.normal(0, 0.01, size=(400, 3))).astype(np.float32)
            mirror = half * np.array([-1, 1, 1], np.float32)
            A = np.concatenate((np.array([[0.0, 0.0, 0.8]], np.float32), half, mirror), 0)
        else:
            A = g.vote_cloud(rng, n, sigma, outl, clusters=ncl)
        At = torch.from_numpy(A)
        t0 = time.time()
        c0 = _calls[0]
        ctr, labels = ref_ms.MeanShiftTorch(bandwidth=bw).fit(At)
        iters = _calls[0] - c0
        print(f"  meanshift {name}: n={len(A)} bw={bw} iters={iters} inliers={int(labels.sum())} "
              f"({time.time() - t0:.0f}s)", flush=True)
        if len(A) <= 4000:       # the restatement is pinned at the same sizes where that is cheap
            orc = meanshift_oracle.MeanShiftOracle(bandwidth=bw)
            octr, olab = orc.fit(At)
            assert torch.equal(ctr, octr) and torch.equal(labels, olab) and orc.n_iter == iters, name
        cases[f"{name}_A"] = A
        cases[f"{name}_bw"] = np.float64(bw)
        cases[f"{name}_ctr"] = ctr.numpy()
        cases[f"{name}_labels"] = np.packbits(labels.numpy())
        cases[f"{name}_iters"] = np.int32(iters)
    cases["names"] = np.array([s[0] for s in specs])
    np.savez_compressed(os.path.join(OUT, "ms_big.npz"), **cases)


def _log_arrays(log, prefix):
    return {f"{prefix}_fit_n": np.array([e["n"] for e in log], np.int32),
            f"{prefix}_fit_iters": np.array([e["iters"] for e in log], np.int32),
            f"{prefix}_fit_ctr": np.stack([e["ctr"] for e in log]).astype(np.float32),
            f"{prefix}_fit_n_in": np.array([e["n_in"] for e in log], np.int32)}


def golden_lm_big():
    f = synth.make_frame("linemod", n_points=12288, seed=2000, lm_obj_id=1)     # bench.py frame 0 (config_id 2)
    pcld, mask, ctr_of, kp_of = g._frame_tensors(f)
    out = {"pcld": f.pcld, "mask": mask.numpy().astype(np.int8), "ctr_of": f.ctr_of, "kp_of": f.kp_of,
           "obj_id": np.int32(f.obj_id), "n_c": np.int32((mask == 1).sum())}
    bs_lm = ref_eval.bs_utils_lm
    for flt in (False, True):
        tag = "flt" if flt else "raw"
        t0 = time.time()
        poses = ref_eval.cal_frame_poses_lm(pcld, mask, ctr_of, kp_of, True, 2, flt, f.obj_id)
        log = _take_log()
        print(f"  cal_frame_poses_lm {tag}: n_c={int((mask == 1).sum())}, {sum(e['iters'] for e in log)} sweeps, "
              f"{time.time() - t0:.0f}s", flush=True)
        out[f"{tag}_pose"] = poses[0]
        out.update(_log_arrays(log, tag))
        out[f"{tag}_ctr_labels"] = np.packbits(log[0]["labels"])
        if flt:   # cheap: also pin the restatement at this size
            oposes, okps = frame_poses_oracle.cal_frame_poses_lm(
                pcld, mask, ctr_of, kp_of, True, 2, flt,
                bs_lm.get_kps(f.obj_id, ds_type="linemod"), bs_lm.get_ctr(f.obj_id, ds_type="linemod"))
            assert np.array_equal(poses[0], oposes[0])
    np.savez_compressed(os.path.join(OUT, "poses_lm_big.npz"), **out)


def golden_ycb_big():
    f = synth.make_frame("ycb", n_points=12288, seed=3000)                       # bench.py --config ycb frame 0
    pcld, mask, ctr_of, kp_of = g._frame_tensors(f)
    c0, c1 = int(f.cls_ids[0]), int(f.cls_ids[1])
    idx = torch.nonzero(mask == c0).flatten()[:250]                              # a slab of object 0 labelled as object 1
    mask = mask.clone()
    mask[idx] = c1
    bs = ref_eval.bs_utils
    t0 = time.time()
    ids, poses = ref_eval.cal_frame_poses(pcld, mask, ctr_of, kp_of, True, 22, True)
    log = _take_log()
    print(f"  cal_frame_poses: classes {ids.tolist()}, {len(log)} fits, {sum(e['iters'] for e in log)} sweeps, "
          f"{time.time() - t0:.0f}s", flush=True)
    oids, oposes, omask, okps = frame_poses_oracle.cal_frame_poses(
        pcld, mask, ctr_of, kp_of, True, 22, True,
        lambda c: bs.get_kps(ref_eval.cls_lst[c - 1]), lambda c: bs.get_ctr(ref_eval.cls_lst[c - 1]),
        ref_eval.config.ycb_r_lst)
    assert np.array_equal(ids, oids) and all(np.array_equal(a, b) for a, b in zip(poses, oposes))
    print(f"  relabelled {int((omask != mask).sum())} points", flush=True)
    out = {"pcld": f.pcld, "mask": mask.numpy().astype(np.int8), "ctr_of": f.ctr_of, "kp_of": f.kp_of,
           "ids": ids, "poses": np.stack(poses), "new_mask": omask.numpy().astype(np.int8), "cls_kps": okps.numpy()}
    out.update(_log_arrays(log, "ref"))
    np.savez_compressed(os.path.join(OUT, "poses_ycb_big.npz"), **out)


def golden_pn2msg_big():
    torch.manual_seed(0)
    ref_model = ref_pvn3d.Pointnet2MSG(input_channels=6)
    testing.randomize_bn_(ref_model, 1)
    ref_model.eval()
    f = synth.make_frame("linemod", n_points=12288, seed=2000, lm_obj_id=1)
    x = torch.from_numpy(f.cld_rgb_nrm)[None]
    t0 = time.time()
    with torch.no_grad():
        y = ref_model(x)                                   # [1,128,12288]
    print(f"  reference Pointnet2MSG forward on CPU at N=12288: {time.time() - t0:.1f}s, out {tuple(y.shape)}", flush=True)
    cols = np.sort(np.random.default_rng(6).choice(12288, 768, replace=False))
    np.savez_compressed(os.path.join(OUT, "pn2msg_big.npz"), cld_rgb_nrm=f.cld_rgb_nrm, cols=cols.astype(np.int32),
                        feats=y[0][:, cols].numpy(), feat_mean=np.float64(y.double().mean()),
                        feat_abs_mean=np.float64(y.double().abs().mean()))


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count())
    which = sys.argv[1:] or ["pn2msg", "ycb", "ms", "lm"]
    for w in which:
        {"ms": golden_meanshift_big, "lm": golden_lm_big, "ycb": golden_ycb_big, "pn2msg": golden_pn2msg_big}[w]()
    print("done")
