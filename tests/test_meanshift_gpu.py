"""Parity of the batched mean-shift kernels with the reference's recorded CPU results.

Bar (BASELINE.json north_star): labels / max_idx / inlier counts bit-exact; voted centres within
1e-4 RELATIVE of the reference's CPU MeanShift (tolerance written here: |dc| <= 1e-4 * |c|).
"""
import os

import numpy as np
import pytest
import torch

from pvn3d_b200.meanshift import MeanShiftTorch

pytestmark = pytest.mark.gpu
CASES = ["tight", "outl10", "outl30", "two", "wide", "bw002", "bw016", "single", "pair_far", "n1200"]
REL_TOL = 1e-4


def _rel_err(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-12))


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("mode", ["certified", "strict", "no_freeze", "early_exit"])
def test_fit_matches_reference_golden(cuda_dev, golden_dir, name, mode):
    """certified   = default: returned seed + witness seeds only; falls back to early_exit when the
                  witnesses cannot prove that the reference's stop rule fires late enough
    strict      = reference stop rule, stationary seeds dropped from the work lists
    no_freeze   = reference stop rule, every seed swept at every iteration (literal reference schedule)
    early_exit  = all seeds, stop as soon as the returned seed is stationary"""
    early_exit = mode in ("early_exit", "certified")
    z = np.load(os.path.join(golden_dir, "ms_cases.npz"))
    A = torch.from_numpy(z[f"{name}_A"]).to(cuda_dev)
    ms = MeanShiftTorch(bandwidth=float(z[f"{name}_bw"]), mode=mode)
    ctr, labels = ms.fit(A)
    assert labels.dtype == torch.bool and labels.shape == (A.size(0),) and ctr.shape == (3,)
    assert np.array_equal(labels.cpu().numpy(), z[f"{name}_labels"]), "labels must be bit-exact"
    assert _rel_err(ctr.cpu().numpy(), z[f"{name}_ctr"]) <= REL_TOL
    iters = int(ms.last_iters[0].item())
    if not early_exit:
        # same global stop rule => same iteration count (a borderline max-shift may move it by one)
        assert abs(iters - int(z[f"{name}_iters"])) <= 1, (iters, int(z[f"{name}_iters"]))
    else:
        assert iters <= int(z[f"{name}_iters"]) + 1


def test_fit_many_equals_individual_fits(cuda_dev, golden_dir):
    z = np.load(os.path.join(golden_dir, "ms_cases.npz"))
    names = ["tight", "outl10", "two", "wide", "single", "pair_far"]
    clouds = [torch.from_numpy(z[f"{n}_A"]).to(cuda_dev) for n in names]
    ms = MeanShiftTorch(bandwidth=0.08, mode="strict")
    ctrs, labels = ms.fit_many(clouds)
    for i, n in enumerate(names):
        c1, l1 = ms.fit(clouds[i])
        assert torch.equal(l1, labels[i])
        assert torch.equal(c1, ctrs[i]), "a fit must not depend on what else is in the batch"
        assert np.array_equal(labels[i].cpu().numpy(), z[f"{n}_labels"])


def test_labels_are_mutable_boolean_index(cuda_dev):
    # callers do ctr_labels[0] = 1 and index with it (pvn3d_eval_utils.py:86-92)
    g = torch.Generator().manual_seed(0)
    A = (torch.randn(200, 3, generator=g) * 0.01 + torch.tensor([0.1, 0.0, 0.8])).to(cuda_dev)
    ctr, labels = MeanShiftTorch(0.08).fit(A)
    labels[0] = 1
    assert A[labels, :].shape[1] == 3


def test_properties_at_full_size(cuda_dev):
    """size-independent properties at the size the bench uses (n_c ~ 3000):
    translation equivariance, permutation invariance of the centre, determinism."""
    rng = np.random.default_rng(1)
    n = 3072
    A = (np.array([0.1, -0.05, 0.8]) + rng.normal(0, 0.005, size=(n, 3))).astype(np.float32)
    out = rng.choice(n, n // 10, replace=False)
    A[out] = rng.uniform([-0.5, -0.4, 0.6], [0.5, 0.4, 1.2], size=(len(out), 3)).astype(np.float32)
    At = torch.from_numpy(A).to(cuda_dev)
    ms = MeanShiftTorch(0.08, mode="strict")
    c0, l0 = ms.fit(At)
    c0b, l0b = ms.fit(At)
    assert torch.equal(c0, c0b) and torch.equal(l0, l0b), "bitwise deterministic run to run"
    shift = torch.tensor([0.25, 0.125, -0.0625], device=cuda_dev)     # exactly representable
    c1, l1 = ms.fit(At + shift)
    assert float((c1 - shift - c0).norm() / c0.norm()) < REL_TOL
    perm = torch.from_numpy(rng.permutation(n)).to(cuda_dev)
    c2, l2 = ms.fit(At[perm])
    assert float((c2 - c0).norm() / c0.norm()) < REL_TOL
    assert int(l2.sum()) == int(l0.sum())
    # the centre is a fixed point: it sits inside the dense cluster
    assert float((c0.cpu() - torch.tensor([0.1, -0.05, 0.8])).norm()) < 0.003
    for mode in ("early_exit", "certified"):
        ce, le = MeanShiftTorch(0.08, mode=mode).fit(At)
        assert torch.equal(le, l0) and float((ce - c0).norm() / c0.norm()) < 1e-5, mode
    cd, ld = MeanShiftTorch(0.08).fit(At)            # the default mode is the certified one
    assert torch.equal(cd, ce) and torch.equal(ld, le)
    # dropping stationary seeds from the work lists changes neither T nor the centre
    msn = MeanShiftTorch(0.08, no_freeze=True)
    cn, ln = msn.fit(At)
    assert torch.equal(ln, l0) and float((cn - c0).norm() / c0.norm()) < 1e-6
    assert abs(int(msn.last_iters[0]) - int(ms.last_iters[0])) <= 1


def test_multi_tile_fit_and_max_iter_cap(cuda_dev):
    """n > 4096 points takes the streamed (multi-tile) sweep; max_iter caps the iteration count at
    max_iter + 1 like the reference loop (it > max_iter breaks after the sweep)."""
    rng = np.random.default_rng(2)
    n = 6000
    A = (np.array([0.0, 0.1, 0.9]) + rng.normal(0, 0.01, size=(n, 3))).astype(np.float32)
    out = rng.choice(n, n // 8, replace=False)
    A[out] = rng.uniform([-0.5, -0.4, 0.6], [0.5, 0.4, 1.2], size=(len(out), 3)).astype(np.float32)
    At = torch.from_numpy(A).to(cuda_dev)
    ms = MeanShiftTorch(0.08, mode="strict")
    c0, l0 = ms.fit(At)
    msn = MeanShiftTorch(0.08, no_freeze=True)
    cn, ln = msn.fit(At)
    assert torch.equal(l0, ln) and float((cn - c0).norm() / c0.norm()) < 1e-6
    assert abs(int(msn.last_iters[0]) - int(ms.last_iters[0])) <= 1
    assert float((c0.cpu() - torch.tensor([0.0, 0.1, 0.9])).norm()) < 0.005
    cc, lc = MeanShiftTorch(0.08, mode="certified").fit(At)
    assert torch.equal(lc, l0) and float((cc - c0).norm() / c0.norm()) < 1e-5
    capped = MeanShiftTorch(0.08, max_iter=4, mode="strict")
    capped.fit(At)
    assert int(capped.last_iters[0]) == 5


def test_cpu_tensor_is_rejected():
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        MeanShiftTorch(0.08).fit(torch.zeros(4, 3))


@pytest.mark.parametrize("kind", ["cluster", "two", "uniform", "coincident", "line", "tiny_bw", "huge_bw", "far_origin", "nan_inf"])
def test_pruned_density_equals_brute_force(cuda_dev, kind):
    """the exact pass that avoids n^2 tests (radial counting sort + triangle inequality) must give the same inlier
    count for EVERY input point as the brute-force pass, on vote-like and on adversarial clouds"""
    rng = np.random.default_rng(len(kind))
    n, bw = 3000, 0.08
    if kind == "cluster":
        A = np.array([0.1, -0.05, 0.8]) + rng.normal(0, 0.005, (n, 3)); A[:300] = rng.uniform([-0.5, -0.4, 0.6], [0.5, 0.4, 1.2], (300, 3))
    elif kind == "two":
        A = np.concatenate([np.array([0.0, 0.0, 0.8]) + rng.normal(0, 0.01, (n // 2, 3)), np.array([0.12, 0.0, 0.8]) + rng.normal(0, 0.01, (n // 2, 3))])
    elif kind == "uniform":
        A = rng.uniform(-0.3, 0.3, (n, 3))
    elif kind == "coincident":
        A = np.tile(np.array([[0.3, 0.2, 0.9]]), (n, 1)); A[::7] += 0.0799999; A[::11] += 0.0800001
    elif kind == "line":
        A = np.stack([np.linspace(0, 2.0, n), np.zeros(n), np.ones(n)], 1)       # spacing 0.67 mm: counts change every point
    elif kind == "tiny_bw":
        bw = 0.004; A = np.array([0.1, 0.0, 0.7]) + rng.normal(0, 0.005, (n, 3))
    elif kind == "huge_bw":
        bw = 5.0; A = rng.uniform(-1, 1, (n, 3))
    elif kind == "nan_inf":       # non-finite votes never count and are never counted (every test on them fails)
        A = np.array([0.1, -0.05, 0.8]) + rng.normal(0, 0.01, (n, 3))
        A[rng.choice(n, 7, replace=False)] = np.nan
        A[rng.choice(n, 5, replace=False), 1] = np.inf
        A[17, 2] = -np.inf
    else:
        A = np.array([120.0, -75.0, 300.0]) + rng.normal(0, 0.03, (n, 3))
    At = torch.from_numpy(A.astype(np.float32)).to(cuda_dev)
    out = {}
    for brute in (False, True):
        ms = MeanShiftTorch(bw, mode="certified")
        ms.brute_density = brute
        ctr, labels = ms.fit(At)
        out[brute] = (ms.last_counts()[:n].clone(), labels.clone(), ctr.clone())
    assert torch.equal(out[False][0], out[True][0]), int((out[False][0] != out[True][0]).sum())
    assert torch.equal(out[False][1], out[True][1])
    # sanity against a torch restatement on the GPU (its contraction order is not pinned: pairs exactly at the
    # threshold may differ, nothing else)
    d = At[:, None, :] - At[None, :, :]
    want = (torch.sqrt((d * d).sum(-1)) < torch.tensor(bw, dtype=torch.float32, device=cuda_dev)).sum(1).int()
    assert int((out[True][0] - want).abs().max()) <= 3
    if kind == "nan_inf":
        bad = ~torch.isfinite(At).all(1)
        assert int(out[False][0][bad].abs().max()) == 0 and not bool(out[False][1][bad].any())
