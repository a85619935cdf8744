"""PROTOTYPE (round-2 starting point, not wired into the library): ball query by cell list with the
reference's exact semantics.

The reference (`ball_query_gpu.cu:9-54`) scans the cloud in index order and keeps the first `nsample`
points with d2 < r^2, padding with the first hit.  Equivalent without the O(N) scan per centre:

  1. bin the cloud into voxels of edge >= r_max (x 1.001 -- far above the fp32 error of the binning even for grids of 1000 voxels: a neighbour is then always within +-1 voxel of
     the centre's voxel, whatever the rounding of the binning), hash the voxel to one of H buckets,
     counting-sort the point indices by bucket (stable: indices ascend inside a bucket);
  2. per centre: the <= 27 DISTINCT buckets of the neighbouring voxels give the candidates (hash
     collisions only add candidates -- the distance test is the reference's, bit for bit);
  3. hits are sorted by index; the first nsample are the reference's answer, the smallest one pads.

`d2` is the reference's contraction fma(dz,dz, fma(dx,dx, dy*dy)) in float32,
emulated exactly here (round-to-odd in float64).  tests/test_oracle_cpu.py checks this against the oracle.
A GPU version: warp per centre, candidates appended by ballot into a shared-memory list, bitonic
selection of the nsample smallest indices; fall back to the index-order scan when a ball holds more
than the list (dense clouds, where the scan exits early anyway).
"""
import numpy as np


def _fma32(a, b, c):
    """correctly rounded float32 fma(a, b, c) for float32 arrays (product exact in float64, sum rounded to odd)"""
    a = a.astype(np.float64); b = b.astype(np.float64); c = c.astype(np.float64)
    s = a * b                                   # exact: 24 x 24 bits
    t = s + c
    bb = t - s
    e = (s - (t - bb)) + (c - bb)               # TwoSum: t + e == s + c exactly
    bits = t.view(np.int64)
    odd = (bits & 1) == 1
    fix = (e != 0) & ~odd & np.isfinite(t)
    toward = np.where(e > 0, np.inf, -np.inf)
    t = np.where(fix, np.nextafter(t, toward), t)
    return t.astype(np.float32)


def ref_d2(c, p):
    """squared distance exactly as the reference kernels compute it (float32, contraction order of the SASS)"""
    dx = (c[..., 0] - p[..., 0]).astype(np.float32)
    dy = (c[..., 1] - p[..., 1]).astype(np.float32)
    dz = (c[..., 2] - p[..., 2]).astype(np.float32)
    return _fma32(dz, dz, _fma32(dx, dx, (dy * dy).astype(np.float32)))


def build_cells(xyz, r_max, n_buckets=4096):
    cell = np.float64(r_max) * 1.001
    lo = xyz.min(0).astype(np.float64)
    ijk = np.floor((xyz.astype(np.float64) - lo) / cell).astype(np.int64)
    h = _hash(ijk, n_buckets)
    order = np.argsort(h, kind="stable").astype(np.int32)           # counting sort on the GPU
    start = np.searchsorted(h[order], np.arange(n_buckets + 1)).astype(np.int32)
    return dict(cell=cell, lo=lo, order=order, start=start, n_buckets=n_buckets)


def _hash(ijk, n_buckets):
    return ((ijk[..., 0] * 73856093) ^ (ijk[..., 1] * 19349663) ^ (ijk[..., 2] * 83492791)) % n_buckets


def ball_query_cells(new_xyz, xyz, radii, nsamples, n_buckets=4096):
    """one cloud: new_xyz [M,3], xyz [N,3] float32 -> [idx[M,ns] int32 for every (radius, nsample)]"""
    cells = build_cells(xyz, max(radii), n_buckets)
    outs = [np.zeros((new_xyz.shape[0], ns), np.int32) for ns in nsamples]
    r2 = [np.float32(np.float32(r) * np.float32(r)) for r in radii]        # ball_query_gpu.cu:22
    off = np.array([(i, j, k) for i in (-1, 0, 1) for j in (-1, 0, 1) for k in (-1, 0, 1)], np.int64)
    for j, c in enumerate(new_xyz):
        ic = np.floor((c.astype(np.float64) - cells["lo"]) / cells["cell"]).astype(np.int64)
        buckets = np.unique(_hash(ic[None, :] + off, cells["n_buckets"]))
        cand = np.concatenate([cells["order"][cells["start"][b]:cells["start"][b + 1]] for b in buckets])
        if cand.size == 0:
            continue
        d2 = ref_d2(c[None, :], xyz[cand])
        for out, rr, ns in zip(outs, r2, nsamples):
            hits = np.sort(cand[d2 < rr])
            if hits.size:
                k = min(ns, hits.size)
                out[j, :k] = hits[:k]
                out[j, k:] = hits[0]
    return outs
