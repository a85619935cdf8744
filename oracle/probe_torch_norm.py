"""TEST INFRASTRUCTURE -- probe which fp32 contraction CPU torch.norm(dim=-1) uses for 3-vectors.

The exact outputs of MeanShiftTorch.fit (labels, num_in, argmax; reference
pvn3d/lib/utils/meanshift_pytorch.py:46-51) are threshold tests on torch.norm of [n,n,3] float32
tensors.  This script checks candidate formulas bit-for-bit.
Run: python oracle/probe_torch_norm.py
"""
import numpy as np
import torch


def fma32(a, b, c):
    # exact fp32 fma through float64 (products of two fp32 are exact in fp64; one rounding of the
    # sum to fp64 then to fp32 can double-round, so use integer-exact path via np.longdouble)
    return (a.astype(np.longdouble) * b.astype(np.longdouble) + c.astype(np.longdouble)).astype(np.float32)


def main():
    g = torch.Generator().manual_seed(0)
    n = 1500
    a = (torch.rand(n, 3, generator=g) - 0.5) * 0.4
    a[:, 2] += 0.8
    diff = a.view(1, n, 3) - a.view(n, 1, 3)          # [n,n,3]
    ref = torch.norm(diff, dim=2).numpy()
    d = diff.numpy()
    x, y, z = d[..., 0], d[..., 1], d[..., 2]
    cands = {
        "sqrt(fma(z,z,fma(y,y,x*x)))": np.sqrt(fma32(z, z, fma32(y, y, (x * x).astype(np.float32)))),
        "sqrt(fma(z,z,fma(x,x,y*y)))": np.sqrt(fma32(z, z, fma32(x, x, (y * y).astype(np.float32)))),
        "sqrt((x*x+y*y)+z*z) plain": np.sqrt(((x * x + y * y) + z * z).astype(np.float32)),
        "sqrt(f64 sum)": np.sqrt((x.astype(np.float64) ** 2 + y.astype(np.float64) ** 2 + z.astype(np.float64) ** 2)).astype(np.float32),
    }
    for k, v in cands.items():
        mism = int((v.astype(np.float32) != ref).sum())
        print(f"{k:36s} mismatches {mism} / {ref.size}")
    for nt in (1, 4, 8):
        torch.set_num_threads(nt)
        r2 = torch.norm(diff, dim=2).numpy()
        print("threads", nt, "identical", bool((r2 == ref).all()))
    v = torch.linalg.vector_norm(diff, dim=2).numpy()
    print("linalg.vector_norm identical", bool((v == ref).all()))


if __name__ == "__main__":
    main()
