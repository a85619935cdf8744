"""TEST INFRASTRUCTURE -- a CPU stand-in for the reference's `_ext` module backed by oracle/pn2.py.

Lets the UNMODIFIED reference Python (pointnet2_utils.py / pointnet2_modules.py / pvn3d.py
Pointnet2MSG) run end to end on CPU tensors in the build container, which is how the golden
feature vectors of tests/golden/pn2msg_*.npz were produced (tests/golden/make_golden_cpu.py).
Same nine names and argument orders as pvn3d/_ext-src/src/bindings.cpp:6-19.
"""
import numpy as np
import torch

from . import pn2


def _np(t):
    return t.detach().cpu().numpy()


def furthest_point_sampling(points, nsamples):
    return torch.from_numpy(pn2.furthest_point_sampling(_np(points), int(nsamples)))


def gather_points(points, idx):
    return torch.from_numpy(pn2.gather_points(_np(points), _np(idx)))


def gather_points_grad(grad_out, idx, n):
    return torch.from_numpy(pn2.gather_points_grad(_np(grad_out), _np(idx), int(n)))


def ball_query(new_xyz, xyz, radius, nsample):
    return torch.from_numpy(pn2.ball_query(_np(new_xyz), _np(xyz), float(np.float32(radius)), int(nsample)))


def group_points(points, idx):
    return torch.from_numpy(pn2.group_points(_np(points), _np(idx)))


def group_points_grad(grad_out, idx, n):
    return torch.from_numpy(pn2.group_points_grad(_np(grad_out), _np(idx), int(n)))


def three_nn(unknowns, knows):
    d, i = pn2.three_nn(_np(unknowns), _np(knows))
    return [torch.from_numpy(d), torch.from_numpy(i)]


def three_interpolate(points, idx, weight):
    return torch.from_numpy(pn2.three_interpolate(_np(points), _np(idx), _np(weight)))


def three_interpolate_grad(grad_out, idx, weight, m):
    return torch.from_numpy(pn2.three_interpolate_grad(_np(grad_out), _np(idx), _np(weight), int(m)))
