"""Deterministic helpers shared by tests, golden-vector scripts and bench.py (no oracle imports)."""
from __future__ import annotations

import torch


def randomize_bn_(model: torch.nn.Module, seed: int = 1) -> torch.nn.Module:
    """Give every BatchNorm non-trivial affine + running statistics (a fresh model has mean 0 / var 1,
    which would hide BN-folding mistakes).  Same values for any module with the same BN layout."""
    g = torch.Generator().manual_seed(seed)
    for m in model.modules():
        if isinstance(m, torch.nn.modules.batchnorm._BatchNorm):
            c = m.num_features
            with torch.no_grad():
                m.weight.copy_(torch.rand(c, generator=g) + 0.5)
                m.bias.copy_(torch.randn(c, generator=g) * 0.1)
                m.running_mean.copy_(torch.randn(c, generator=g) * 0.1)
                m.running_var.copy_(torch.rand(c, generator=g) + 0.5)
    return model


def seeded_pointnet2msg(seed: int = 0, bn_seed: int = 1, input_channels: int = 6):
    """torch.manual_seed(seed); Pointnet2MSG() with default init; randomised BN; eval()."""
    from .pointnet2 import Pointnet2MSG

    torch.manual_seed(seed)
    model = Pointnet2MSG(input_channels=input_channels)
    randomize_bn_(model, bn_seed)
    return model.eval()
