"""Shared helpers for the parity tests."""
import numpy as np
import torch

import synthetic


def T(a, device="cpu"):
    return torch.from_numpy(np.ascontiguousarray(a)).to(device)


def rand_triple(seed0, N, C=64, F=128, device="cpu"):
    """The three torch.rand draws of one NeRFNetwork.forward, as tests/golden/make_golden.py's RandStream
    hands them out: call 0 -> [N,C], call 1 -> [N,1], call 2 -> [N,F,1]."""
    return (T(synthetic.uniforms(seed0, (N, C)), device), T(synthetic.uniforms(seed0 + 1, (N, 1)), device),
            T(synthetic.uniforms(seed0 + 2, (N, F, 1)), device))


def bits_equal(a, b):
    a = a.detach().cpu().contiguous().numpy() if isinstance(a, torch.Tensor) else np.ascontiguousarray(a)
    b = b.detach().cpu().contiguous().numpy() if isinstance(b, torch.Tensor) else np.ascontiguousarray(b)
    if a.shape != b.shape or a.dtype != b.dtype:
        return False
    if a.dtype.kind == "f":
        # exact value equality; +0 == -0 (torch.sum's zero sign is an implementation detail), NaN == NaN
        return bool(np.all((a == b) | (np.isnan(a) & np.isnan(b))))
    return bool(np.array_equal(a, b))
