"""Deterministic synthetic weights for parity tests and benchmarks (test infrastructure).

There is no network for checkpoints, so every test/bench weight set is
generated from (name, shape, seed) alone: the same call yields the same
state_dict in the build container (where the golden fixtures are produced from
the real reference) and on the GPU box (where the CUDA path is checked).
Each tensor has its own generator seeded from crc32(name), so the values do
not depend on dict order or on which other tensors exist.
"""
from __future__ import annotations

import zlib
from typing import Dict, Mapping, Sequence

import torch


def _gen(name: str, seed: int) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed((seed * 1000003 + zlib.crc32(name.encode())) % (2 ** 62))
    return g


def synth_tensor(name: str, shape: Sequence[int], seed: int = 0) -> torch.Tensor:
    shape = tuple(int(s) for s in shape)
    g = _gen(name, seed)
    leaf = name.rsplit(".", 1)[-1]
    if leaf == "g":                                   # RMSNorm gains: near 1 but not all ones
        return 1.0 + 0.1 * torch.randn(shape, generator=g)
    if leaf == "mem_kv":                              # dd:163,207 randn init
        return torch.randn(shape, generator=g)
    if leaf == "bias":
        return 0.05 * torch.randn(shape, generator=g)
    if leaf == "weight" and len(shape) == 1:          # GroupNorm gains (VAE decoder): near 1 but not all ones
        return 1.0 + 0.1 * torch.randn(shape, generator=g)
    if leaf == "weight" and len(shape) >= 2:          # conv / linear: variance 1/fan_in
        fan_in = 1
        for s in shape[1:]:
            fan_in *= s
        bound = (3.0 / fan_in) ** 0.5
        return (torch.rand(shape, generator=g) * 2 - 1) * bound
    return torch.randn(shape, generator=g)


def synth_state_dict(shapes: Mapping[str, Sequence[int]], seed: int = 0) -> Dict[str, torch.Tensor]:
    """shapes: name -> shape (e.g. {k: v.shape for k, v in module.state_dict().items()})."""
    return {k: synth_tensor(k, s, seed) for k, s in shapes.items()}
