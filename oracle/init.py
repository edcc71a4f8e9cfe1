"""Seeded, device-independent, NON-ZERO parameter initialisation shared by the oracle and the B200 path.

Why not the upstream initialisation: upstream wraps the last layer of every residual path in ``zero_module``
(ResBlock out-conv, SpatialTransformer proj_out, UNet ``out`` conv, all 13 ControlNet zero-convs, the last
hint conv), so "random-init weights from the yaml" gives eps == 0 and control == 0 and parity would pass
vacuously (SURVEY.md §7 hard part 1).  Every parameter is therefore drawn from a counter-based hash
(splitmix64 finaliser over ``seed``, a CRC of the state-dict key and the flat element index) mapped to a
fan-in-scaled uniform.  Integer torch ops wrap identically on CPU and CUDA, so the values are bit-identical
on every device and need no 4.9 GB checkpoint to travel.

Test infrastructure (see oracle/__init__.py).
"""
from __future__ import annotations

import math
import zlib

import torch

_M64 = (1 << 64) - 1


def _i64(v: int) -> int:
    v &= _M64
    return v - (1 << 64) if v >= (1 << 63) else v


_GOLD = _i64(0x9E3779B97F4A7C15)
_C1 = _i64(0xBF58476D1CE4E5B9)
_C2 = _i64(0x94D049BB133111EB)


def _lsr(z: torch.Tensor, s: int) -> torch.Tensor:
    """logical shift right on int64 (torch's >> is arithmetic)"""
    return (z >> s) & ((1 << (64 - s)) - 1)


def hash_uniform(shape, seed: int, stream: int, device="cpu") -> torch.Tensor:
    """float32 tensor of U[-1, 1) values; element i = f(seed, stream, i), identical on every device."""
    n = int(math.prod(shape)) if len(shape) else 1
    base = _i64((seed * 0x9E3779B97F4A7C15) ^ (stream * 0xD1B54A32D192ED03))
    z = torch.arange(n, dtype=torch.int64, device=device) * _GOLD + base
    z = (z ^ _lsr(z, 30)) * _C1
    z = (z ^ _lsr(z, 27)) * _C2
    z = z ^ _lsr(z, 31)
    f = _lsr(z, 40).to(torch.float32) * (1.0 / 16777216.0)  # 24 exact bits in [0,1)
    return (f * 2.0 - 1.0).reshape(shape)


def init_rule(key: str, shape) -> tuple[float, float]:
    """(centre, half-width) of the uniform a parameter is drawn from.

    weights (ndim >= 2): std = 1/sqrt(fan_in)  -> half-width sqrt(3/fan_in)   (keeps activations O(1))
    norm gammas (1-D ``weight``): 1 +- 0.2 ;  all biases / norm betas: +- 0.1
    """
    if len(shape) >= 2:
        fan_in = math.prod(shape[1:])
        return 0.0, math.sqrt(3.0 / fan_in)
    if key.endswith("weight"):
        return 1.0, 0.2
    return 0.0, 0.1


@torch.no_grad()
def seeded_state_dict(module: torch.nn.Module, seed: int = 0, prefix: str = "", device=None) -> dict:
    """Fill ``module``'s parameters in place and return {prefix+key: tensor}."""
    out = {}
    for key, p in module.named_parameters():
        full = prefix + key
        centre, half = init_rule(full, tuple(p.shape))
        dev = p.device if device is None else device
        u = hash_uniform(tuple(p.shape), seed, zlib.crc32(full.encode()), device=dev)
        p.copy_((u * half + centre).to(p.dtype))
        out[full] = p
    return out
