"""Synthetic (seeded, non-zero, device-independent) weights and inputs for benchmarks and smoke runs.

BASELINE.json asks for "random-init weights from base_diffusion_makeup.yaml"; upstream's own init zeroes the last layer
of every residual path (eps == 0, control == 0), which would make any benchmark and any parity check vacuous, so every
parameter is drawn from a counter-based hash instead (SURVEY.md §7 hard part 1).  The generator is the same function
the test oracle uses (tests assert equality), so both sides see identical weights without shipping a 4.9 GB checkpoint.
"""
from __future__ import annotations

import math
import zlib

import torch

_M64 = (1 << 64) - 1


def _i64(v: int) -> int:
    v &= _M64
    return v - (1 << 64) if v >= (1 << 63) else v


_GOLD, _C1, _C2 = _i64(0x9E3779B97F4A7C15), _i64(0xBF58476D1CE4E5B9), _i64(0x94D049BB133111EB)


def _lsr(z, s):
    return (z >> s) & ((1 << (64 - s)) - 1)


def hash_uniform(shape, seed: int, stream: int, device="cuda") -> torch.Tensor:
    """U[-1, 1) float32; element i depends only on (seed, stream, i): splitmix64 finaliser, top 24 bits"""
    n = int(math.prod(shape)) if len(shape) else 1
    base = _i64((seed * 0x9E3779B97F4A7C15) ^ (stream * 0xD1B54A32D192ED03))
    z = torch.arange(n, dtype=torch.int64, device=device) * _GOLD + base
    z = (z ^ _lsr(z, 30)) * _C1
    z = (z ^ _lsr(z, 27)) * _C2
    z = z ^ _lsr(z, 31)
    return (_lsr(z, 40).to(torch.float32) * (1.0 / 16777216.0) * 2.0 - 1.0).reshape(shape)


def synthetic_tensor(key: str, shape, seed: int = 0, device="cuda") -> torch.Tensor:
    """weights: std 1/sqrt(fan_in); norm gammas 1 +- 0.2; biases / betas +- 0.1"""
    if len(shape) >= 2:
        centre, half = 0.0, math.sqrt(3.0 / math.prod(shape[1:]))
    elif key.endswith("weight"):
        centre, half = 1.0, 0.2
    else:
        centre, half = 0.0, 0.1
    return hash_uniform(tuple(shape), seed, zlib.crc32(key.encode()), device) * half + centre


def synthetic_state_dict(ldm, seed: int = 0, device="cuda") -> dict:
    """upstream-keyed fp32 state dict for a B200ControlLDM (``control_model.*`` + ``model.diffusion_model.*``)"""
    sd = {}
    for prefix, net in (("control_model.", ldm.control_model), ("model.diffusion_model.", ldm.model.diffusion_model)):
        for k, shape in net.upstream_shapes().items():
            sd[prefix + k] = synthetic_tensor(prefix + k, shape, seed, device)
    return sd


def synthetic_first_stage_state_dict(decoder, seed: int = 0, device="cuda") -> dict:
    """upstream-keyed fp32 state dict (``first_stage_model.*``) for a B200FirstStageDecoder"""
    return {"first_stage_model." + k: synthetic_tensor("first_stage_model." + k, shape, seed, device)
            for k, shape in decoder.upstream_shapes().items()}


def synthetic_batch(B_global: int, image_hw: int = 256, context_dim: int = 768, seed: int = 1234, device="cuda"):
    """Synthetic source / reference pairs, text context and start noise for the WHOLE job, drawn from one seeded
    generator so that results do not depend on how many GPUs share the batch.
    src, ref ~ U[0,1) [B,3,H,W] (images in [0,1], diffdata/datasets.py:776-781); ctx ~ N(0,1) [B,77,D] stands in
    for CLIP('makeup transfer'); x_T ~ N(0,1) [B,4,H/8,W/8]."""
    g = torch.Generator(device=device).manual_seed(seed)
    h = image_hw // 8
    return {
        "src": torch.rand(B_global, 3, image_hw, image_hw, device=device, generator=g),
        "ref": torch.rand(B_global, 3, image_hw, image_hw, device=device, generator=g),
        "ctx": torch.randn(B_global, 77, context_dim, device=device, generator=g),
        "uc_ctx": torch.randn(B_global, 77, context_dim, device=device, generator=g),
        "x_T": torch.randn(B_global, 4, h, h, device=device, generator=g),
    }


def synthetic_clip_state_dict(cfg, seed=0, device="cpu"):
    """random-init weights of the CLIP text tower under HuggingFace's parameter names (benchmarks / smoke runs: no
    checkpoint is available offline); N(0, 0.02) matrices, small biases, LayerNorm weights around 1"""
    g = torch.Generator().manual_seed(seed)
    C, I, V, T = cfg["hidden_size"], cfg["intermediate_size"], cfg["vocab_size"], cfg["max_position_embeddings"]
    n = lambda *shape, std=0.02, mean=0.0: (torch.randn(*shape, generator=g) * std + mean).to(device)  # noqa: E731
    sd = {"text_model.embeddings.token_embedding.weight": n(V, C), "text_model.embeddings.position_embedding.weight": n(T, C),
          "text_model.final_layer_norm.weight": n(C, std=0.1, mean=1.0), "text_model.final_layer_norm.bias": n(C, std=0.05)}
    for i in range(cfg["num_hidden_layers"]):
        p = f"text_model.encoder.layers.{i}."
        for name in ("q_proj", "k_proj", "v_proj", "out_proj"):
            sd[p + f"self_attn.{name}.weight"], sd[p + f"self_attn.{name}.bias"] = n(C, C), n(C, std=0.05)
        sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"] = n(I, C), n(I, std=0.05)
        sd[p + "mlp.fc2.weight"], sd[p + "mlp.fc2.bias"] = n(C, I), n(C, std=0.05)
        for ln in ("layer_norm1", "layer_norm2"):
            sd[p + ln + ".weight"], sd[p + ln + ".bias"] = n(C, std=0.1, mean=1.0), n(C, std=0.05)
    return sd
