"""Makeup interpolation sweep (BASELINE.json configs[4]; SURVEY.md §8(d) "Config 5").

The reference ships no interpolation code (README.md:25 and res/04_exp_interplote.png only show results), so the sweep
is defined through the unchanged ``apply_model`` surface: for reference k and blend weight w the ControlNet hint is

    c_concat = cat(source, (1 - w) * ref_k + w * ref_{(k+1) mod R})          (source first: makeup_diffuse.py:56)

which yields R * len(weights) independent samples of one source — an ordinary batch for the sampler, sharded across
ranks like any other (dist.sample_sharded).  Host-side tensor assembly only; all compute is the normal denoising path.
"""
from __future__ import annotations

import torch


def interpolation_hints(src: torch.Tensor, refs: torch.Tensor, weights) -> torch.Tensor:
    """src [1,3,H,W] or [3,H,W] in [0,1]; refs [R,3,H,W]; weights: iterable of floats in [0,1].
    Returns the hint batch [R * len(weights), 6, H, W], ordered reference-major (k, then w)."""
    if src.dim() == 3:
        src = src[None]
    if src.shape[0] != 1 or src.shape[1:] != refs.shape[1:]:
        raise ValueError(f"expected one source image matching the references, got {tuple(src.shape)} vs {tuple(refs.shape)}")
    w = torch.as_tensor(list(weights), dtype=refs.dtype, device=refs.device)
    if w.numel() == 0 or bool((w < 0).any()) or bool((w > 1).any()):
        raise ValueError("blend weights must be a non-empty sequence in [0, 1]")
    nxt = torch.roll(refs, shifts=-1, dims=0)
    blend = (1 - w)[None, :, None, None, None] * refs[:, None] + w[None, :, None, None, None] * nxt[:, None]  # [R, W, 3, H, W]
    blend = blend.reshape(-1, *refs.shape[1:])
    return torch.cat([src.expand(blend.shape[0], -1, -1, -1), blend], 1)


def interpolation_cond(src, refs, weights, context):
    """cond dict for the whole sweep; context [1 or B, 77, D] is the (constant) prompt encoding"""
    hint = interpolation_hints(src, refs, weights)
    ctx = context.expand(hint.shape[0], -1, -1).contiguous() if context.shape[0] == 1 else context
    if ctx.shape[0] != hint.shape[0]:
        raise ValueError("context batch must be 1 or R * len(weights)")
    return {"c_crossattn": [ctx], "c_concat": [hint]}
