"""Batch sharding across the GPUs of one box.

Every (source, reference, noise) sample denoises independently — nothing in diffmk/cddim.py:9-100 or apply_model mixes
batch rows — so ranks take contiguous batch chunks, weights are replicated (re-generated / loaded per rank) and the
only data-path collective is ONE all-gather of the final latents over NCCL (NVLink 5 / NVSwitch).  The last DDIM update
kernel writes x_0 directly into this rank's slice of the gather buffer (``out=``), so the collective runs in place with
no staging copy.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def shard_bounds(n_global: int, rank: int, world: int):
    """contiguous chunk [lo, hi) of rank; the first n_global % world ranks take one extra sample"""
    base, extra = divmod(n_global, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def sample_sharded(sampler, S, batch_global, shape, cond_local, x_T_local, rank=0, world=1, group=None, **kw):
    """DDIM-sample this rank's chunk and all-gather the final latents.

    cond_local / x_T_local hold this rank's rows only.  Returns the [batch_global, C, H, W] latents on every rank.
    Requires equal chunk sizes when world > 1 (all_gather_into_tensor)."""
    lo, hi = shard_bounds(batch_global, rank, world)
    b = hi - lo
    C, H, W = shape
    if world > 1 and batch_global % world:
        raise ValueError("sample_sharded needs batch_global % world == 0")
    gathered = torch.empty(batch_global, C, H, W, dtype=torch.float32, device=x_T_local.device)
    mine = gathered[lo:hi]
    sampler.make_schedule(ddim_num_steps=S, ddim_eta=kw.pop("eta", 0.0), verbose=False)
    steps = np.flip(sampler.ddim_timesteps)
    x = x_T_local
    for i, step in enumerate(steps):
        ts = torch.full((b,), int(step), device=x.device, dtype=torch.long)
        last = i == len(steps) - 1
        x, _ = sampler.denoising_step(x, cond_local, ts, index=len(steps) - i - 1, out=mine if last else None, **kw)
    if world > 1:
        dist.all_gather_into_tensor(gathered, mine, group=group)  # in place: input is the rank's own slice
    return gathered
