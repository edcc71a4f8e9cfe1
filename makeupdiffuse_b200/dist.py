"""Batch sharding across the GPUs of one box.

Every (source, reference, noise) sample denoises independently — nothing in diffmk/cddim.py:9-100 or apply_model mixes
batch rows — so ranks take contiguous batch chunks, weights are replicated (re-generated / loaded per rank) and the
only data-path collective is ONE all-gather of the final latents over NVLink 5 / NVSwitch, done in one of two ways:

* NCCL (default): the last DDIM update kernel writes x_0 directly into this rank's slice of the gather buffer
  (``out=``), and ``all_gather_into_tensor`` runs in place on it — no staging copy.
* fused (``fused_gather=True``): the gather buffer is torch symmetric memory (every rank's copy is mapped into every
  other rank); the last DDIM update kernel stores its x_0 slice into ALL ranks' buffers itself
  (``mkd_ddim_update_peers``: plain stores to peer addresses over NVLink), and a signal barrier between the ranks
  replaces the collective launch.  Same bits as the NCCL path (tests/test_sharding.py, 2 GPUs).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

_symm_cache: dict = {}


def shard_bounds(n_global: int, rank: int, world: int):
    """contiguous chunk [lo, hi) of rank; the first n_global % world ranks take one extra sample"""
    base, extra = divmod(n_global, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def _symmetric_gather_buffer(shape, device, group):
    """persistent symmetric-memory gather buffer + its rendezvous handle for this shape (collective on first use)"""
    import torch.distributed._symmetric_memory as symm_mem
    group = group or dist.group.WORLD
    key = (tuple(shape), str(device), group.group_name)
    hit = _symm_cache.get(key)
    if hit is None:
        buf = symm_mem.empty(shape, dtype=torch.float32, device=device)
        hit = _symm_cache[key] = (buf, symm_mem.rendezvous(buf, group.group_name))
    return hit


def sample_sharded(sampler, S, batch_global, shape, cond_local, x_T_local, rank=0, world=1, group=None,
                   fused_gather=False, **kw):
    """DDIM-sample this rank's chunk and all-gather the final latents.

    cond_local / x_T_local hold this rank's rows only.  Returns the [batch_global, C, H, W] latents on every rank.
    Requires equal chunk sizes when world > 1."""
    lo, hi = shard_bounds(batch_global, rank, world)
    b = hi - lo
    C, H, W = shape
    if world > 1 and batch_global % world:
        raise ValueError("sample_sharded needs batch_global % world == 0")
    fused = bool(fused_gather) and world > 1
    hdl = peer_ptrs = None
    if fused:
        gathered, hdl = _symmetric_gather_buffer((batch_global, C, H, W), x_T_local.device, group)
        off = lo * C * H * W * 4
        peer_ptrs = [int(p) + off for p in hdl.buffer_ptrs]  # this rank's slice inside every rank's buffer
        hdl.barrier()  # nobody still reads the previous result out of the buffers we are about to overwrite
    else:
        gathered = torch.empty(batch_global, C, H, W, dtype=torch.float32, device=x_T_local.device)
    mine = gathered[lo:hi]
    sampler.make_schedule(ddim_num_steps=S, ddim_eta=kw.pop("eta", 0.0), verbose=False)
    steps = np.flip(sampler.ddim_timesteps)
    own = hasattr(sampler, "begin_loop")
    if own:
        sampler.begin_loop(steps)  # this loop reads its cond afresh, timestep embeddings of all steps at once (B200DDIMSampler)
        kw = dict(kw)
    x = x_T_local
    for i, step in enumerate(steps):
        ts = torch.full((b,), int(step), device=x.device, dtype=torch.long)
        last = i == len(steps) - 1
        if own:
            kw["t_value"] = int(step)
        x, _ = sampler.denoising_step(x, cond_local, ts, index=len(steps) - i - 1, out=mine if last else None,
                                      peer_ptrs=peer_ptrs if last else None, **kw)
    if fused:
        hdl.barrier()  # every rank's final update kernel has stored its slice everywhere
        return gathered.clone()  # the symmetric buffer is reused by the next call
    if world > 1:
        dist.all_gather_into_tensor(gathered, mine, group=group)  # in place: input is the rank's own slice
    return gathered
