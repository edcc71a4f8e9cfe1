"""Multi-GPU path on CPU: world_size-2 gloo processes run the sharded sampler (makeupdiffuse_b200/dist.py) with the
C-ABI calls replaced by the torch fakes; the gathered result must equal the single-process result (batch rows are
independent; the only collective is one all-gather of the final latents)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))

from makeupdiffuse_b200.dist import shard_bounds  # noqa: E402

PARAMS = dict(model_channels=32, num_heads=2, context_dim=32)
BG, H, S = 4, 8, 4


def _setup_model():
    import fake_ops
    from makeupdiffuse_b200 import B200ControlLDM, B200DDIMSampler, ops
    from makeupdiffuse_b200.synth import synthetic_state_dict
    for name in fake_ops.ALL:
        setattr(ops, name, getattr(fake_ops, name))
    m = B200ControlLDM(PARAMS, PARAMS, dtype=torch.float32, device="cpu")
    m.load_state_dict(synthetic_state_dict(m, 0, device="cpu"))
    return B200DDIMSampler(m)


def _data():
    g = torch.Generator().manual_seed(7)
    return {"ctx": torch.randn(BG, 77, 32, generator=g), "hint": torch.rand(BG, 6, 8 * H, 8 * H, generator=g),
            "x_T": torch.randn(BG, 4, H, H, generator=g)}


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    from makeupdiffuse_b200.dist import sample_sharded
    s = _setup_model()
    d = _data()
    lo, hi = shard_bounds(BG, rank, world)
    cond = {"c_crossattn": [d["ctx"][lo:hi]], "c_concat": [d["hint"][lo:hi]]}
    with torch.no_grad():
        full = sample_sharded(s, S, BG, (4, H, H), cond, d["x_T"][lo:hi].contiguous(), rank, world)
    torch.save(full, os.path.join(out_dir, f"r{rank}.pt"))
    dist.destroy_process_group()


def test_shard_bounds():
    assert [shard_bounds(128, r, 8) for r in (0, 7)] == [(0, 16), (112, 128)]
    assert [shard_bounds(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert shard_bounds(5, 0, 1) == (0, 5)


def test_two_rank_gloo_equals_single_process(tmp_path):
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = torch.load(tmp_path / "r0.pt"), torch.load(tmp_path / "r1.pt")
    assert torch.equal(r0, r1) and r0.shape == (BG, 4, H, H)
    s = _setup_model()
    d = _data()
    cond = {"c_crossattn": [d["ctx"]], "c_concat": [d["hint"]]}
    with torch.no_grad():
        ref, _ = s.sample(S, BG, (4, H, H), cond, eta=0.0, x_T=d["x_T"], verbose=False)
    assert float((r0 - ref).abs().max()) < 1e-4 * float(ref.abs().max())
