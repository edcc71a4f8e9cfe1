"""A few eager UNet+ControlNet evaluations at the bench configuration — the short command ncu wraps."""
import argparse
import sys

import torch

sys.path.insert(0, ".")
from makeupdiffuse_b200 import B200ControlLDM, _lib  # noqa: E402
from makeupdiffuse_b200.synth import synthetic_batch, synthetic_state_dict  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--B", type=int, default=16)
ap.add_argument("--size", type=int, default=256)
ap.add_argument("--evals", type=int, default=2)
a = ap.parse_args()
dev = torch.device("cuda", 0)
m = B200ControlLDM(dtype=torch.bfloat16, device=dev)
m.load_state_dict(synthetic_state_dict(m, 0, dev))
d = synthetic_batch(a.B, a.size, 768, device=dev)
cond = {"c_crossattn": [d["ctx"]], "c_concat": [torch.cat([d["src"], d["ref"]], 1)]}
t = torch.full((a.B,), 501, device=dev, dtype=torch.long)
lib = _lib.load()
m.precompute_time_embeddings([501])  # as inside the sampler's loop: embeddings of the loop's timesteps from the table
m.set_step(501, a.B)
for i in range(a.evals):
    if i == a.evals - 1:
        torch.cuda.profiler.start()  # `ncu --profile-from-start off` then sees exactly one warm evaluation
    n0 = lib.mkd_launch_count()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    eps = m.apply_model(d["x_T"], t, cond)
    e1.record()
    torch.cuda.synchronize()
    print(f"eval {i}: {lib.mkd_launch_count() - n0} library launches, {e0.elapsed_time(e1):.2f} ms (eager, host-bound)", flush=True)
torch.cuda.profiler.stop()
print("eps", float(eps.float().std()))
