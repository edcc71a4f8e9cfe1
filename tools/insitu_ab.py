"""In-situ A/B of the two tensor-core GEMM kernels, per shape.

The conv2d launches of one UNet+ControlNet evaluation (batch 16, 256^2) are replayed in program order as ONE CUDA graph
(cold weights, activations as the network leaves them — what bench.py's roofline times).  Baseline: every launch on the
single-CTA kernel.  Then, one shape class at a time, that class alone is switched to the CTA-pair kernel (debug hook
mkd_debug_set_pair_auto, read by the host-side dispatch at capture time) and the whole list is timed again: the
difference is what the pair kernel buys for that class where it actually runs.  Back-to-back micro-benchmarks of one
launch (tools/gemm_bench.py) keep the weights in L2 and overstate it."""
import argparse
import ctypes as C
import sys

import torch

sys.path.insert(0, ".")
from makeupdiffuse_b200 import B200ControlLDM, _lib, ops  # noqa: E402
from makeupdiffuse_b200.synth import synthetic_batch, synthetic_state_dict  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--B", type=int, default=16)
ap.add_argument("--size", type=int, default=256)
ap.add_argument("--reps", type=int, default=20)
a = ap.parse_args()
dev = torch.device("cuda", 0)
lib = _lib.load()
lib.mkd_debug_set_pair_auto.argtypes = [C.c_int]
lib.mkd_debug_set_pair_auto.restype = None

m = B200ControlLDM(dtype=torch.bfloat16, device=dev)
m.load_state_dict(synthetic_state_dict(m, 0, dev))
m.concurrent = False
d = synthetic_batch(a.B, a.size, 768, device=dev)
cond = {"c_crossattn": [d["ctx"]], "c_concat": [torch.cat([d["src"], d["ref"]], 1)]}
t = torch.full((a.B,), 501, device=dev, dtype=torch.long)
m.apply_model(d["x_T"], t, cond)
torch.cuda.synchronize()
ops.PROFILE = []
m.apply_model(d["x_T"], t, cond)
torch.cuda.synchronize()
prof, ops.PROFILE = [r for r in ops.PROFILE if r["op"] == "conv2d" and r["path"] == _lib.PATH_TCGEN05], None


def key(r):
    ds = r["desc"]
    epi = ("e" if ds.emb else "") + ("r32" if ds.residual and ds.residual_dtype == _lib.MKD_F32 else "r16" if ds.residual else "") + \
          ("y" if ds.y else "") + ("Y" if ds.y32 else "") + ("s" if ds.stats else "") + (f"a{ds.act}" if ds.act else "")
    return (r["M"], r["K"], r["C"] * r["R"] * r["R"], r["R"], r["stride"], r["up"], epi)


keys = {}
for r in prof:
    keys.setdefault(key(r), []).append(r)
flops = sum(r["flops"] for r in prof)


def replay_ms(pair_keys):
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for cap in (False, True):
            ctx = torch.cuda.graph(g, stream=side) if cap else None
            if ctx:
                ctx.__enter__()
            for r in prof:
                lib.mkd_debug_set_pair_auto(1 if key(r) in pair_keys else 0)
                ops.run_conv_desc(r["desc"])
            if ctx:
                ctx.__exit__(None, None, None)
    torch.cuda.current_stream().wait_stream(side)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / a.reps


base = replay_ms(set())
base2 = replay_ms(set())
allp = replay_ms(set(keys))
print(f"{len(prof)} launches, {flops / 1e12:.3f} TFLOP: all single {base:.3f} ms (repeat {base2:.3f}), all pair-auto {allp:.3f} ms")
print("     M      N      K  R s u epi        count   delta_us_per_launch (pair - single; negative = pair faster)")
wins = []
for k, rs in sorted(keys.items(), key=lambda kv: -sum(r["flops"] for r in kv[1])):
    ms = replay_ms({k})
    delta = (ms - 0.5 * (base + base2)) * 1e3 / len(rs)
    print(f"{k[0]:6d} {k[1]:6d} {k[2]:6d}  {k[3]} {k[4]} {k[5]} {k[6]:10s} {len(rs):5d}   {delta:+8.2f}", flush=True)
    if delta < -0.3:
        wins.append(k)
best = replay_ms(set(wins))
print(f"pair on the {len(wins)} winning classes only: {best:.3f} ms  ({flops / best / 1e9:.1f} TFLOP/s)")
lib.mkd_debug_set_pair_auto(1)
