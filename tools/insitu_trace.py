"""%globaltimer timeline of the tensor-core GEMM launches INSIDE one eager, single-stream UNet+ControlNet evaluation.

Needs the stamps compiled in:  MKD_TRACE=1 python -m makeupdiffuse_b200.build --force
Every conv2d call gets its own trace buffer (the pointer travels in the launch parameters), so no synchronisation is
added between launches.  Per launch: when its CTAs entered, finished the prologue, saw the first full stage, and left,
relative to the first entry; plus the distance from the previous GEMM launch's last exit to this one's first entry."""
import argparse
import ctypes
import sys

import torch

sys.path.insert(0, ".")
from makeupdiffuse_b200 import B200ControlLDM, _lib, ops  # noqa: E402
from makeupdiffuse_b200.synth import synthetic_batch, synthetic_state_dict  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--B", type=int, default=16)
ap.add_argument("--size", type=int, default=256)
ap.add_argument("--first", type=int, default=0)
ap.add_argument("--count", type=int, default=400)
a = ap.parse_args()
dev = torch.device("cuda", 0)
lib = _lib.load()
lib.mkd_debug_set_trace.argtypes = [ctypes.c_void_p]
WORDS = 148 * 16 + 148 * 16 * 8

m = B200ControlLDM(dtype=torch.bfloat16, device=dev)
m.load_state_dict(synthetic_state_dict(m, 0, dev))
m.concurrent = False
d = synthetic_batch(a.B, a.size, 768, device=dev)
cond = {"c_crossattn": [d["ctx"]], "c_concat": [torch.cat([d["src"], d["ref"]], 1)]}
t = torch.full((a.B,), 501, device=dev, dtype=torch.long)
for _ in range(2):
    m.apply_model(d["x_T"], t, cond)
torch.cuda.synchronize()

orig = ops.conv2d
recs = []
bufs = torch.zeros(a.count, WORDS, dtype=torch.int64, device=dev)
n = {"i": 0}


def traced(x2d, w, y2d, **kw):
    i = n["i"]
    n["i"] += 1
    if a.first <= i < a.first + a.count:
        lib.mkd_debug_set_trace(bufs[i - a.first].data_ptr())
        orig(x2d, w, y2d, **kw)
        lib.mkd_debug_set_trace(None)
        R = kw.get("R", 1)
        epi = ("e" if kw.get("emb") is not None else "") + ("r" if kw.get("residual") is not None else "") + \
              ("y" if y2d is not None else "") + ("Y" if kw.get("y32") is not None else "") + ("s" if kw.get("stats") is not None else "") + \
              (f"a{kw['act']}" if kw.get("act") else "")
        recs.append((i, kw["N"] * kw["H"] * kw["W"], w.shape[0], x2d.shape[1] * R * R, kw.get("stride", 1), int(bool(kw.get("upsample"))), epi))
    else:
        orig(x2d, w, y2d, **kw)


ops.conv2d = traced
torch.cuda._sleep(int(0.05 * 1.9e9))  # let the host run ahead: launches queue back to back
m.apply_model(d["x_T"], t, cond)
torch.cuda.synchronize()
ops.conv2d = orig
tr = bufs.cpu()
print("  #      M      N      K s u epi     kern  CTAs | gap  | entry_spread prologue  full0  | first_exit median_exit last_exit (us since first entry)")
prev_exit = None
tot = 0.0
for j, (i, M, N, K, s, u, epi) in enumerate(recs):
    tt = tr[j, :148 * 16].reshape(148, 16)
    used = tt[:, 0] > 0
    if not bool(used.any()):
        print(f"{i:3d} {M:6d} {N:6d} {K:6d} {s} {u} {epi:7s} (not a traced kernel)")
        prev_exit = None
        continue
    pair = bool((tt[used, 11] > 0).any())
    ex = tt[used, 11 if pair else 8].double()
    t0 = float(tt[used, 0].min())
    f = lambda col: (float(col[col > 0].median()) - t0) / 1e3 if bool((col > 0).any()) else float("nan")  # noqa: E731
    gap = (t0 - prev_exit) / 1e3 if prev_exit is not None else float("nan")
    prev_exit = float(ex.max())
    tot += (float(ex.max()) - t0) / 1e3
    print(f"{i:3d} {M:6d} {N:6d} {K:6d} {s} {u} {epi:7s} {'pair' if pair else 'sngl'} {int(used.sum()):5d} | {gap:5.1f} | "
          f"{(float(tt[used, 0].max()) - t0) / 1e3:6.2f} {f(tt[used, 1].double()):8.2f} {f(tt[used, 3].double()):7.2f} | "
          f"{(float(ex.min()) - t0) / 1e3:8.2f} {(float(ex.median()) - t0) / 1e3:8.2f} {(float(ex.max()) - t0) / 1e3:8.2f}"
          + (f" | mma_u0 {f(tt[used, 4].double()):6.2f} acc0 {f(tt[used, 5].double()):6.2f} panel0 {f(tt[used, 6].double()):6.2f} "
             f"lastpanel {f(tt[used, 10].double()):6.2f} drained {f(tt[used, 9].double()):6.2f}" if pair else
             f" | mma_u0 {f(tt[used, 4].double()):6.2f} acc0 {f(tt[used, 5].double()):6.2f} epi0 {f(tt[used, 6].double()):6.2f} epilast {f(tt[used, 7].double()):6.2f}"))
print(f"sum of (last exit - first entry) over {len(recs)} launches: {tot / 1e3:.3f} ms")
