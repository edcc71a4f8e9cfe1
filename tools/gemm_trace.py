"""Per-role %globaltimer timeline of one tcgen05 GEMM launch (debug hook mkd_debug_set_trace).
Needs a library built with the stamps compiled in:  MKD_TRACE=1 python -m makeupdiffuse_b200.build --force"""
import ctypes
import math
import sys

import torch

sys.path.insert(0, ".")
from makeupdiffuse_b200 import _lib as L  # noqa: E402
from makeupdiffuse_b200 import ops  # noqa: E402

lib = L.load()
lib.mkd_debug_set_trace.argtypes = [ctypes.c_void_p]
DEV = "cuda"
NAMES = ["entry", "prologue", "tma0 issued", "full0", "mma u0 done", "acc0 ready", "epi u0 done", "epi last done", "exit",
         "p0 bar1", "p0 tmem ld", "p0 staged", "p0 phase2", "p1 staged"]
ORDER = [0, 1, 2, 3, 4, 5, 9, 10, 11, 12, 13, 6, 7, 8]
SHAPES = {"sq320": (16, 32, 32, 320, 320, 1), "conv320": (16, 32, 32, 320, 320, 3),
          "tiny": (1, 1, 128, 64, 160, 1), "sq1280_m256": (16, 4, 4, 1280, 1280, 1)}
if len(sys.argv) > 1 and sys.argv[1] == "dual":  # statistics-emitting 3x3 convs (DUAL-eligible: run with MKD_DUAL=1 / 0)
    SHAPES = {"conv320_stats": (16, 32, 32, 320, 320, 3), "conv960_stats": (16, 32, 32, 960, 320, 3)}
if len(sys.argv) > 1 and sys.argv[1] == "nsweep":  # same A, K = 2880; N-tile = 160 / 80 / 64 / 32 (BN picked from N)
    SHAPES = {"N320_bn160": (16, 32, 32, 320, 320, 3), "N240_bn80": (16, 32, 32, 320, 240, 3),
              "N256_bn64": (16, 32, 32, 320, 256, 3), "N32_bn32": (16, 32, 32, 320, 32, 3)}
for name, (N, H, W, C, K, R) in SHAPES.items():
    M = N * H * W
    x = torch.randn(M, C, device=DEV).bfloat16()
    w = (torch.randn(K, R, R, C, device=DEV) / math.sqrt(C * R * R)).bfloat16()
    y = torch.empty(M, K, device=DEV, dtype=torch.bfloat16)
    extra = {}
    if name.endswith("_stats"):
        extra = dict(stats=torch.empty(M // 128, K, 2, device=DEV), y32=torch.empty(M, K, device=DEV),
                     emb=torch.randn(N, K, device=DEV).bfloat16())
    d = ops.make_conv_desc(x, w, y, N=N, H=H, W=W, R=R, S=R, pad=R // 2, bias=torch.randn(K, device=DEV), **extra)
    for _ in range(3):
        ops.run_conv_desc(d)
    tr = torch.zeros(148 * 16, dtype=torch.int64, device=DEV)
    torch.cuda.synchronize()
    lib.mkd_debug_set_trace(tr.data_ptr())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ops.run_conv_desc(d)
    e1.record()
    torch.cuda.synchronize()
    lib.mkd_debug_set_trace(None)
    t = tr.cpu().reshape(148, 16)[:, :14]
    used = t[:, 0] > 0
    t0 = int(t[used, 0].min())
    kbl = C * R * R // 64
    mm = t[used][:, 4].double() - t[used][:, 3].double()
    print(f"== {name}: k-blocks/unit {kbl}, main loop of unit 0: median {float(mm.median()) / 1e3:.2f} us -> {float(mm.median()) / kbl:.0f} ns per k-block")
    print(f"== {name}: event time {1e3 * e0.elapsed_time(e1):.1f} us, CTAs {int(used.sum())}; times in us since first CTA entry")
    for slot in ORDER:
        nm = NAMES[slot]
        col = t[used, slot].double()
        col = col[col > 0]
        if len(col):
            print(f"   {nm:14s} min {(col.min() - t0) / 1e3:7.2f}  median {(col.median() - t0) / 1e3:7.2f}  max {(col.max() - t0) / 1e3:7.2f}")
