"""Per-role %globaltimer timeline of one CTA-pair GEMM launch (debug hook mkd_debug_set_trace).
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
NAMES = ["entry", "prologue done", "tma0 issued", "full0 (mma)", "mma u0 issued", "acc0 ready (epi)", "panel0 computed",
         "store0 issued", "last store issued", "stores drained", "last panel computed", "exit", "panel0 tmem loaded",
         "panel0 written", "cluster sync 1", "tma0 about to issue"]
ORDER = [0, 14, 1, 15, 2, 3, 4, 5, 12, 13, 6, 7, 10, 8, 9, 11]
SHAPES = {"sq320_plain": (16, 32, 32, 320, 320, 1, "plain"), "sq320_res32": (16, 32, 32, 320, 320, 1, "res32"),
          "qkv_320": (16, 32, 32, 320, 960, 1, "plain"), "conv320": (16, 32, 32, 320, 320, 3, "emb32"),
          "sq1280_plain": (16, 8, 8, 1280, 1280, 1, "plain"), "conv1280": (16, 8, 8, 1280, 1280, 3, "emb32"),
          "ff1_320": (16, 32, 32, 320, 2560, 1, "geglu"),
          "conv1280_4x4": (16, 4, 4, 1280, 1280, 3, "emb32"), "sq1280_m256": (16, 4, 4, 1280, 1280, 1, "plain"),
          "conv2560_4x4": (16, 4, 4, 2560, 1280, 3, "emb32"), "ff2_1280": (16, 8, 8, 5120, 1280, 1, "res32")}
only = sys.argv[1] if len(sys.argv) > 1 else ""
ws = torch.empty(96 << 20, dtype=torch.uint8, device=DEV)
for name, (N, H, W, C, K, R, epi) in SHAPES.items():
    if only and only not in name:
        continue
    M = N * H * W
    x = torch.randn(M, C, device=DEV).bfloat16()
    w = (torch.randn(K, R, R, C, device=DEV) / math.sqrt(C * R * R)).bfloat16()
    kw = dict(N=N, H=H, W=W, R=R, S=R, pad=R // 2, bias=torch.randn(K, device=DEV), workspace=ws, path=L.PATH_TCGEN05_PAIR)
    y = torch.empty(M, K // 2 if epi == "geglu" else K, device=DEV, dtype=torch.bfloat16)
    if epi == "res32":
        r = torch.randn(M, K, device=DEV)
        y = None
        kw.update(residual=r, y32=r)
    elif epi == "emb32":
        y = None
        kw.update(emb=torch.randn(N, K, device=DEV).bfloat16(), y32=torch.empty(M, K, device=DEV))
    elif epi == "geglu":
        kw.update(act=L.ACT_GEGLU, geglu_block=128)
    d = ops.make_conv_desc(x, w, y, **kw)
    for _ in range(3):
        ops.run_conv_desc(d)
    tr = torch.zeros(148 * 16 + 148 * 16 * 8, dtype=torch.int64, device=DEV)
    torch.cuda.synchronize()
    lib.mkd_debug_set_trace(tr.data_ptr())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ops.run_conv_desc(d)
    e1.record()
    torch.cuda.synchronize()
    lib.mkd_debug_set_trace(None)
    t = tr.cpu()[:148 * 16].reshape(148, 16)
    det = tr.cpu()[148 * 16:].reshape(148, 16, 8)
    used = t[:, 0] > 0
    t0 = int(t[used, 0].min())
    print(f"== {name}: event time {1e3 * e0.elapsed_time(e1):.1f} us, CTAs {int(used.sum())}; times in us since first CTA entry")
    for slot in ORDER:
        col = t[used, slot].double()
        col = col[col > 0]
        if len(col):
            print(f"   {NAMES[slot]:20s} min {(col.min() - t0) / 1e3:7.2f}  median {(col.median() - t0) / 1e3:7.2f}  max {(col.max() - t0) / 1e3:7.2f}   ({len(col)} CTAs)")
    # epilogue detail of CTA 0 and 1: clock64 stamps per panel (warp 4 / 8 lane 0), relative to the first
    for cta in (0, 1, 2):
        dd = det[cta]
        base = int(dd[dd > 0].min()) if bool((dd > 0).any()) else 0
        print(f"   CTA {cta} panels (clk since first stamp): start, slot free, tmem data, written, fenced, arrived")
        for gidx in range(16):
            if int(dd[gidx, 0]) > 0:
                print("      panel %2d: " % gidx + " ".join("%7d" % (int(v) - base) if int(v) > 0 else "      -" for v in dd[gidx, :6]))
