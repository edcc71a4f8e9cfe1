"""Micro-benchmark of the tcgen05 conv/GEMM kernel on the shapes that dominate one UNet+ControlNet eval.
Back-to-back launches replayed from a CUDA graph between two events (no host gaps), L2 warm or flushed (--flush)."""
import argparse
import math
import sys

import torch

sys.path.insert(0, ".")
from makeupdiffuse_b200 import _lib as L  # noqa: E402
from makeupdiffuse_b200 import ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--only", default="")
ap.add_argument("--flush", action="store_true")
a = ap.parse_args()
DEV = "cuda"
ws = torch.empty(96 << 20, dtype=torch.uint8, device=DEV)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)

SHAPES = {
    # name: (N, H, W, C, K, R, epilogue)
    "sq320_res32": (16, 32, 32, 320, 320, 1, "res32"),
    "sq320_plain": (16, 32, 32, 320, 320, 1, "plain"),
    "sq640_res32": (16, 16, 16, 640, 640, 1, "res32"),
    "sq1280_res32": (16, 8, 8, 1280, 1280, 1, "res32"),
    "sq1280_m256": (16, 4, 4, 1280, 1280, 1, "plain"),
    "conv320": (16, 32, 32, 320, 320, 3, "emb32"),
    "conv640": (16, 16, 16, 640, 640, 3, "emb32"),
    "conv1280": (16, 8, 8, 1280, 1280, 3, "emb32"),
    "conv1280_4x4": (16, 4, 4, 1280, 1280, 3, "emb32"),
    "conv2560_4x4": (16, 4, 4, 2560, 1280, 3, "emb32"),
    "conv320_c1": (16, 32, 32, 320, 320, 3, "emb32st"),   # ResBlock conv1 as the network runs it: + GroupNorm statistics
    "conv320_c2": (16, 32, 32, 320, 320, 3, "res32st"),   # ResBlock conv2: fp32 residual, fp32 out, statistics
    "conv960_c1": (16, 32, 32, 960, 320, 3, "emb32st"),
    "up640": (16, 16, 16, 640, 640, 3, "upst"),           # Upsample conv: 16x16 -> 32x32, N = 640
    "ff1_320": (16, 32, 32, 320, 2560, 1, "geglu"),
    "ff2_320": (16, 32, 32, 1280, 320, 1, "res32"),
    "qkv_320": (16, 32, 32, 320, 960, 1, "plain"),
    "big_k": (16, 32, 32, 960, 320, 3, "emb32"),
}
for name, (N, H, W, C, K, R, epi) in SHAPES.items():
    if a.only and a.only not in name:
        continue
    M = N * H * W
    x = torch.randn(M, C, device=DEV).bfloat16()
    w = (torch.randn(K, R, R, C, device=DEV) / math.sqrt(C * R * R)).bfloat16()
    bias = torch.randn(K, device=DEV)
    kw = dict(N=N, H=H, W=W, R=R, S=R, pad=R // 2, bias=bias, workspace=ws)
    Ko = K
    if epi == "plain":
        y = torch.empty(M, K, device=DEV, dtype=torch.bfloat16)
        args = (x, w, y)
    elif epi == "res32":
        r = torch.randn(M, K, device=DEV)
        args = (x, w, None)
        kw.update(residual=r, y32=r)
    elif epi == "emb32":
        e = torch.randn(N, K, device=DEV).bfloat16()
        y32 = torch.empty(M, K, device=DEV)
        args = (x, w, None)
        kw.update(emb=e, y32=y32)
    elif epi in ("emb32st", "res32st", "upst"):
        Mo = M * 4 if epi == "upst" else M
        st = torch.empty((Mo + 127) // 128, K, 2, device=DEV)
        y32 = torch.empty(Mo, K, device=DEV)
        if epi == "emb32st":
            args = (x, w, None)
            kw.update(emb=torch.randn(N, K, device=DEV).bfloat16(), y32=y32, stats=st)
        elif epi == "res32st":
            args = (x, w, None)  # as the network issues it: fp32 residual in, fp32 out, statistics, no bf16 copy
            kw.update(residual=torch.randn(Mo, K, device=DEV), y32=y32, stats=st)
        else:
            args = (x, w, torch.empty(Mo, K, device=DEV, dtype=torch.bfloat16))
            kw.update(upsample=True, y32=y32, stats=st)
            M = Mo
    elif epi == "geglu":
        Ko = K // 2
        y = torch.empty(M, Ko, device=DEV, dtype=torch.bfloat16)
        args = (x, w, y)
        kw.update(act=L.ACT_GEGLU, geglu_block=80)
    d = ops.make_conv_desc(*args, **kw)
    for _ in range(3):
        ops.run_conv_desc(d)
    torch.cuda.synchronize()
    # launches are replayed from a CUDA graph: a Python/ctypes launch costs ~12 us of host time, which would hide every
    # kernel shorter than that; --flush interleaves a 256 MB memset whose own time is measured and subtracted
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        g, gf = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            for _ in range(a.iters):
                if a.flush:
                    flush.zero_()
                ops.run_conv_desc(d)
        with torch.cuda.graph(gf, stream=side):
            for _ in range(a.iters):
                flush.zero_()
    torch.cuda.current_stream().wait_stream(side)

    def run(graph):
        graph.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        graph.replay()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1)
    us = 1e3 * (run(g) - (run(gf) if a.flush else 0.0)) / a.iters
    fl = 2.0 * M * K * C * R * R
    print(f"{name:14s} M={M:6d} N={K:5d} K={C * R * R:6d} {epi:6s}: {us:8.1f} us  {fl / us / 1e6:8.1f} TFLOP/s", flush=True)
