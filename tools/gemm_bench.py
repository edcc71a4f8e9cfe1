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
ap.add_argument("--path", default="auto", help="auto | single | pair | both (single then pair, side by side)")
a = ap.parse_args()
PATHS = {"auto": L.PATH_AUTO, "single": L.PATH_TCGEN05_SINGLE, "pair": L.PATH_TCGEN05_PAIR}
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
    # the rest of one evaluation's GEMM mix (profiles/r01_conv_table.json)
    "qkv_640": (16, 16, 16, 640, 1920, 1, "plain"),
    "qkv_1280": (16, 8, 8, 1280, 3840, 1, "plain"),
    "ff1_640": (16, 16, 16, 640, 5120, 1, "geglu"),
    "ff1_1280": (16, 8, 8, 1280, 10240, 1, "geglu"),
    "ff2_640": (16, 16, 16, 2560, 640, 1, "res32"),
    "ff2_1280": (16, 8, 8, 5120, 1280, 1, "res32"),
    "sq640_plain": (16, 16, 16, 640, 640, 1, "plain"),
    "sq1280_plain": (16, 8, 8, 1280, 1280, 1, "plain"),
    "po320_st": (16, 32, 32, 320, 320, 1, "res32yst"),    # SpatialTransformer proj_out: fp32 residual, bf16 + fp32 out, statistics
    "zc320_inj": (16, 32, 32, 320, 320, 1, "inject"),     # zero-conv injection: bf16 slot += alpha * (acc + bias), statistics
    "conv640_c2": (16, 16, 16, 640, 640, 3, "res32st"),
    "conv1920_c1": (16, 16, 16, 1920, 640, 3, "emb32st"),
    "conv1280_c2": (16, 8, 8, 1280, 1280, 3, "res32"),
    "conv2560_8x8": (16, 8, 8, 2560, 1280, 3, "emb32"),
    "skip_1x1": (16, 16, 16, 320, 640, 1, "y32"),
    # the stacked trunk's launches (names g2_*): the same layer of the UNet encoder and the ControlNet trunk on two stacked
    # batches of 16, mkd_conv_desc.wgroups = 2 (twice the rows of the shapes above, two weight sets)
    "g2_conv320_c1": (32, 32, 32, 320, 320, 3, "emb32st"),
    "g2_conv320_c2": (32, 32, 32, 320, 320, 3, "res32st"),
    "g2_sq320_res32": (32, 32, 32, 320, 320, 1, "res32"),
    "g2_conv640": (32, 16, 16, 640, 640, 3, "emb32"),
    "g2_sq640_res32": (32, 16, 16, 640, 640, 1, "res32"),
    "g2_conv1280": (32, 8, 8, 1280, 1280, 3, "emb32"),
    "g2_sq1280_res32": (32, 8, 8, 1280, 1280, 1, "res32"),
    "g2_conv1280_4x4": (32, 4, 4, 1280, 1280, 3, "emb32"),
    "g2_sq1280_m512": (32, 4, 4, 1280, 1280, 1, "res32"),
    "g2_conv2560_4x4": (32, 4, 4, 2560, 1280, 3, "emb32"),
    "m512_conv1280": (8, 8, 8, 1280, 1280, 3, "emb32"),    # 512 rows, one network (8x8 level of configs[3]'s 8 rows per GPU)
    "m512_conv2560": (8, 8, 8, 2560, 1280, 3, "emb32"),
}
for name, (N, H, W, C, K, R, epi) in SHAPES.items():
    if a.only and a.only not in name:
        continue
    M = N * H * W
    wg = 2 if name.startswith("g2_") else 1
    x = torch.randn(M, C, device=DEV).bfloat16()
    w = (torch.randn(wg * K, R, R, C, device=DEV) / math.sqrt(C * R * R)).bfloat16()
    bias = torch.randn(wg * K, device=DEV)
    kw = dict(N=N, H=H, W=W, R=R, S=R, pad=R // 2, bias=bias, workspace=ws, wgroups=wg)
    Ko = K
    if epi == "plain":
        y = torch.empty(M, K, device=DEV, dtype=torch.bfloat16)
        args = (x, w, y)
    elif epi == "res32":
        r = torch.randn(M, K, device=DEV)
        args = (x, w, None)
        kw.update(residual=r, y32=r)
    elif epi == "y32":
        args = (x, w, None)
        kw.update(y32=torch.empty(M, K, device=DEV))
    elif epi == "res32yst":
        st = torch.empty(M // 128, K, 2, device=DEV)
        args = (x, w, torch.empty(M, K, device=DEV, dtype=torch.bfloat16))
        kw.update(residual=torch.randn(M, K, device=DEV), y32=torch.empty(M, K, device=DEV), stats=st)
    elif epi == "inject":
        st = torch.empty(M // 128, K, 2, device=DEV)
        slot = torch.randn(M, K, device=DEV).bfloat16()
        args = (x, w, slot)
        kw.update(residual=slot, alpha=0.7, stats=st)
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
        kw.update(act=L.ACT_GEGLU, geglu_block=80)  # (per path below: 80 for the single-CTA kernel, 128 for the pair kernel)
    fl = 2.0 * M * K * C * R * R
    results = []
    for pname in (("single", "pair") if a.path == "both" else (a.path,)):
        kw2 = dict(kw, path=PATHS[pname])
        if epi == "geglu":
            kw2["geglu_block"] = 80 if pname == "single" else 128
        try:
            d = ops.make_conv_desc(*args, **kw2)
            for _ in range(3):
                ops.run_conv_desc(d)
            torch.cuda.synchronize()
        except RuntimeError as ex:
            results.append(f"{pname}: n/a ({str(ex)[-40:]})")
            continue
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
        results.append(f"{pname}: {us:7.1f} us {fl / us / 1e6:7.1f} TFLOP/s")
    print(f"{name:14s} M={M:6d} N={K:5d} K={C * R * R:6d} {epi:8s}: " + "   ".join(results), flush=True)
