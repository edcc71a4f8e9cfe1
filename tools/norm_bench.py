"""Bandwidth of the norm kernels on the shapes of one UNet+ControlNet step (batch 16): GB/s of algorithmic bytes
(1 read + 1 write per element) against the measured HBM copy peak.  Warm = inputs L2-resident (as inside a step, where
the producer has just written them); --flush = a 256 MB write between launches (HBM-resident inputs)."""
import json
import sys

import torch

sys.path.insert(0, ".")
from makeupdiffuse_b200 import ops  # noqa: E402

DEV = "cuda"
FLUSH = "--flush" in sys.argv
flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"]


def timeit(fn, reps=20):
    """launches are captured into a CUDA graph (a Python/ctypes call costs ~14 us of host time, more than most of these
    kernels), replayed, and timed with two events; --flush interleaves a 256 MB memset (its time is measured alone and
    subtracted)"""
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            fn()
        g, gf = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            for _ in range(reps):
                if FLUSH:
                    flush.zero_()
                fn()
        with torch.cuda.graph(gf, stream=side):
            for _ in range(reps):
                if FLUSH:
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
    return 1e3 * (run(g) - (run(gf) if FLUSH else 0.0)) / reps


print("mode:", "flushed (HBM-resident inputs)" if FLUSH else "warm (L2-resident inputs)", " HBM peak", peak, "GB/s")
N = 16
for name, HW, C in [("32x32 C320", 1024, 320), ("32x32 C640 (cat)", 1024, 640), ("16x16 C640", 256, 640),
                    ("16x16 C1920 (cat)", 256, 1920), ("8x8 C1280", 64, 1280), ("4x4 C2560 (cat)", 16, 2560)]:
    M = N * HW
    x = torch.randn(M, C, device=DEV)
    y = torch.empty(M, C, device=DEV, dtype=torch.bfloat16)
    g, b = torch.ones(C, device=DEV), torch.zeros(C, device=DEV)
    ws = torch.empty(ops.groupnorm_workspace_bytes(N) // 4, device=DEV)
    nbytes = M * C * 6
    us = timeit(lambda: ops.groupnorm(x, y, N, g, b, 1e-5, True, ws))
    line = f"groupnorm(+SiLU) fp32->bf16 {name:18s} {us:7.1f} us {nbytes / us / 1e3:7.0f} GB/s ({nbytes / us / 1e3 / peak:.2f} of peak)"
    if HW % 128 == 0:
        st = torch.randn(M // 128, C, 2, device=DEV).abs() + 1.0
        st[..., 1] += st[..., 0] ** 2 * 128
        us2 = timeit(lambda: ops.groupnorm_apply(x, y, N, g, b, 1e-5, True, st))
        line += f" | apply (fused stats) {us2:7.1f} us {nbytes / us2 / 1e3:7.0f} GB/s ({nbytes / us2 / 1e3 / peak:.2f})"
    print(line, flush=True)
for name, M, C in [("32x32 C320", 16384, 320), ("16x16 C640", 4096, 640), ("8x8 C1280", 1024, 1280)]:
    x = torch.randn(M, C, device=DEV)
    y = torch.empty(M, C, device=DEV, dtype=torch.bfloat16)
    g, b = torch.ones(C, device=DEV), torch.zeros(C, device=DEV)
    us = timeit(lambda: ops.layernorm(x, y, g, b))
    nbytes = M * C * 6
    print(f"layernorm fp32->bf16        {name:18s} {us:7.1f} us {nbytes / us / 1e3:7.0f} GB/s ({nbytes / us / 1e3 / peak:.2f} of peak)", flush=True)
