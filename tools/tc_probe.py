"""First-light probe for the tcgen05 GEMM kernel: each stage runs in its own process (a faulting kernel kills the
context), prints a decoded error picture, and never hangs the box (run under `timeout`)."""
import math
import sys

import torch

sys.path.insert(0, ".")
from makeupdiffuse_b200 import _lib as L  # noqa: E402
from makeupdiffuse_b200 import ops  # noqa: E402

torch.manual_seed(0)
DEV = "cuda"


def run(M, K, N, mode="rand", conv=None, **kw):
    if conv:
        Nn, H, W = conv
        M = Nn * H * W
        R = 3
    else:
        Nn, H, W, R = 1, 1, M, 1
    if mode == "eye":
        x = torch.zeros(M, K, device=DEV)
        x[torch.arange(M), torch.arange(M) % K] = 1.0
    else:
        x = torch.randn(M, K, device=DEV)
    x = x.bfloat16()
    w = (torch.randn(N, R, R, K, device=DEV) / math.sqrt(K * R * R)).bfloat16()
    y = torch.full((M, N), 7.0, device=DEV, dtype=torch.bfloat16)
    ops.conv2d(x, w, y, N=Nn, H=H, W=W, R=R, S=R, pad=R // 2, path=L.PATH_TCGEN05, **kw)
    torch.cuda.synchronize()
    if conv:
        xr = x.float().reshape(Nn, H, W, K).permute(0, 3, 1, 2)
        ref = torch.nn.functional.conv2d(xr, w.float().permute(0, 3, 1, 2), padding=1).permute(0, 2, 3, 1).reshape(M, N)
    else:
        ref = x.float() @ w.float().reshape(N, K).t()
    err = (y.float() - ref).abs()
    relerr = float((y.float() - ref).norm() / ref.norm())
    print(f"M={M} K={K} N={N} mode={mode} conv={conv}: rel={relerr:.3e} max={float(err.max()):.3e} "
          f"untouched={int((y == 7.0).sum())}", flush=True)
    if relerr > 1e-2:
        bad = (err > 0.05).nonzero()
        print("  first bad (row, col):", bad[:8].tolist(), flush=True)
        rows_bad = (err > 0.05).any(1).nonzero().flatten()
        cols_bad = (err > 0.05).any(0).nonzero().flatten()
        print("  bad rows:", rows_bad[:32].tolist(), "... count", len(rows_bad), flush=True)
        print("  bad cols:", cols_bad[:32].tolist(), "... count", len(cols_bad), flush=True)
        print("  y[0,:8]  ", y[0, :8].float().tolist(), flush=True)
        print("  ref[0,:8]", ref[0, :8].tolist(), flush=True)
    return relerr


STAGES = {
    "a": lambda: run(128, 64, 32, "eye"),
    "b": lambda: run(128, 64, 32),
    "c": lambda: run(128, 256, 32),
    "d": lambda: run(128, 128, 160),
    "e": lambda: run(1000, 320, 960),
    "f": lambda: run(0, 320, 320, conv=(2, 32, 32)),
    "g": lambda: run(0, 1280, 1280, conv=(5, 4, 4)),
    "h": lambda: run(0, 64, 64, conv=(1, 8, 8)),
}

if __name__ == "__main__":
    ops.device_ok(0)
    r = STAGES[sys.argv[1]]()
    sys.exit(0 if r < 1e-2 else 1)
