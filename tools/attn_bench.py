"""Times mkd_attention on the SpatialTransformer shapes of one UNet+ControlNet step (batch 16, 256^2 and 512^2 levels).
MKD_ATTN=mma selects the older mma.sync kernel for an A/B run."""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, ".")
from makeupdiffuse_b200 import ops  # noqa: E402

DEV = "cuda"
SHAPES = [("self 32x32 d40", 16, 8, 1024, 1024, 40), ("self 16x16 d80", 16, 8, 256, 256, 80),
          ("self 8x8 d160", 16, 8, 64, 64, 160), ("self 4x4 d160", 16, 8, 16, 16, 160),
          ("cross 32x32 d40", 16, 8, 1024, 77, 40), ("cross 16x16 d80", 16, 8, 256, 77, 80),
          ("cross 8x8 d160", 16, 8, 64, 77, 160), ("self 64x64 d40 (512^2)", 8, 8, 4096, 4096, 40),
          ("self 32x32 d80 (512^2)", 8, 8, 1024, 1024, 80)]
ONLY = sys.argv[1] if len(sys.argv) > 1 else ""
print("kernel:", os.environ.get("MKD_ATTN", "tcgen05"))
for name, B, heads, Nq, Nkv, d in SHAPES:
    if ONLY and ONLY not in name:
        continue
    C = heads * d
    g = torch.Generator(device=DEV).manual_seed(0)
    if Nq == Nkv:
        qkv = torch.randn(B * Nq, 3 * C, device=DEV, generator=g).bfloat16()
        q, k, v = qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:]
    else:
        q = torch.randn(B * Nq, C, device=DEV, generator=g).bfloat16()
        kv = torch.randn(B * Nkv, 2 * C, device=DEV, generator=g).bfloat16()
        k, v = kv[:, :C], kv[:, C:]
    o = torch.empty(B * Nq, C, device=DEV, dtype=torch.bfloat16)
    run = lambda: ops.attention(q, k, v, o, B=B, heads=heads, Nq=Nq, Nkv=Nkv, d=d, scale=d ** -0.5)  # noqa: E731
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20
    e0.record()
    for _ in range(reps):
        run()
    e1.record()
    torch.cuda.synchronize()
    us = 1e3 * e0.elapsed_time(e1) / reps
    sp = lambda t, n: t.float().reshape(B, n, heads, d).permute(0, 2, 1, 3)  # noqa: E731
    ref = F.scaled_dot_product_attention(sp(q, Nq), sp(k, Nkv), sp(v, Nkv)).permute(0, 2, 1, 3).reshape(B * Nq, C)
    err = float((o.float() - ref).norm() / ref.norm())
    fl = 4.0 * B * heads * Nq * Nkv * d
    print(f"{name:26s} {us:8.1f} us  {fl / us / 1e6:7.1f} TFLOP/s  rel-L2 {err:.2e}", flush=True)
