"""Back-to-back launches of one pair-kernel shape (no host synchronisation in between): protocol races show up here."""
import math
import sys

import torch

sys.path.insert(0, ".")
from makeupdiffuse_b200 import _lib as L, ops  # noqa: E402

M, C, K = 16384, 1280, 320
x = torch.randn(M, C, device="cuda").bfloat16()
w = (torch.randn(K, 1, 1, C, device="cuda") / math.sqrt(C)).bfloat16()
b = torch.randn(K, device="cuda")
r = torch.randn(M, K, device="cuda")
y = torch.empty(M, K, device="cuda", dtype=torch.bfloat16)
ref = None
for it in range(int(sys.argv[1]) if len(sys.argv) > 1 else 300):
    ops.conv2d(x, w, y, N=1, H=1, W=M, bias=b, residual=r, path=L.PATH_TCGEN05_PAIR)
    if it % 50 == 0:
        torch.cuda.synchronize()
        if ref is None:
            ref = (x.float() @ w.reshape(K, C).float().t() + b + r)
            print("rel err", float((y.float() - ref).norm() / ref.norm()))
        print("iter", it, "ok", flush=True)
torch.cuda.synchronize()
print("done")
