"""Times B200FirstStageDecoder.decode on a batch of 256^2 images (yaml-sized decoder, synthetic seeded weights)."""
import sys

import torch

sys.path.insert(0, ".")
from makeupdiffuse_b200 import B200FirstStageDecoder  # noqa: E402
from makeupdiffuse_b200.synth import synthetic_first_stage_state_dict  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
dev = torch.device("cuda", 0)
m = B200FirstStageDecoder(dtype=torch.bfloat16)
sd = synthetic_first_stage_state_dict(m, 0, dev)
m.load_state_dict(sd)
z = torch.randn(B, 4, 32, 32, device=dev)
for _ in range(2):
    out = m.decode(z)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    out = m.decode(z)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
fl = 0.6222e12 * B  # 0.622 TFLOP per 256^2 image (convs + mid attention, enumerated from the module tree)
print(f"decode batch {B} x 256^2: {ms:.2f} ms  ({ms / B:.2f} ms / image, {fl / ms / 1e9:.0f} TFLOP/s over the whole decode), "
      f"out {tuple(out.shape)} finite={bool(torch.isfinite(out).all())}")
