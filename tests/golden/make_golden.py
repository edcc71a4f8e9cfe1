"""Generates tests/golden/ddim_tiny_v1.npz from the ORACLE (CPU, fp32) — run from the repo root:

    python tests/golden/make_golden.py

The reference ships no golden vectors for this path and its ldm/cldm dependency is not importable,
so these fixtures pin the oracle against ITSELF over time (any later edit of oracle/ that changes results trips
tests/test_golden.py) and give the GPU tests a device-independent target: reduced-width networks with the yaml's
topology, seeded hash weights, a 6-step DDIM run with and without classifier-free guidance.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import MKDDIMSampler, OracleControlLDM, seeded_state_dict  # noqa: E402

PARAMS = dict(model_channels=64, num_heads=4, context_dim=64)
B, H, S = 2, 16, 6


def inputs():
    g = torch.Generator().manual_seed(20240601)
    return {"ctx": torch.randn(B, 77, 64, generator=g), "uc_ctx": torch.randn(B, 77, 64, generator=g),
            "hint": torch.rand(B, 6, 8 * H, 8 * H, generator=g), "x_T": torch.randn(B, 4, H, H, generator=g)}


def main():
    torch.set_num_threads(4)
    m = OracleControlLDM(control_params=PARAMS, unet_params=PARAMS).eval()
    seeded_state_dict(m, 0)
    i = inputs()
    cond = {"c_crossattn": [i["ctx"]], "c_concat": [i["hint"]]}
    uc = {"c_crossattn": [i["uc_ctx"]], "c_concat": [i["hint"]]}
    s = MKDDIMSampler(m)
    s.make_schedule(S, ddim_eta=0.0, verbose=False)
    steps = np.flip(s.ddim_timesteps)  # NB: 1000 // 6 = 166 -> 7 timesteps (upstream quirk when T % S != 0); kept
    total = len(steps)
    out = {"timesteps": np.asarray(steps), "x_T_head": i["x_T"].flatten()[:16].numpy()}
    with torch.no_grad():
        x = i["x_T"]
        eps_all, x_all = [], []
        for k, step in enumerate(steps):
            ts = torch.full((B,), int(step), dtype=torch.long)
            eps_all.append(m.apply_model(x, ts, cond).numpy())
            x, _ = s.denoising_step(x, cond, ts, index=total - k - 1)
            x_all.append(x.numpy())
        out["eps"] = np.stack(eps_all)          # [S, B, 4, H, H] teacher-forced targets
        out["x"] = np.stack(x_all)              # x_t after each step
        xc = s.reconstruct(i["x_T"], cond, t_start=S, unconditional_guidance_scale=9.0, unconditional_conditioning=uc)
        out["x0_cfg9"] = xc.numpy()
        ctrl = m.control_model(x=i["x_T"], hint=i["hint"], timesteps=torch.full((B,), int(steps[0])), context=i["ctx"])
        out["control_norms"] = np.asarray([float(c.norm()) for c in ctrl])
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ddim_tiny_v1.npz")
    np.savez_compressed(path, **{k: np.asarray(v, dtype=np.float32 if np.asarray(v).dtype.kind == "f" else None) for k, v in out.items()})
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
