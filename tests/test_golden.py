"""Golden fixtures (tests/golden/ddim_tiny_v1.npz, made by tests/golden/make_golden.py from the oracle):
CPU: the oracle still reproduces them; GPU: the CUDA path matches them through the public sampler API."""
import os

import numpy as np
import pytest
import torch

from oracle import MKDDIMSampler, OracleControlLDM, seeded_state_dict

HERE = os.path.dirname(os.path.abspath(__file__))
G = np.load(os.path.join(HERE, "golden", "ddim_tiny_v1.npz"))
PARAMS = dict(model_channels=64, num_heads=4, context_dim=64)
B, H, S = 2, 16, 6


def inputs(device="cpu"):
    g = torch.Generator().manual_seed(20240601)
    d = {"ctx": torch.randn(B, 77, 64, generator=g), "uc_ctx": torch.randn(B, 77, 64, generator=g),
         "hint": torch.rand(B, 6, 8 * H, 8 * H, generator=g), "x_T": torch.randn(B, 4, H, H, generator=g)}
    return {k: v.to(device) for k, v in d.items()}


def rel(a, b):
    a, b = torch.as_tensor(a).float().cpu(), torch.as_tensor(b).float().cpu()
    return float((a - b).norm() / b.norm())


def test_oracle_reproduces_golden():
    torch.set_num_threads(4)
    m = OracleControlLDM(control_params=PARAMS, unet_params=PARAMS).eval()
    seeded_state_dict(m, 0)
    i = inputs()
    assert np.allclose(i["x_T"].flatten()[:16].numpy(), G["x_T_head"])
    cond = {"c_crossattn": [i["ctx"]], "c_concat": [i["hint"]]}
    s = MKDDIMSampler(m)
    s.make_schedule(S, ddim_eta=0.0, verbose=False)
    assert list(np.flip(s.ddim_timesteps)) == list(G["timesteps"])
    with torch.no_grad():
        x, _ = s.sample(S, B, (4, H, H), cond, eta=0.0, x_T=i["x_T"], verbose=False)
        e0 = m.apply_model(i["x_T"], torch.full((B,), int(G["timesteps"][0])), cond)
    assert rel(e0, G["eps"][0]) < 1e-5
    assert rel(x, G["x"][-1]) < 1e-4


@pytest.mark.gpu
@pytest.mark.parametrize("dtype,tol_eps,tol_cfg", [(torch.float32, 1e-4, 2e-3), (torch.bfloat16, 2e-2, 0.35)])
def test_cuda_path_matches_golden(dtype, tol_eps, tol_cfg):
    from makeupdiffuse_b200 import B200ControlLDM, B200DDIMSampler
    from makeupdiffuse_b200.synth import synthetic_state_dict
    m = B200ControlLDM(PARAMS, PARAMS, dtype=dtype)
    m.load_state_dict(synthetic_state_dict(m, 0))
    i = inputs("cuda")
    cond = {"c_crossattn": [i["ctx"]], "c_concat": [i["hint"]]}
    uc = {"c_crossattn": [i["uc_ctx"]], "c_concat": [i["hint"]]}
    s = B200DDIMSampler(m)
    s.make_schedule(S, ddim_eta=0.0, verbose=False)
    # teacher-forced: feed the golden x_t of every step
    xs = [i["x_T"]] + [torch.as_tensor(G["x"][k]).cuda() for k in range(len(G["timesteps"]) - 1)]
    for k, step in enumerate(G["timesteps"]):
        ts = torch.full((B,), int(step), device="cuda", dtype=torch.long)
        eps = m.apply_model(xs[k], ts, cond)
        assert rel(eps, G["eps"][k]) < tol_eps, (k, rel(eps, G["eps"][k]))
    x0 = s.reconstruct(i["x_T"], cond, t_start=S, unconditional_guidance_scale=9.0, unconditional_conditioning=uc)
    assert rel(x0, G["x0_cfg9"]) < tol_cfg, rel(x0, G["x0_cfg9"])  # free-running, CFG 9 amplifies differences
