"""Fixture produced by EXECUTING the reference's own diffmk/cddim.py (tests/golden/make_golden_ref_sampler.py, run in
the build container against /root/reference; committed as tests/golden/ref_sampler_v1.npz).  It pins the sampler rows
A4 / A5 of SURVEY.md §8(a) — CFG batching order and combine, coefficient gather, pred_x0 / dir_xt / x_prev, the noise
term and its RNG consumption, the truncated reverse loop, the use_original_steps tables — against reference code, not
against a restatement.

CPU: the oracle sampler and the product's host logic (kernel contract restated by tests/fake_ops.py) reproduce it.
GPU: B200DDIMSampler (fused mkd_ddim_update kernel, incl. the CFG combine) reproduces it.
"""
import os

import numpy as np
import pytest
import torch

import fake_ops
from makeupdiffuse_b200 import B200DDIMSampler, ops
from oracle import MKDDIMSampler
from toy_denoiser import ToyDenoiser, toy_inputs

HERE = os.path.dirname(os.path.abspath(__file__))
G = np.load(os.path.join(HERE, "golden", "ref_sampler_v1.npz"))
# fp32 elementwise chains evaluated in a different association (fused kernel / host-side coefficient folding):
# a few ulp per step, compounding over <= 50 steps of a contractive map
TOL_STEP, TOL_LOOP = 2e-6, 2e-5


def err(a, ref):
    a, ref = torch.as_tensor(a).float().cpu(), torch.as_tensor(ref).float()
    return float((a - ref).abs().max() / ref.abs().max())


def run_cases(Sampler, device, with_rng_cases):
    m = ToyDenoiser(device)
    i = toy_inputs(device)
    c = {"c_crossattn": [i["ctx"]], "c_concat": [i["hint"]]}
    u = {"c_crossattn": [i["uc_ctx"]], "c_concat": [i["hint"]]}
    B = i["x_T"].shape[0]
    worst = {}
    with torch.no_grad():
        s = Sampler(m)
        s.make_schedule(50, ddim_eta=0.0, verbose=False)
        for index in (49, 25, 0):
            ts = torch.full((B,), int(s.ddim_timesteps[index]), dtype=torch.long, device=device)
            xp, p0 = s.denoising_step(i["x_T"], c, ts, index=index)
            worst[f"A_step{index}"] = max(err(xp, G[f"A_step{index}_x_prev"]), err(p0, G[f"A_step{index}_pred_x0"]))
            assert worst[f"A_step{index}"] < TOL_STEP
        worst["A_loop"] = err(s.reconstruct(i["x_T"], c, t_start=50), G["A_reconstruct50"])
        assert worst["A_loop"] < TOL_LOOP

        s = Sampler(m)
        s.make_schedule(20, ddim_eta=0.0, verbose=False)
        seen = []
        x = s.reconstruct(i["x_T"], c, t_start=12, unconditional_guidance_scale=9.0, unconditional_conditioning=u, callback=seen.append)
        worst["B_cfg9_loop"] = err(x, G["B_reconstruct12_cfg9"])
        assert worst["B_cfg9_loop"] < TOL_LOOP
        assert seen == list(G["B_callback_args"])
        ts = torch.full((B,), int(s.ddim_timesteps[7]), dtype=torch.long, device=device)
        xp, p0 = s.denoising_step(i["x_T"], [i["ctx"], i["hint"]], ts, index=7, unconditional_guidance_scale=3.5,
                                  unconditional_conditioning=[i["uc_ctx"], i["hint"]])
        worst["B_list"] = max(err(xp, G["B_list_x_prev"]), err(p0, G["B_list_pred_x0"]))
        xp, p0 = s.denoising_step(i["x_T"], i["ctx"], ts, index=7, unconditional_guidance_scale=3.5, unconditional_conditioning=i["uc_ctx"])
        worst["B_tensor"] = max(err(xp, G["B_tensor_x_prev"]), err(p0, G["B_tensor_pred_x0"]))
        assert worst["B_list"] < TOL_STEP and worst["B_tensor"] < TOL_STEP
        m.calls.clear()
        xp, _ = s.denoising_step(i["x_T"], c, ts, index=7, unconditional_guidance_scale=1.0, unconditional_conditioning=u)
        assert err(xp, G["B_scale1_x_prev"]) < TOL_STEP
        assert m.calls == [int(G["B_apply_model_calls_scale1"])] == [B]  # one un-doubled call (cddim.py:15-16)

        if with_rng_cases:  # torch's CPU generator: only comparable when the sampler draws on the CPU
            s = Sampler(m)
            s.make_schedule(20, ddim_eta=0.5, verbose=False)
            torch.manual_seed(1234)
            xp, p0 = s.denoising_step(i["x_T"], c, ts, index=7, temperature=0.8)
            worst["C_eta05"] = max(err(xp, G["C_eta05_x_prev"]), err(p0, G["C_eta05_pred_x0"]))
            assert worst["C_eta05"] < TOL_STEP
            assert np.array_equal(torch.randn(4).numpy(), G["C_rng_after"])  # exactly one randn(x.shape) consumed
            torch.manual_seed(1234)
            worst["C_loop"] = err(s.reconstruct(i["x_T"], c, t_start=5), G["C_reconstruct5_eta05"])
            assert worst["C_loop"] < TOL_LOOP
            torch.manual_seed(1234)
            xp, _ = s.denoising_step(i["x_T"], c, ts, index=7, repeat_noise=True)
            worst["C_repeat"] = err(xp, G["C_repeat_noise_x_prev"])
            assert worst["C_repeat"] < TOL_STEP

        s = Sampler(m)
        s.make_schedule(50, ddim_eta=0.0, verbose=False)
        assert int(G["D_missing_attr_raises"]) == 1
        with pytest.raises(AttributeError):  # cddim.py:54 reads the table off the model
            s.reconstruct(i["x_T"], c, t_start=4, use_original_steps=True)
        m.ddim_sigmas_for_original_num_steps = torch.zeros(1000, device=device)
        worst["D_original"] = err(s.reconstruct(i["x_T"], c, t_start=4, use_original_steps=True), G["D_original_steps_reconstruct4"])
        assert worst["D_original"] < TOL_LOOP
    return worst


def test_oracle_sampler_matches_reference_cddim():
    print(run_cases(MKDDIMSampler, "cpu", True))


def test_product_host_logic_matches_reference_cddim(monkeypatch):
    for name in fake_ops.ALL:
        monkeypatch.setattr(ops, name, getattr(fake_ops, name))
    print(run_cases(B200DDIMSampler, "cpu", True))


@pytest.mark.gpu
def test_b200_sampler_matches_reference_cddim():
    print(run_cases(B200DDIMSampler, "cuda", False))
