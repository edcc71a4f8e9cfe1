"""Generates tests/golden/ref_sampler_v1.npz by EXECUTING THE REFERENCE'S OWN ``diffmk/cddim.py`` — run here, in the
build container, from the repo root (the GPU box has no /root/reference; tests only read the committed .npz):

    python tests/golden/make_golden_ref_sampler.py

What this pins.  ``MKDDIMSampler.denoising_step`` (cddim.py:9-79) and ``MKDDIMSampler.reconstruct`` (cddim.py:81-100)
are in-repo reference code; the only thing that keeps the file from importing is its line 2,
``from ldm.models.diffusion.ddim import *`` (third-party lllyasviel/ControlNet, not vendored).  This script installs a
small stand-in for that ONE module (``DDIMSampler.__init__`` / ``make_schedule`` and ``noise_like``, restated from the
published upstream, device-agnostic) and then loads the reference file unmodified from where it lies.  Every number in
the fixture is therefore produced by the reference's own statements for: the CFG batching order [uncond; cond] and the
combine (:18-40), the coefficient gather by ``index`` incl. ``use_original_steps`` (:51-59), pred_x0 (:63), dir_xt
(:74), the noise term and its RNG consumption (:75), x_prev (:78), and the truncated reverse loop (:83-99).
What it does NOT pin: the schedule values themselves (stand-in; covered by KAT K1) and the networks behind
``apply_model`` (a closed-form toy denoiser here).

tests/test_golden.py replays the same toy denoiser through the oracle sampler (CPU) and through B200DDIMSampler
(GPU, fused mkd_ddim_update kernel) and compares with this fixture.
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from toy_denoiser import ToyDenoiser, toy_inputs  # noqa: E402  (shared with the tests)

REF = "/root/reference/diffmk/cddim.py"


def _install_upstream_stand_in():
    """ldm.models.diffusion.ddim: just what cddim.py needs from its star import (torch, np, DDIMSampler, noise_like)."""

    class DDIMSampler(object):
        def __init__(self, model, schedule="linear", **kwargs):
            self.model = model
            self.ddpm_num_timesteps = model.num_timesteps
            self.schedule = schedule

        def register_buffer(self, name, attr):
            setattr(self, name, attr)

        def make_schedule(self, ddim_num_steps, ddim_discretize="uniform", ddim_eta=0., verbose=True):
            c = self.ddpm_num_timesteps // ddim_num_steps
            self.ddim_timesteps = np.asarray(list(range(0, self.ddpm_num_timesteps, c))) + 1
            alphas_cumprod = self.model.alphas_cumprod
            to_torch = lambda x: x.clone().detach().to(torch.float32)  # noqa: E731
            self.register_buffer('betas', to_torch(self.model.betas))
            self.register_buffer('alphas_cumprod', to_torch(alphas_cumprod))
            self.register_buffer('alphas_cumprod_prev', to_torch(self.model.alphas_cumprod_prev))
            self.register_buffer('sqrt_one_minus_alphas_cumprod', to_torch(np.sqrt(1. - alphas_cumprod.cpu())))
            ac = alphas_cumprod.cpu().numpy()
            alphas = ac[self.ddim_timesteps]
            alphas_prev = np.asarray([ac[0]] + ac[self.ddim_timesteps[:-1]].tolist())
            sigmas = ddim_eta * np.sqrt((1 - alphas_prev) / (1 - alphas) * (1 - alphas / alphas_prev))
            self.register_buffer('ddim_sigmas', torch.as_tensor(sigmas))
            self.register_buffer('ddim_alphas', torch.as_tensor(alphas))
            self.register_buffer('ddim_alphas_prev', alphas_prev)
            self.register_buffer('ddim_sqrt_one_minus_alphas', torch.as_tensor(np.sqrt(1. - alphas)))
            self.register_buffer('ddim_sigmas_for_original_num_steps', ddim_eta * torch.sqrt(
                (1 - self.alphas_cumprod_prev) / (1 - self.alphas_cumprod) * (1 - self.alphas_cumprod / self.alphas_cumprod_prev)))

    def noise_like(shape, device, repeat=False):
        if repeat:
            return torch.randn((1, *shape[1:]), device=device).repeat(shape[0], *((1,) * (len(shape) - 1)))
        return torch.randn(shape, device=device)

    names = ["ldm", "ldm.models", "ldm.models.diffusion", "ldm.models.diffusion.ddim"]
    mods = {n: types.ModuleType(n) for n in names}
    for n in names[:-1]:
        mods[n].__path__ = []
    leaf = mods[names[-1]]
    leaf.torch, leaf.np, leaf.DDIMSampler, leaf.noise_like = torch, np, DDIMSampler, noise_like
    leaf.__all__ = ["torch", "np", "DDIMSampler", "noise_like"]
    sys.modules.update(mods)


def load_reference_sampler():
    _install_upstream_stand_in()
    spec = importlib.util.spec_from_file_location("ref_cddim", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)  # the reference file, unmodified, read from /root/reference
    return mod.MKDDIMSampler


def main():
    torch.set_num_threads(2)
    Sampler = load_reference_sampler()
    m = ToyDenoiser()
    i = toy_inputs()
    c_dict = {"c_crossattn": [i["ctx"]], "c_concat": [i["hint"]]}
    u_dict = {"c_crossattn": [i["uc_ctx"]], "c_concat": [i["hint"]]}
    out = {}
    with torch.no_grad():
        # case A: S = 50, eta = 0, no guidance: three single steps + the full reverse loop from x_T
        s = Sampler(m)
        s.make_schedule(50, ddim_eta=0.0, verbose=False)
        B = i["x_T"].shape[0]
        for index in (49, 25, 0):
            ts = torch.full((B,), int(s.ddim_timesteps[index]), dtype=torch.long)
            xp, p0 = s.denoising_step(i["x_T"], c_dict, ts, index=index)
            out[f"A_step{index}_x_prev"], out[f"A_step{index}_pred_x0"] = xp.numpy(), p0.numpy()
        out["A_reconstruct50"] = s.reconstruct(i["x_T"], c_dict, t_start=50).numpy()
        # case B: S = 20, guidance 9 (diffusion_makeup.py:308), truncated loop t_start = 12; dict conditioning
        s = Sampler(m)
        s.make_schedule(20, ddim_eta=0.0, verbose=False)
        seen = []
        out["B_reconstruct12_cfg9"] = s.reconstruct(i["x_T"], c_dict, t_start=12, unconditional_guidance_scale=9.0,
                                                    unconditional_conditioning=u_dict, callback=seen.append).numpy()
        out["B_callback_args"] = np.asarray(seen)
        ts = torch.full((B,), int(s.ddim_timesteps[7]), dtype=torch.long)
        # list and bare-tensor conditioning take the other two branches of the CFG concatenation (:32-38)
        xp, p0 = s.denoising_step(i["x_T"], [i["ctx"], i["hint"]], ts, index=7, unconditional_guidance_scale=3.5,
                                  unconditional_conditioning=[i["uc_ctx"], i["hint"]])
        out["B_list_x_prev"], out["B_list_pred_x0"] = xp.numpy(), p0.numpy()
        xp, p0 = s.denoising_step(i["x_T"], i["ctx"], ts, index=7, unconditional_guidance_scale=3.5,
                                  unconditional_conditioning=i["uc_ctx"])
        out["B_tensor_x_prev"], out["B_tensor_pred_x0"] = xp.numpy(), p0.numpy()
        # guidance scale exactly 1 takes the single-call branch even with an unconditional conditioning (:15-16)
        xp, _ = s.denoising_step(i["x_T"], c_dict, ts, index=7, unconditional_guidance_scale=1.0, unconditional_conditioning=u_dict)
        out["B_scale1_x_prev"] = xp.numpy()
        out["B_apply_model_calls_scale1"] = np.asarray(m.calls[-1])
        # case C: eta = 0.5 (sigma > 0): the noise term, temperature, and the RNG stream (one randn per step)
        s = Sampler(m)
        s.make_schedule(20, ddim_eta=0.5, verbose=False)
        torch.manual_seed(1234)
        xp, p0 = s.denoising_step(i["x_T"], c_dict, ts, index=7, temperature=0.8)
        out["C_eta05_x_prev"], out["C_eta05_pred_x0"] = xp.numpy(), p0.numpy()
        out["C_rng_after"] = torch.randn(4).numpy()  # next draws of the global generator after ONE step
        torch.manual_seed(1234)
        out["C_reconstruct5_eta05"] = s.reconstruct(i["x_T"], c_dict, t_start=5).numpy()
        torch.manual_seed(1234)
        xp, _ = s.denoising_step(i["x_T"], c_dict, ts, index=7, repeat_noise=True)
        out["C_repeat_noise_x_prev"] = xp.numpy()
        # case D: use_original_steps (the 1000-step tables of the model, :51-54 and :83).  cddim.py:54 reads
        # ddim_sigmas_for_original_num_steps off the MODEL: without it the reference raises AttributeError (recorded),
        # with it the loop walks timesteps 0..t_start-1 of the 1000-step tables
        s = Sampler(m)
        s.make_schedule(50, ddim_eta=0.0, verbose=False)
        try:
            s.reconstruct(i["x_T"], c_dict, t_start=4, use_original_steps=True)
            out["D_missing_attr_raises"] = np.asarray(0)
        except AttributeError:
            out["D_missing_attr_raises"] = np.asarray(1)
        m.ddim_sigmas_for_original_num_steps = torch.zeros(1000)
        out["D_original_steps_reconstruct4"] = s.reconstruct(i["x_T"], c_dict, t_start=4, use_original_steps=True).numpy()
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_sampler_v1.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", len(out), "arrays")


if __name__ == "__main__":
    main()
