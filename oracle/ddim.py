"""Oracle DDIM sampler: the loop the reference runs 50x per image.

``MKDDIMSampler.denoising_step`` / ``reconstruct`` follow ``diffmk/cddim.py:9-79`` / ``:81-100`` line by line
in meaning (not in text): CFG batching order [uncond; cond] (:18-39), combine e_u + s*(e_c - e_u) (:40),
coefficient gather by ``index`` (:51-59), pred_x0 (:63), dir_xt (:74), noise drawn every step even when
sigma == 0 (:75), x_prev (:78).  ``make_schedule`` / ``sample`` / ``ddim_sampling`` / ``decode`` /
``stochastic_encode`` restate the inherited upstream ``ldm.models.diffusion.ddim.DDIMSampler`` (not vendored;
SURVEY.md §8(a) rows A2/A3); pinned by KATs K1, K2, K5, K6.

Unlike upstream nothing is force-moved to "cuda": buffers live on ``model.device`` so the oracle runs on CPU.

Test infrastructure (see oracle/__init__.py).
"""
from __future__ import annotations

import numpy as np
import torch


def ddim_timesteps_uniform(num_ddim: int, num_ddpm: int) -> np.ndarray:
    """c = T // S ; arange(0, T, c) + 1   (S=50 -> 1, 21, ..., 981)."""
    c = num_ddpm // num_ddim
    return np.asarray(list(range(0, num_ddpm, c))) + 1


def ddim_parameters(alphacums: np.ndarray, steps: np.ndarray, eta: float):
    a = alphacums[steps]
    a_prev = np.asarray([alphacums[0]] + alphacums[steps[:-1]].tolist())
    sig = eta * np.sqrt((1 - a_prev) / (1 - a) * (1 - a / a_prev))
    return sig, a, a_prev


def _cat_uncond_first(uc, c):
    """cddim.py:20-38: every tensor leaf becomes cat([uncond, cond]) along batch."""
    if isinstance(c, dict):
        assert isinstance(uc, dict)
        return {k: ([torch.cat([uc[k][i], c[k][i]]) for i in range(len(c[k]))] if isinstance(c[k], list)
                    else torch.cat([uc[k], c[k]])) for k in c}
    if isinstance(c, list):
        assert isinstance(uc, list)
        return [torch.cat([uc[i], c[i]]) for i in range(len(c))]
    return torch.cat([uc, c])


class DDIMSampler:
    def __init__(self, model, schedule="linear", **kwargs):
        self.model = model
        self.ddpm_num_timesteps = model.num_timesteps
        self.schedule = schedule

    def _f32(self, a):
        return torch.as_tensor(np.asarray(a), dtype=torch.float32).to(self.model.device)

    def make_schedule(self, ddim_num_steps, ddim_discretize="uniform", ddim_eta=0.0, verbose=True):
        assert ddim_discretize == "uniform"
        m = self.model
        self.ddim_timesteps = ddim_timesteps_uniform(ddim_num_steps, self.ddpm_num_timesteps)
        ac = m.alphas_cumprod.detach().float()
        assert ac.shape[0] == self.ddpm_num_timesteps, "alphas have to be defined for each timestep"
        self.betas = m.betas.float()
        self.alphas_cumprod = ac
        self.alphas_cumprod_prev = m.alphas_cumprod_prev.float()
        self.sqrt_alphas_cumprod = ac.sqrt()
        self.sqrt_one_minus_alphas_cumprod = (1.0 - ac).sqrt()
        sig, a, a_prev = ddim_parameters(ac.cpu().numpy(), self.ddim_timesteps, ddim_eta)
        # upstream keeps a / sqrt(1-a) as numpy (float32) and sigmas / a_prev as well; torch.full() in the
        # step casts whatever it gets to fp32 (SURVEY note N1)
        self.ddim_sigmas = self._f32(sig)
        self.ddim_alphas = self._f32(a)
        self.ddim_alphas_prev = np.asarray(a_prev)
        self.ddim_sqrt_one_minus_alphas = self._f32(np.sqrt(1.0 - a))
        acp = self.alphas_cumprod_prev
        self.ddim_sigmas_for_original_num_steps = ddim_eta * torch.sqrt(
            (1 - acp) / (1 - ac) * (1 - ac / acp))

    # -- sampling from noise (upstream sample/ddim_sampling) ----------------------------------------------
    def sample(self, S, batch_size, shape, conditioning=None, callback=None, normals_sequence=None,
               img_callback=None, quantize_x0=False, eta=0.0, mask=None, x0=None, temperature=1.0,
               noise_dropout=0.0, score_corrector=None, corrector_kwargs=None, verbose=True, x_T=None,
               log_every_t=100, unconditional_guidance_scale=1.0, unconditional_conditioning=None,
               dynamic_threshold=None, ucg_schedule=None, **kwargs):
        self.make_schedule(ddim_num_steps=S, ddim_eta=eta, verbose=verbose)
        C, H, W = shape
        return self.ddim_sampling(conditioning, (batch_size, C, H, W), callback=callback,
                                  img_callback=img_callback, quantize_denoised=quantize_x0, mask=mask, x0=x0,
                                  noise_dropout=noise_dropout, temperature=temperature,
                                  score_corrector=score_corrector, corrector_kwargs=corrector_kwargs, x_T=x_T,
                                  log_every_t=log_every_t,
                                  unconditional_guidance_scale=unconditional_guidance_scale,
                                  unconditional_conditioning=unconditional_conditioning,
                                  dynamic_threshold=dynamic_threshold)

    def ddim_sampling(self, cond, shape, x_T=None, callback=None, img_callback=None, quantize_denoised=False,
                      mask=None, x0=None, log_every_t=100, temperature=1.0, noise_dropout=0.0,
                      score_corrector=None, corrector_kwargs=None, unconditional_guidance_scale=1.0,
                      unconditional_conditioning=None, dynamic_threshold=None, **kwargs):
        dev = self.model.device
        b = shape[0]
        img = torch.randn(shape, device=dev) if x_T is None else x_T
        inter = {"x_inter": [img], "pred_x0": [img]}
        steps = np.flip(self.ddim_timesteps)
        total = steps.shape[0]
        for i, step in enumerate(steps):
            index = total - i - 1
            ts = torch.full((b,), int(step), device=dev, dtype=torch.long)
            if mask is not None:
                img = self.model.q_sample(x0, ts) * mask + (1.0 - mask) * img
            img, pred_x0 = self.p_sample_ddim(
                img, cond, ts, index=index, quantize_denoised=quantize_denoised, temperature=temperature,
                noise_dropout=noise_dropout, score_corrector=score_corrector,
                corrector_kwargs=corrector_kwargs, unconditional_guidance_scale=unconditional_guidance_scale,
                unconditional_conditioning=unconditional_conditioning, dynamic_threshold=dynamic_threshold)
            if callback:
                callback(i)
            if img_callback:
                img_callback(pred_x0, i)
            if index % log_every_t == 0 or index == total - 1:
                inter["x_inter"].append(img)
                inter["pred_x0"].append(pred_x0)
        return img, inter

    def p_sample_ddim(self, *a, **k):
        with torch.no_grad():
            return self._step(*a, **k)

    def stochastic_encode(self, x0, t, use_original_steps=False, noise=None):
        sa = self.sqrt_alphas_cumprod if use_original_steps else self.ddim_alphas.sqrt()
        s1 = self.sqrt_one_minus_alphas_cumprod if use_original_steps else self.ddim_sqrt_one_minus_alphas
        noise = torch.randn_like(x0) if noise is None else noise
        ex = lambda v: v.gather(-1, t).reshape(-1, 1, 1, 1)  # noqa: E731
        return ex(sa) * x0 + ex(s1) * noise

    def decode(self, x_latent, cond, t_start, **kw):
        with torch.no_grad():
            return self._reverse(x_latent, cond, t_start, **kw)

    # -- one x_t -> x_{t-1} update -------------------------------------------------------------------------
    def _step(self, x, c, t, index, repeat_noise=False, use_original_steps=False, quantize_denoised=False,
              temperature=1.0, noise_dropout=0.0, score_corrector=None, corrector_kwargs=None,
              unconditional_guidance_scale=1.0, unconditional_conditioning=None, dynamic_threshold=None):
        b, dev = x.shape[0], x.device
        m = self.model
        if unconditional_conditioning is None or unconditional_guidance_scale == 1.0:
            out = m.apply_model(x, t, c)
        else:
            both = m.apply_model(torch.cat([x] * 2), torch.cat([t] * 2),
                                 _cat_uncond_first(unconditional_conditioning, c))
            e_u, e_c = both.chunk(2)
            out = e_u + unconditional_guidance_scale * (e_c - e_u)
        e_t = m.predict_eps_from_z_and_v(x, t, out) if m.parameterization == "v" else out
        if score_corrector is not None:
            assert m.parameterization == "eps", "not implemented"
            e_t = score_corrector.modify_score(m, e_t, x, t, c, **corrector_kwargs)

        if use_original_steps:
            # cddim.py:54 reads this table off the MODEL (upstream registers it on the sampler): a model without the
            # attribute raises AttributeError in the reference, and so does this restatement
            A, AP, S1, SG = (m.alphas_cumprod, m.alphas_cumprod_prev, m.sqrt_one_minus_alphas_cumprod,
                             m.ddim_sigmas_for_original_num_steps)
        else:
            A, AP, S1, SG = (self.ddim_alphas, self.ddim_alphas_prev, self.ddim_sqrt_one_minus_alphas,
                             self.ddim_sigmas)
        full = lambda v: torch.full((b, 1, 1, 1), float(v[index]), device=dev)  # noqa: E731
        a_t, a_prev, sigma_t, s1m = full(A), full(AP), full(SG), full(S1)

        if m.parameterization != "v":
            pred_x0 = (x - s1m * e_t) / a_t.sqrt()
        else:
            pred_x0 = m.predict_start_from_z_and_v(x, t, out)
        if quantize_denoised:
            pred_x0, _, *_ = m.first_stage_model.quantize(pred_x0)
        if dynamic_threshold is not None:
            raise NotImplementedError()
        dir_xt = (1.0 - a_prev - sigma_t ** 2).sqrt() * e_t
        shape = (1, *x.shape[1:]) if repeat_noise else x.shape
        draw = torch.randn(shape, device=dev)
        if repeat_noise:
            draw = draw.repeat(b, *((1,) * (len(x.shape) - 1)))
        noise = sigma_t * draw * temperature
        if noise_dropout > 0.0:
            noise = torch.nn.functional.dropout(noise, p=noise_dropout)
        return a_prev.sqrt() * pred_x0 + dir_xt + noise, pred_x0

    def _reverse(self, x_latent, cond, t_start, unconditional_guidance_scale=1.0,
                 unconditional_conditioning=None, use_original_steps=False, callback=None):
        steps = np.arange(self.ddpm_num_timesteps) if use_original_steps else self.ddim_timesteps
        steps = steps[:t_start]
        total = steps.shape[0]
        x = x_latent
        for i, step in enumerate(np.flip(steps)):
            ts = torch.full((x_latent.shape[0],), int(step), device=x_latent.device, dtype=torch.long)
            x, _ = self._step(x, cond, ts, index=total - i - 1, use_original_steps=use_original_steps,
                              unconditional_guidance_scale=unconditional_guidance_scale,
                              unconditional_conditioning=unconditional_conditioning)
            if callback:
                callback(i)
        return x


class MKDDIMSampler(DDIMSampler):
    """diffmk/cddim.py:5-100 — grad-enabled twins of p_sample_ddim / decode."""

    def denoising_step(self, x, c, t, index, **kw):
        return self._step(x, c, t, index, **kw)

    def reconstruct(self, x_latent, cond, t_start, unconditional_guidance_scale=1.0,
                    unconditional_conditioning=None, use_original_steps=False, callback=None):
        return self._reverse(x_latent, cond, t_start, unconditional_guidance_scale,
                             unconditional_conditioning, use_original_steps, callback)
