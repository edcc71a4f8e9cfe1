"""Oracle stand-in for the LatentDiffusion/ControlLDM object the sampler holds as ``self.model``.

Only the surface the sampler touches (SURVEY.md §8(b) level B1/B2): ``apply_model``, ``parameterization``,
the schedule buffers, ``num_timesteps``, ``betas``, ``device``; plus ``q_sample`` /
``predict_start_from_noise`` used at the x_p entry (``diffmk/diffusion_makeup.py:384-389``).

Test infrastructure (see oracle/__init__.py).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from .nets import ControlNet, ControlledUnetModel


def linear_beta_alphas_cumprod(timesteps=1000, linear_start=0.00085, linear_end=0.0120) -> np.ndarray:
    """yaml:4-8 with upstream's ``linear`` schedule: beta = linspace(sqrt(b0), sqrt(b1), T, f64)**2;
    returns cumprod(1-beta) in float64 (callers store fp32)."""
    betas = np.linspace(linear_start ** 0.5, linear_end ** 0.5, timesteps, dtype=np.float64) ** 2
    return np.cumprod(1.0 - betas, axis=0)


class _DiffusionWrapper(nn.Module):
    def __init__(self, unet):
        super().__init__()
        self.diffusion_model = unet


class OracleControlLDM(nn.Module):
    """State-dict prefixes match upstream: ``control_model.*`` and ``model.diffusion_model.*``."""

    def __init__(self, control_params=None, unet_params=None, timesteps=1000, linear_start=0.00085,
                 linear_end=0.0120, parameterization="eps", only_mid_control=False, scale_factor=0.18215):
        super().__init__()
        self.control_model = ControlNet(**(control_params or {}))
        self.model = _DiffusionWrapper(ControlledUnetModel(**(unet_params or {})))
        self.control_scales = [1.0] * 13
        self.only_mid_control = only_mid_control
        self.parameterization = parameterization
        self.scale_factor = scale_factor
        self.num_timesteps = int(timesteps)
        ac = linear_beta_alphas_cumprod(timesteps, linear_start, linear_end)
        betas = np.linspace(linear_start ** 0.5, linear_end ** 0.5, timesteps, dtype=np.float64) ** 2
        f32 = lambda a: torch.tensor(a, dtype=torch.float32)  # noqa: E731
        self.register_buffer("betas", f32(betas))
        self.register_buffer("alphas_cumprod", f32(ac))
        self.register_buffer("alphas_cumprod_prev", f32(np.append(1.0, ac[:-1])))
        self.register_buffer("sqrt_alphas_cumprod", f32(np.sqrt(ac)))
        self.register_buffer("sqrt_one_minus_alphas_cumprod", f32(np.sqrt(1.0 - ac)))
        self.register_buffer("sqrt_recip_alphas_cumprod", f32(np.sqrt(1.0 / ac)))
        self.register_buffer("sqrt_recipm1_alphas_cumprod", f32(np.sqrt(1.0 / ac - 1)))

    @property
    def device(self):
        return self.betas.device

    # diffmk/makeup_diffuse.py:152-170 (== upstream ControlLDM.apply_model)
    def apply_model(self, x_noisy, t, cond, return_all=False, *args, **kwargs):
        assert isinstance(cond, dict)
        unet = self.model.diffusion_model
        cond_txt = torch.cat(cond["c_crossattn"], 1)
        if cond["c_concat"] is None:
            eps = unet(x=x_noisy, timesteps=t, context=cond_txt, control=None,
                       only_mid_control=self.only_mid_control)
        else:
            control = self.control_model(x=x_noisy, hint=torch.cat(cond["c_concat"], 1), timesteps=t,
                                         context=cond_txt)
            control = [c * s for c, s in zip(control, self.control_scales)]
            eps = unet(x=x_noisy, timesteps=t, context=cond_txt, control=control,
                       only_mid_control=self.only_mid_control)
        if not return_all:
            return eps
        return eps, self.predict_start_from_noise(x_noisy, t, eps)

    @staticmethod
    def _extract(a, t, x):
        return a.gather(-1, t).reshape(t.shape[0], *((1,) * (x.dim() - 1)))

    def q_sample(self, x_start, t, noise=None):
        noise = torch.randn_like(x_start) if noise is None else noise
        return (self._extract(self.sqrt_alphas_cumprod, t, x_start) * x_start +
                self._extract(self.sqrt_one_minus_alphas_cumprod, t, x_start) * noise)

    def predict_start_from_noise(self, x_t, t, noise):
        return (self._extract(self.sqrt_recip_alphas_cumprod, t, x_t) * x_t -
                self._extract(self.sqrt_recipm1_alphas_cumprod, t, x_t) * noise)
