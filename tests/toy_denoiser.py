"""TEST-ONLY closed-form denoiser behind the sampler surface (B1/B2 of SURVEY.md §8(b)): everything the samplers read
off ``model`` (cddim.py:16,39,42,51-54) without the networks.  Shared by tests/golden/make_golden_ref_sampler.py (which
drives the REFERENCE's diffmk/cddim.py with it) and tests/test_golden.py (oracle and B200 samplers)."""
import numpy as np
import torch


class ToyDenoiser:
    parameterization = "eps"
    num_timesteps = 1000

    def __init__(self, device="cpu"):
        self.device = torch.device(device)
        # register_schedule "linear" with yaml:4-8 (linear_start 0.00085, linear_end 0.012, 1000 steps), stored fp32
        betas = np.linspace(0.00085 ** 0.5, 0.012 ** 0.5, 1000, dtype=np.float64) ** 2
        ac = np.cumprod(1.0 - betas)
        f = lambda v: torch.tensor(v, dtype=torch.float32, device=self.device)  # noqa: E731
        self.betas = f(betas)
        self.alphas_cumprod = f(ac)
        self.alphas_cumprod_prev = f(np.append(1.0, ac[:-1]))
        self.sqrt_one_minus_alphas_cumprod = f(np.sqrt(1.0 - ac))
        self.calls = []  # batch size of every apply_model call

    @staticmethod
    def _leaves(c):
        if isinstance(c, dict):
            return c["c_crossattn"][0], (c["c_concat"][0] if c.get("c_concat") else None)
        if isinstance(c, list):
            return c[0], (c[1] if len(c) > 1 else None)
        return c, None

    def apply_model(self, x, t, c):
        """eps depends on x, on t, on the context row and on the hint row of the SAME batch index, so a wrong CFG
        batching order, a wrong timestep or a mixed-up batch row all change the result."""
        self.calls.append(int(x.shape[0]))
        ctx, hint = self._leaves(c)
        b = x.shape[0]
        k = ctx.float().mean(dim=tuple(range(1, ctx.dim()))).view(b, 1, 1, 1)
        h = hint.float().mean(dim=tuple(range(1, hint.dim()))).view(b, 1, 1, 1) if hint is not None else 0.0
        tt = t.float().view(b, 1, 1, 1) / 1000.0
        return torch.tanh(0.7 * x.float() + 2.0 * k) * (0.5 + tt) + 0.25 * h - 0.1 * tt


def toy_inputs(device="cpu", B=3, h=8):
    g = torch.Generator().manual_seed(20261018)
    d = {"ctx": torch.randn(B, 7, 16, generator=g), "uc_ctx": torch.randn(B, 7, 16, generator=g),
         "hint": torch.rand(B, 6, 8 * h, 8 * h, generator=g), "x_T": torch.randn(B, 4, h, h, generator=g)}
    return {k: v.to(device) for k, v in d.items()}
