"""``B200DDIMSampler``: drop-in for ``MKDDIMSampler`` (diffmk/cddim.py:5-100) and the upstream ``DDIMSampler`` it
extends — same constructor, same ``make_schedule`` / ``sample`` / ``ddim_sampling`` / ``p_sample_ddim`` /
``denoising_step`` / ``reconstruct`` / ``decode`` / ``stochastic_encode`` signatures and return values
(SURVEY.md §8(b) level B1).

Per step it runs ``model.apply_model`` (one CUDA-graph replay of the fused UNet+ControlNet kernels when
``use_cuda_graph`` is on) followed by ONE fused kernel for the whole x_t -> x_{t-1} update including the
classifier-free-guidance combine (cddim.py:39-40, 56-78), instead of ~12 elementwise launches + 4 ``torch.full``.
The RNG stream is consumed exactly like the reference: one ``randn(x.shape)`` per step even when sigma == 0
(cddim.py:75).  Unlike upstream nothing is force-moved to "cuda": buffers follow ``model.device``.
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops


def _uniform_ddim_timesteps(num_ddim, num_ddpm):
    c = num_ddpm // num_ddim
    return np.asarray(list(range(0, num_ddpm, c))) + 1


def _cat_uncond_first(uc, c):
    """cddim.py:20-38 — [uncond; cond] along batch for every tensor leaf"""
    if isinstance(c, dict):
        assert isinstance(uc, dict)
        return {k: ([torch.cat([uc[k][i], c[k][i]]) for i in range(len(c[k]))] if isinstance(c[k], list)
                    else torch.cat([uc[k], c[k]])) for k in c}
    if isinstance(c, list):
        assert isinstance(uc, list)
        return [torch.cat([uc[i], c[i]]) for i in range(len(c))]
    return torch.cat([uc, c])


class B200DDIMSampler:
    def __init__(self, model, schedule="linear", use_cuda_graph=True, **kwargs):
        self.model = model
        self.ddpm_num_timesteps = model.num_timesteps
        self.schedule = schedule
        self.use_cuda_graph = use_cuda_graph
        self._cfg_cache = None
        self._graph = None
        self.graph_launches = 0  # kernels executed through graph replays (they bypass the library's launch counter)

    # ---- schedule (upstream make_schedule / make_ddim_sampling_parameters) -----------------------------------
    def make_schedule(self, ddim_num_steps, ddim_discretize="uniform", ddim_eta=0.0, verbose=True):
        if ddim_discretize != "uniform":
            raise NotImplementedError(f'There is no ddim discretization method called "{ddim_discretize}"')
        m = self.model
        self.ddim_timesteps = _uniform_ddim_timesteps(ddim_num_steps, self.ddpm_num_timesteps)
        ac = m.alphas_cumprod.detach().float()
        assert ac.shape[0] == self.ddpm_num_timesteps, "alphas have to be defined for each timestep"
        dev = m.device
        self.betas = m.betas.float()
        self.alphas_cumprod = ac
        self.alphas_cumprod_prev = m.alphas_cumprod_prev.float()
        self.sqrt_alphas_cumprod = ac.sqrt()
        self.sqrt_one_minus_alphas_cumprod = (1.0 - ac).sqrt()
        acn = ac.cpu().numpy()
        a = acn[self.ddim_timesteps]
        a_prev = np.asarray([acn[0]] + acn[self.ddim_timesteps[:-1]].tolist())
        sig = ddim_eta * np.sqrt((1 - a_prev) / (1 - a) * (1 - a / a_prev))
        f32 = lambda v: torch.as_tensor(np.asarray(v), dtype=torch.float32).to(dev)  # noqa: E731
        self.ddim_sigmas = f32(sig)
        self.ddim_alphas = f32(a)
        self.ddim_alphas_prev = a_prev
        self.ddim_sqrt_one_minus_alphas = f32(np.sqrt(1.0 - a))
        acp = self.alphas_cumprod_prev
        self.ddim_sigmas_for_original_num_steps = ddim_eta * torch.sqrt((1 - acp) / (1 - ac) * (1 - ac / acp))
        # host copies of the four per-step coefficients, rounded exactly like the reference's fp32 torch ops
        self._coef_ddim = self._coefficients(self.ddim_alphas, self.ddim_alphas_prev, self.ddim_sqrt_one_minus_alphas,
                                             self.ddim_sigmas)
        self._coef_orig = None

    @staticmethod
    def _coefficients(A, AP, S1, SG):
        """per index: (sqrt(1-a_t), sqrt(a_t), sqrt(a_prev), sqrt(1 - a_prev - sigma^2), sigma) as fp32 values
        (cddim.py:56-59 builds fp32 tensors with torch.full; :63,74,78 take fp32 sqrt)"""
        f = lambda v: torch.as_tensor(np.asarray(v.cpu() if torch.is_tensor(v) else v), dtype=torch.float32)  # noqa: E731
        a, ap, s1, sg = f(A), f(AP), f(S1), f(SG)
        return torch.stack([s1, a.sqrt(), ap.sqrt(), (1.0 - ap - sg ** 2).sqrt(), sg], 1).tolist()

    # ---- sampling from noise ----------------------------------------------------------------------------------
    def sample(self, S, batch_size, shape, conditioning=None, callback=None, normals_sequence=None, img_callback=None,
               quantize_x0=False, eta=0.0, mask=None, x0=None, temperature=1.0, noise_dropout=0.0, score_corrector=None,
               corrector_kwargs=None, verbose=True, x_T=None, log_every_t=100, unconditional_guidance_scale=1.0,
               unconditional_conditioning=None, dynamic_threshold=None, ucg_schedule=None, **kwargs):
        self.make_schedule(ddim_num_steps=S, ddim_eta=eta, verbose=verbose)
        C, H, W = shape
        return self.ddim_sampling(conditioning, (batch_size, C, H, W), callback=callback, img_callback=img_callback,
                                  quantize_denoised=quantize_x0, mask=mask, x0=x0, noise_dropout=noise_dropout,
                                  temperature=temperature, score_corrector=score_corrector,
                                  corrector_kwargs=corrector_kwargs, x_T=x_T, log_every_t=log_every_t,
                                  unconditional_guidance_scale=unconditional_guidance_scale,
                                  unconditional_conditioning=unconditional_conditioning,
                                  dynamic_threshold=dynamic_threshold)

    @torch.no_grad()
    def ddim_sampling(self, cond, shape, x_T=None, callback=None, img_callback=None, quantize_denoised=False, mask=None,
                      x0=None, log_every_t=100, temperature=1.0, noise_dropout=0.0, score_corrector=None,
                      corrector_kwargs=None, unconditional_guidance_scale=1.0, unconditional_conditioning=None,
                      dynamic_threshold=None, **kwargs):
        dev = self.model.device
        b = shape[0]
        steps = np.flip(self.ddim_timesteps)
        self.begin_loop(steps)
        img = torch.randn(shape, device=dev) if x_T is None else x_T
        inter = {"x_inter": [img], "pred_x0": [img]}
        total = steps.shape[0]
        for i, step in enumerate(steps):
            index = total - i - 1
            ts = torch.full((b,), int(step), device=dev, dtype=torch.long)
            if mask is not None:
                img = self.model.q_sample(x0, ts) * mask + (1.0 - mask) * img
            img, pred_x0 = self.denoising_step(
                img, cond, ts, index=index, quantize_denoised=quantize_denoised, temperature=temperature,
                noise_dropout=noise_dropout, score_corrector=score_corrector, corrector_kwargs=corrector_kwargs,
                unconditional_guidance_scale=unconditional_guidance_scale,
                unconditional_conditioning=unconditional_conditioning, dynamic_threshold=dynamic_threshold,
                t_value=int(step))
            if callback:
                callback(i)
            if img_callback:
                img_callback(pred_x0, i)
            if index % log_every_t == 0 or index == total - 1:
                inter["x_inter"].append(img)
                inter["pred_x0"].append(pred_x0)
        return img, inter

    def begin_loop(self, steps=None):
        """A sampling loop starts: whatever was hoisted for an earlier cond (the model's hint features and K/V, the
        doubled CFG cond) is dropped, so the loop reads its conditioning tensors as they are NOW — also when the caller
        refilled the same tensors in a way torch's version counter does not see.  Loops driven from outside
        (makeupdiffuse_b200.dist.sample_sharded, a caller's own loop over denoising_step) call this themselves.
        steps: the loop's timesteps, when known: the model computes the timestep embeddings of all of them at once, and
        denoising_step(..., t_value=step) selects one per step."""
        self._cfg_cache = None
        inv = getattr(self.model, "invalidate_cond_cache", None)
        if inv is not None:
            inv()
        pre = getattr(self.model, "precompute_time_embeddings", None)
        if pre is not None and steps is not None and len(steps):
            pre([int(v) for v in steps])

    def p_sample_ddim(self, *a, **k):
        with torch.no_grad():
            return self.denoising_step(*a, **k)

    def stochastic_encode(self, x0, t, use_original_steps=False, noise=None):
        sa = self.sqrt_alphas_cumprod if use_original_steps else self.ddim_alphas.sqrt()
        s1 = self.sqrt_one_minus_alphas_cumprod if use_original_steps else self.ddim_sqrt_one_minus_alphas
        noise = torch.randn_like(x0) if noise is None else noise
        ex = lambda v: v.gather(-1, t).reshape(-1, 1, 1, 1)  # noqa: E731
        return ex(sa) * x0 + ex(s1) * noise

    def decode(self, x_latent, cond, t_start, **kw):
        with torch.no_grad():
            return self.reconstruct(x_latent, cond, t_start, **kw)

    # ---- the model call, optionally as a CUDA-graph replay -------------------------------------------------------
    def _eps(self, x, t, c, t_value=None):
        """model.apply_model(x, t, c); with use_cuda_graph the ~10^3 kernel launches of one UNet+ControlNet step are
        captured once per (cond, batch shape) and replayed (launch-bound otherwise: SURVEY.md §7 hard part 6).
        t_value: the loop's timestep as a host int when every row of ``t`` holds it (the sampler's own loops): the model
        then takes the ResBlock timestep embeddings from the table computed at the start of the loop (begin_loop(steps))
        instead of running the two embedding MLPs again."""
        set_step = getattr(self.model, "set_step", None)
        if set_step is None or t_value is None:
            return self._eps_call(x, t, c, False)
        try:
            return self._eps_call(x, t, c, set_step(t_value, x.shape[0]))
        finally:
            set_step(None)

    def _eps_call(self, x, t, c, emb_selected):
        prepare = getattr(self.model, "_prepare", None)
        if not (self.use_cuda_graph and x.is_cuda and prepare is not None):
            return self.model.apply_model(x, t, c)
        from . import _lib
        g = self._graph
        # the graph only depends on shapes: every per-cond tensor it reads (hint features, cross-attention K/V) lives
        # in a static arena buffer that model._prepare() refills in place when the cond changes
        m = self.model
        nets = [getattr(m, "control_model", None), getattr(getattr(m, "model", None), "diffusion_model", None)]
        key = (tuple(x.shape), c["c_concat"] is None, tuple(tuple(v.shape) for v in c["c_crossattn"]), id(m),
               getattr(m, "only_mid_control", False), tuple(getattr(m, "control_scales", ())),
               # a graph holds raw pointers and the launch structure of the moment it was captured: reloaded weights
               # (new tensors), another stream layout or another GroupNorm path need a new capture
               getattr(m, "_weights_epoch", 0), getattr(m, "concurrent", None), getattr(m, "grouped", None), emb_selected,
               tuple((getattr(n, "fused_gn_stats", None), getattr(n, "fuse_gn_tail", None)) for n in nets))
        if g is None or g["key"] != key:
            sx, st = x.clone(), t.clone()
            self.model.apply_model(sx, st, c)  # warm-up: allocates every static buffer, fills the cond cache
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            n0 = _lib.load().mkd_launch_count()
            with torch.cuda.graph(graph):
                out = self.model.apply_model(sx, st, c)
            g = self._graph = {"key": key, "graph": graph, "x": sx, "t": st, "out": out,
                               "launches": _lib.load().mkd_launch_count() - n0}
        # no-op when the cond is unchanged; otherwise recomputes the hoisted tensors eagerly (in the stacked form too when the
        # captured evaluation runs the two trunks as one network)
        ug = getattr(m, "_use_grouped", None)
        if ug is not None:
            prepare(c, c["c_concat"] is not None and ug(x.shape[0], x.shape[2], x.shape[3]))
        else:
            prepare(c)
        g["x"].copy_(x)
        g["t"].copy_(t)
        g["graph"].replay()
        self.graph_launches += g["launches"]
        return g["out"]

    # ---- one x_t -> x_{t-1} update: diffmk/cddim.py:9-79 ---------------------------------------------------------
    def denoising_step(self, x, c, t, index, repeat_noise=False, use_original_steps=False, quantize_denoised=False,
                       temperature=1.0, noise_dropout=0.0, score_corrector=None, corrector_kwargs=None,
                       unconditional_guidance_scale=1.0, unconditional_conditioning=None, dynamic_threshold=None,
                       out=None, peer_ptrs=None, t_value=None):
        """``out`` (extension, optional): tensor that receives x_prev — e.g. this rank's slice of an all-gather
        buffer, so the last update of a sharded run lands directly where the collective reads it.
        ``peer_ptrs`` (extension, optional): device pointers of that same slice inside EVERY rank's gather buffer (peer
        memory); the update kernel then stores x_prev to all of them — the all-gather fused into the kernel.
        ``t_value`` (extension, optional): the timestep as a host int when every row of ``t`` holds it (see _eps)."""
        m = self.model
        if m.parameterization != "eps":
            raise NotImplementedError("B200 path implements the yaml's parameterization: eps (yaml:50)")
        if score_corrector is not None or quantize_denoised:
            raise NotImplementedError("score_corrector / quantize_denoised are unused by the reference path")
        if dynamic_threshold is not None:
            raise NotImplementedError()
        b = x.shape[0]
        x = x.float().contiguous()
        cfg = not (unconditional_conditioning is None or unconditional_guidance_scale == 1.0)
        if not cfg:
            e = self._eps(x, t, c, t_value)
        else:
            cc = self._cfg_cache  # the doubled cond is step-invariant: build it once, not 50 times
            if cc is None or cc[0] is not c or cc[1] is not unconditional_conditioning:
                cc = self._cfg_cache = (c, unconditional_conditioning, _cat_uncond_first(unconditional_conditioning, c))
            e = self._eps(torch.cat([x] * 2), torch.cat([t] * 2), cc[2], t_value)
        if use_original_steps:
            if self._coef_orig is None:
                # cddim.py:54 reads the sigma table off the MODEL; without that attribute the reference raises
                # AttributeError, and so does this
                self._coef_orig = self._coefficients(m.alphas_cumprod, m.alphas_cumprod_prev,
                                                     m.sqrt_one_minus_alphas_cumprod,
                                                     m.ddim_sigmas_for_original_num_steps)
            s1m, sq_at, sq_ap, dirc, sigma = self._coef_orig[index]
        else:
            s1m, sq_at, sq_ap, dirc, sigma = self._coef_ddim[index]
        # the reference draws the noise every step, also when sigma == 0 (cddim.py:75): keep the RNG stream in step
        shape = (1, *x.shape[1:]) if repeat_noise else x.shape
        draw = torch.randn(shape, device=x.device)
        noise = None
        if sigma != 0.0:
            noise = draw.repeat(b, *((1,) * (x.dim() - 1))) if repeat_noise else draw
            if noise_dropout > 0.0:
                # dropout acts on sigma*noise*temperature in the reference; it commutes with the scalar factors
                noise = torch.nn.functional.dropout(noise, p=noise_dropout)
            noise = noise.contiguous()
        x_prev, pred_x0 = (torch.empty_like(x) if out is None else out), torch.empty_like(x)
        ops.ddim_update(x, e.contiguous(), x_prev, sqrt_one_minus_at=s1m, sqrt_at=sq_at, sqrt_a_prev=sq_ap, dir_coef=dirc,
                        sigma_t=sigma, temperature=temperature, noise=noise, pred_x0=pred_x0,
                        cfg_scale=float(unconditional_guidance_scale) if cfg else None, peer_ptrs=peer_ptrs)
        return x_prev, pred_x0

    # ---- truncated reverse loop from a caller-supplied x_t: diffmk/cddim.py:81-100 -------------------------------
    def reconstruct(self, x_latent, cond, t_start, unconditional_guidance_scale=1.0, unconditional_conditioning=None,
                    use_original_steps=False, callback=None):
        steps = np.arange(self.ddpm_num_timesteps) if use_original_steps else self.ddim_timesteps
        steps = steps[:t_start]
        total = steps.shape[0]
        self.begin_loop(steps)
        x = x_latent
        for i, step in enumerate(np.flip(steps)):
            ts = torch.full((x_latent.shape[0],), int(step), device=x_latent.device, dtype=torch.long)
            x, _ = self.denoising_step(x, cond, ts, index=total - i - 1, use_original_steps=use_original_steps,
                                       unconditional_guidance_scale=unconditional_guidance_scale,
                                       unconditional_conditioning=unconditional_conditioning, t_value=int(step))
            if callback:
                callback(i)
        return x
