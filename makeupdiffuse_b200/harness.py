"""Test-harness I/O of the reference's inference run, on the B200 path.

What ``runs/test.py`` drives per batch is ``test_step`` -> ``log_results`` -> ``save_local``
(``diffmk/diffusion_makeup.py:332-411``) over the pair list of ``TestFixed_Dataset`` (``diffdata/datasets.py:728-784``).
This module mirrors those three pieces for a batch that already holds tensors (decoding image files, face parsing
masks and landmarks belongs to the teacher pipeline, which stays out of scope — SURVEY.md §2):

* ``read_pair_list`` / ``pair_basenames`` — ``datasets.py:738-741, 760-764`` ("<non-makeup> <makeup>" per line,
  ``img_name = '<src>&<ref>'``);
* ``log_results`` — ``diffusion_makeup.py:360-411``: x_p -> ``get_z`` -> reconstruction, control_src / control_ref,
  ground_truth, one-step x_0 prediction at a random t >= t_min ("sample_ddmp"), DDIM samples, guided DDIM samples, and
  the ``test_pairs`` bookkeeping (``:376-381``), and the text panel ``conditioning`` (``:372``, upstream
  ``ldm.util.log_txt_as_img``: the image names drawn on a white 256 x 256 canvas — ``txt_panels`` below; upstream loads
  ``font/DejaVuSans.ttf`` from the ControlNet checkout, here a TTF path can be passed and PIL's built-in font is the default,
  so the glyph pixels may differ while key, shape, value range and line wrapping are the reference's);
* ``sample_log`` — upstream ``ControlLDM.sample_log`` as called at ``:393-408`` (shape from ``c_concat``);
* ``test_step`` / ``save_local`` — ``:332-358``: clamp, ``make_grid(nrow = number of panels)``, rescale, HWC, uint8 and
  PNG.  The grid / rescale / uint8 conversion is one CUDA pass (``mkd_image_grid_u8``), byte-identical to the
  reference's torchvision + numpy sequence; only the uint8 grid crosses PCIe.
* ``write_test_pairs`` — ``:326-330``.
"""
from __future__ import annotations

import os

import torch

from . import ops
from .sampler import B200DDIMSampler


def read_pair_list(path):
    """lines "<non-makeup image> <makeup image>" -> (non_makeup_names, makeup_names)   (datasets.py:738-741)"""
    with open(path, "r") as f:
        rows = [ln.strip().split(" ") for ln in f.readlines() if ln.strip()]
    return [r[0] for r in rows], [r[1] for r in rows]


def pair_basenames(non_makeup_names, makeup_names):
    """img_name of sample i: '<basename of source>&<basename of reference>'   (datasets.py:760-764)"""
    base = lambda n: os.path.basename(n).split(".")[0]  # noqa: E731
    return ["%s&%s" % (base(s), base(r)) for s, r in zip(non_makeup_names, makeup_names)]


def test_pair_rows(batch_idx, img_names):
    """diffusion_makeup.py:376-381"""
    return [["%04d-%d" % (batch_idx, i + 1), "non-makeup/%s.png" % n.split("&")[0], "makeup/%s.png" % n.split("&")[1]]
            for i, n in enumerate(img_names)]


def write_test_pairs(path, test_pairs):
    """diffusion_makeup.py:326-330"""
    with open(path, "w") as f:
        for p in test_pairs:
            f.write("%s %s %s\n" % (p[0], p[1], p[2]))


def txt_panels(wh, captions, size=10, font_path=None):
    """upstream ``ldm.util.log_txt_as_img`` as called at diffusion_makeup.py:372 with ((256, 256), img_names, size=16): one white
    RGB canvas of ``wh`` = (width, height) per caption, the caption in black from the top-left corner, wrapped every
    ``int(40 * width / 256)`` characters; returns [B, 3, H, W] fp32 in [-1, 1] on the host (pure host-side drawing)."""
    import numpy as np
    from PIL import Image, ImageDraw, ImageFont
    font = ImageFont.truetype(font_path, size=size) if font_path else ImageFont.load_default(size=size)
    nc = int(40 * (wh[0] / 256))
    out = []
    for cap in captions:
        canvas = Image.new("RGB", tuple(wh), color="white")
        lines = "\n".join(cap[i:i + nc] for i in range(0, len(cap), nc))
        try:
            ImageDraw.Draw(canvas).text((0, 0), lines, fill="black", font=font)
        except UnicodeEncodeError:  # upstream skips captions its font cannot encode and keeps the blank canvas
            pass
        out.append(np.array(canvas).transpose(2, 0, 1) / 127.5 - 1.0)
    return torch.tensor(np.stack(out), dtype=torch.float32)


def sample_log(model, cond, batch_size, ddim, ddim_steps, sampler=None, **kwargs):
    """upstream ControlLDM.sample_log: DDIM over shape (4, h / 8, w / 8) taken from the hint"""
    assert ddim, "the reference only samples with DDIM (ddim_steps is set in its configs)"
    sampler = sampler or B200DDIMSampler(model)
    _, _, h, w = cond["c_concat"][0].shape
    return sampler.sample(ddim_steps, batch_size, (4, h // 8, w // 8), cond, verbose=False, **kwargs)


@torch.no_grad()
def log_results(model, batch, batch_idx=0, *, ddim_steps=50, ddim_eta=0.0, sample=True, unconditional_guidance_scale=9.0,
                t_min=0, test_pairs=None, sampler=None, generator=None, font_path=None):
    """diffusion_makeup.py:360-411.  ``batch``: 'pgt_sr' [B,3,H,W] in [-1,1] (the teacher output x_p), 'src_img' /
    'ref_img' [B,3,H,W] in [0,1], 'c_crossattn' [B,77,768] or 'tokens' [B,77] or 'txt' (list of prompts), optional
    'img_name'.  Returns the reference's dict of image batches (device tensors in about [-1, 1])."""
    dev = model.device
    pgt_sr = batch["pgt_sr"].to(dev, torch.float32)
    c_cat = model.assemble_hint(batch["src_img"].to(dev, torch.float32), batch["ref_img"].to(dev, torch.float32))
    if "c_crossattn" in batch:
        c = batch["c_crossattn"].to(dev, torch.float32)
    else:
        c = model.get_learned_conditioning(batch["tokens"] if "tokens" in batch else batch["txt"])
    log = {}
    zn = torch.randn(pgt_sr.shape[0], 4, pgt_sr.shape[2] // 8, pgt_sr.shape[3] // 8, device=dev, generator=generator)
    z = model.get_z(pgt_sr, zn)  # posterior sample (makeup_diffuse.py:37-40) with the noise drawn here
    log["reconstruction"] = model.decode_first_stage(z)
    src, ref = torch.chunk(c_cat, 2, dim=1)
    log["control_src"] = src * 2.0 - 1.0
    log["control_ref"] = ref * 2.0 - 1.0
    if "img_name" in batch:  # diffusion_makeup.py:372 (the reference's batches always carry img_name)
        log["conditioning"] = txt_panels((256, 256), batch["img_name"], size=16, font_path=font_path).to(dev)
    log["ground_truth"] = pgt_sr
    if test_pairs is not None and "img_name" in batch:
        test_pairs.extend(test_pair_rows(batch_idx, batch["img_name"]))
    b = z.shape[0]
    cond = {"c_concat": [c_cat], "c_crossattn": [c]}
    t = torch.randint(t_min, model.num_timesteps, (b,), device=dev, generator=generator).long()
    noise = torch.randn(z.shape, device=dev, generator=generator)
    x_noisy = model.q_sample(x_start=z, t=t, noise=noise)
    _, x_recon = model.apply_model(x_noisy, t, cond, return_all=True)
    log["sample_ddmp"] = model.decode_first_stage(x_recon)
    sampler = sampler or B200DDIMSampler(model)
    if sample:
        samples, _ = sample_log(model, cond, b, True, ddim_steps, sampler=sampler, eta=ddim_eta)
        log["samples"] = model.decode_first_stage(samples)
    if unconditional_guidance_scale > 1.0:
        uc_full = {"c_concat": [c_cat], "c_crossattn": [model.get_unconditional_conditioning(b)]}
        samples_cfg, _ = sample_log(model, cond, b, True, ddim_steps, sampler=sampler, eta=ddim_eta,
                                    unconditional_guidance_scale=unconditional_guidance_scale,
                                    unconditional_conditioning=uc_full)
        log[f"samples_cfg_scale_{unconditional_guidance_scale:.2f}"] = model.decode_first_stage(samples_cfg)
    return log


def save_local(images, batch_idx, root, rescale=True, clamp=True, write_png=True):
    """diffusion_makeup.py:344-358 (+ the clamp of test_step, :340-341).  Returns {key: uint8 HWC grid on the host}."""
    nrow = len(images)  # (sic: the reference uses the number of panels as make_grid's nrow)
    out = {}
    for k, v in images.items():
        grid = ops.image_grid_u8(v.float(), nrow=nrow, padding=2, clamp=clamp, rescale=rescale).cpu().numpy()
        out[k] = grid
        if write_png:
            from PIL import Image
            path = os.path.join(root, "{}_{:04}.png".format(k, batch_idx))
            os.makedirs(os.path.split(path)[0], exist_ok=True)
            Image.fromarray(grid).save(path)
    return out


def test_step(model, batch, batch_idx, root, **kw):
    """diffusion_makeup.py:332-342"""
    save_kw = {k: kw.pop(k) for k in ("rescale", "clamp", "write_png") if k in kw}
    images = log_results(model, batch, batch_idx, **kw)
    return save_local(images, batch_idx, root, **save_kw)


test_pair_rows.__test__ = False  # (helpers named test_* are not pytest tests)
test_step.__test__ = False
