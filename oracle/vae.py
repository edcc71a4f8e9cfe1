"""Oracle restatement of the first-stage DECODER the reference runs right after sampling:
``decode_first_stage(z)`` = ``first_stage_model.decode(z / scale_factor)`` (``diffmk/makeups.py:260-262``; call sites
``diffmk/diffusion_makeup.py:389,396,409``), with ``first_stage_model`` = upstream ``ldm.models.autoencoder.AutoencoderKL``
configured by ``diffmodels/base_diffusion_makeup.yaml:86-105`` (``embed_dim 4, z_channels 4, ch 128, ch_mult 1,2,4,4,
num_res_blocks 2, attn_resolutions [], out_ch 3``):

    decode(z) = Decoder(post_quant_conv(z))

The module tree / op order are those of upstream ``ldm.modules.diffusionmodules.model.Decoder`` (NOT vendored in the
reference), restated in plain fp32 PyTorch with upstream parameter names (``post_quant_conv``, ``decoder.conv_in``,
``decoder.mid.block_1|attn_1|block_2``, ``decoder.up.{level}.block.{i}``, ``decoder.up.{level}.upsample.conv``,
``decoder.norm_out``, ``decoder.conv_out``) so that the ``first_stage_model.*`` keys of an SD-1.5 checkpoint load.
Structural pin (tests/test_vae.py): 49 490 179 decoder parameters + 20 in ``post_quant_conv`` — the public SD VAE
(83 653 863 = encoder 34 163 592 + quant_conv 72 + these).

PARITY UNPINNED by the reference like the rest of the oracle (oracle/__init__.py).  Test infrastructure.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F


def _norm(c):
    return nn.GroupNorm(32, c, eps=1e-6, affine=True)


def _swish(x):
    return x * torch.sigmoid(x)


class ResnetBlock(nn.Module):
    """x + conv2(swish(norm2(conv1(swish(norm1(x))))));  1x1 ``nin_shortcut`` on x when the width changes (temb unused)"""

    def __init__(self, cin, cout):
        super().__init__()
        self.norm1, self.conv1 = _norm(cin), nn.Conv2d(cin, cout, 3, padding=1)
        self.norm2, self.conv2 = _norm(cout), nn.Conv2d(cout, cout, 3, padding=1)
        if cin != cout:
            self.nin_shortcut = nn.Conv2d(cin, cout, 1)

    def forward(self, x):
        h = self.conv1(_swish(self.norm1(x)))
        h = self.conv2(_swish(self.norm2(h)))
        return (self.nin_shortcut(x) if hasattr(self, "nin_shortcut") else x) + h


class AttnBlock(nn.Module):
    """single-head self-attention over all pixels: x + proj_out(softmax(q^T k / sqrt(C)) v)"""

    def __init__(self, c):
        super().__init__()
        self.norm = _norm(c)
        self.q, self.k, self.v, self.proj_out = (nn.Conv2d(c, c, 1) for _ in range(4))

    def forward(self, x):
        h = self.norm(x)
        b, c, hh, ww = h.shape
        q = self.q(h).reshape(b, c, hh * ww).permute(0, 2, 1)   # b, hw, c
        k = self.k(h).reshape(b, c, hh * ww)                    # b, c, hw
        v = self.v(h).reshape(b, c, hh * ww)
        w = torch.bmm(q, k) * (int(c) ** -0.5)
        w = F.softmax(w, dim=2)
        h = torch.bmm(v, w.permute(0, 2, 1)).reshape(b, c, hh, ww)
        return x + self.proj_out(h)


class Upsample(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, padding=1)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))


class Decoder(nn.Module):
    def __init__(self, ch=128, out_ch=3, ch_mult=(1, 2, 4, 4), num_res_blocks=2, z_channels=4, **unused):
        super().__init__()
        self.num_resolutions, self.num_res_blocks = len(ch_mult), num_res_blocks
        block_in = ch * ch_mult[-1]
        self.conv_in = nn.Conv2d(z_channels, block_in, 3, padding=1)
        self.mid = nn.Module()
        self.mid.block_1 = ResnetBlock(block_in, block_in)
        self.mid.attn_1 = AttnBlock(block_in)
        self.mid.block_2 = ResnetBlock(block_in, block_in)
        self.up = nn.ModuleList()
        for i_level in reversed(range(self.num_resolutions)):
            block_out = ch * ch_mult[i_level]
            up = nn.Module()
            up.block = nn.ModuleList()
            for _ in range(num_res_blocks + 1):
                up.block.append(ResnetBlock(block_in, block_out))
                block_in = block_out
            if i_level != 0:
                up.upsample = Upsample(block_in)
            self.up.insert(0, up)  # up[0] is the full-resolution level, like upstream
        self.norm_out = _norm(block_in)
        self.conv_out = nn.Conv2d(block_in, out_ch, 3, padding=1)

    def forward(self, z):
        h = self.conv_in(z)
        h = self.mid.block_2(self.mid.attn_1(self.mid.block_1(h)))
        for i_level in reversed(range(self.num_resolutions)):
            for blk in self.up[i_level].block:
                h = blk(h)
            if i_level != 0:
                h = self.up[i_level].upsample(h)
        return self.conv_out(_swish(self.norm_out(h)))


class OracleFirstStageDecoder(nn.Module):
    """``first_stage_model`` restricted to what decoding needs; state-dict keys = upstream's under ``first_stage_model.``"""

    def __init__(self, embed_dim=4, ddconfig=None):
        super().__init__()
        dd = dict(ch=128, out_ch=3, ch_mult=(1, 2, 4, 4), num_res_blocks=2, z_channels=4)
        dd.update(ddconfig or {})
        self.decoder = Decoder(**dd)
        self.post_quant_conv = nn.Conv2d(embed_dim, dd["z_channels"], 1)

    def decode(self, z):
        return self.decoder(self.post_quant_conv(z))


def decode_first_stage(first_stage: OracleFirstStageDecoder, z, scale_factor=0.18215):
    """upstream LatentDiffusion.decode_first_stage / diffmk/makeups.py:260-262 (scale_factor: yaml:47)"""
    return first_stage.decode(1.0 / scale_factor * z)


# ---- encode side: SURVEY.md §8(f) rank 2 (x_p entry) -------------------------------------------------------------------
class Downsample(nn.Module):
    """upstream Downsample(with_conv=True): zero-pad the bottom / right edge by one, then 3x3 stride-2 conv without padding"""

    def __init__(self, c):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, stride=2, padding=0)

    def forward(self, x):
        return self.conv(F.pad(x, (0, 1, 0, 1), mode="constant", value=0))


class Encoder(nn.Module):
    """upstream ldm.modules.diffusionmodules.model.Encoder for yaml:88-105 (double_z: 2 * z_channels output channels)"""

    def __init__(self, ch=128, ch_mult=(1, 2, 4, 4), num_res_blocks=2, in_channels=3, z_channels=4, double_z=True, **unused):
        super().__init__()
        self.num_resolutions, self.num_res_blocks = len(ch_mult), num_res_blocks
        self.conv_in = nn.Conv2d(in_channels, ch, 3, padding=1)
        in_mult = (1,) + tuple(ch_mult)
        self.down = nn.ModuleList()
        block_in = ch
        for i_level in range(self.num_resolutions):
            block_in, block_out = ch * in_mult[i_level], ch * ch_mult[i_level]
            down = nn.Module()
            down.block = nn.ModuleList()
            for _ in range(num_res_blocks):
                down.block.append(ResnetBlock(block_in, block_out))
                block_in = block_out
            if i_level != self.num_resolutions - 1:
                down.downsample = Downsample(block_in)
            self.down.append(down)
        self.mid = nn.Module()
        self.mid.block_1 = ResnetBlock(block_in, block_in)
        self.mid.attn_1 = AttnBlock(block_in)
        self.mid.block_2 = ResnetBlock(block_in, block_in)
        self.norm_out = _norm(block_in)
        self.conv_out = nn.Conv2d(block_in, 2 * z_channels if double_z else z_channels, 3, padding=1)

    def forward(self, x):
        h = self.conv_in(x)
        for i_level in range(self.num_resolutions):
            for blk in self.down[i_level].block:
                h = blk(h)
            if i_level != self.num_resolutions - 1:
                h = self.down[i_level].downsample(h)
        h = self.mid.block_2(self.mid.attn_1(self.mid.block_1(h)))
        return self.conv_out(_swish(self.norm_out(h)))


class OracleFirstStageEncoder(nn.Module):
    """``first_stage_model`` restricted to what encoding needs (``encoder.*`` + ``quant_conv``).  ``encode`` returns the
    moments (mean, logvar clamped to [-30, 20]) of upstream's DiagonalGaussianDistribution."""

    def __init__(self, embed_dim=4, ddconfig=None):
        super().__init__()
        dd = dict(ch=128, ch_mult=(1, 2, 4, 4), num_res_blocks=2, in_channels=3, z_channels=4, double_z=True)
        dd.update(ddconfig or {})
        self.encoder = Encoder(**dd)
        self.quant_conv = nn.Conv2d(2 * dd["z_channels"], 2 * embed_dim, 1)

    def encode(self, x):
        mean, logvar = torch.chunk(self.quant_conv(self.encoder(x)), 2, dim=1)
        return mean, torch.clamp(logvar, -30.0, 20.0)


def get_z(first_stage: OracleFirstStageEncoder, x, noise, scale_factor=0.18215):
    """diffmk/makeup_diffuse.py:37-40: z = scale_factor * posterior.sample(), sample = mean + exp(logvar / 2) * noise"""
    mean, logvar = first_stage.encode(x)
    return scale_factor * (mean + torch.exp(0.5 * logvar) * noise)
