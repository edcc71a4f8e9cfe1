"""Oracle restatement of the two networks ``apply_model`` drives: ControlNet and ControlledUnetModel.

Built from the hyper-parameters of ``diffmodels/base_diffusion_makeup.yaml:52-84``; the dataflow between the
two nets follows ``diffmk/makeup_diffuse.py:157-168``.  The module trees / op order are those of upstream
lllyasviel/ControlNet (``cldm.cldm``, ``ldm.modules.diffusionmodules.openaimodel``, ``ldm.modules.attention``;
NOT vendored in the reference) restated here in plain fp32 PyTorch.  Sub-module and parameter NAMES are kept
equal to upstream's (``input_blocks.{i}.{j}``, ``zero_convs.{i}.0``, ``transformer_blocks.0.attn1.to_q`` ...)
so that an SD-1.5 / ControlNet checkpoint would load; the parameter-count and key-name pins of SURVEY.md
§8(c) are asserted in tests/test_oracle_kat.py.

Test infrastructure (see oracle/__init__.py).
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F


def timestep_embedding(t: torch.Tensor, dim: int, max_period: float = 10000.0) -> torch.Tensor:
    """[cos(t*f_k), sin(t*f_k)], f_k = exp(-ln(max_period) * k / (dim/2)), k = 0..dim/2-1  (SURVEY A8c)."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(half, dtype=torch.float32, device=t.device) / half)
    args = t[:, None].float() * freqs[None]
    emb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
    if dim % 2:
        emb = torch.cat([emb, torch.zeros_like(emb[:, :1])], dim=-1)
    return emb


class GroupNorm32(nn.GroupNorm):
    """GroupNorm whose statistics are always computed in fp32 (upstream GroupNorm32)."""

    def forward(self, x):
        return super().forward(x.float()).type(x.dtype)


class EmbedSequential(nn.Sequential):
    """Sequential that hands (emb) to ResBlocks and (context) to SpatialTransformers (upstream
    TimestepEmbedSequential)."""

    def forward(self, x, emb=None, context=None):
        for layer in self:
            if isinstance(layer, ResBlock):
                x = layer(x, emb)
            elif isinstance(layer, SpatialTransformer):
                x = layer(x, context)
            else:
                x = layer(x)
        return x


class Downsample(nn.Module):
    def __init__(self, ch):
        super().__init__()
        self.op = nn.Conv2d(ch, ch, 3, stride=2, padding=1)

    def forward(self, x):
        return self.op(x)


class Upsample(nn.Module):
    def __init__(self, ch):
        super().__init__()
        self.conv = nn.Conv2d(ch, ch, 3, padding=1)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=2, mode="nearest"))


class ResBlock(nn.Module):
    """h = conv(SiLU(GN(x))); h += Linear(SiLU(emb)); h = conv(SiLU(GN(h))); out = skip(x) + h   (SURVEY A8a)."""

    def __init__(self, cin, emb_ch, cout):
        super().__init__()
        self.in_layers = nn.Sequential(GroupNorm32(32, cin), nn.SiLU(), nn.Conv2d(cin, cout, 3, padding=1))
        self.emb_layers = nn.Sequential(nn.SiLU(), nn.Linear(emb_ch, cout))
        self.out_layers = nn.Sequential(GroupNorm32(32, cout), nn.SiLU(), nn.Dropout(0.0),
                                        nn.Conv2d(cout, cout, 3, padding=1))
        self.skip_connection = nn.Identity() if cin == cout else nn.Conv2d(cin, cout, 1)

    def forward(self, x, emb):
        h = self.in_layers(x)
        h = h + self.emb_layers(emb)[:, :, None, None]
        h = self.out_layers(h)
        return self.skip_connection(x) + h


class CrossAttention(nn.Module):
    def __init__(self, query_dim, context_dim, heads, dim_head):
        super().__init__()
        inner = heads * dim_head
        context_dim = query_dim if context_dim is None else context_dim
        self.heads, self.scale = heads, dim_head ** -0.5
        self.to_q = nn.Linear(query_dim, inner, bias=False)
        self.to_k = nn.Linear(context_dim, inner, bias=False)
        self.to_v = nn.Linear(context_dim, inner, bias=False)
        self.to_out = nn.Sequential(nn.Linear(inner, query_dim), nn.Dropout(0.0))

    def forward(self, x, context=None):
        context = x if context is None else context
        b, n, _ = x.shape
        h = self.heads

        def split(t):  # b n (h d) -> b h n d
            return t.reshape(b, t.shape[1], h, -1).permute(0, 2, 1, 3)

        q, k, v = split(self.to_q(x)), split(self.to_k(context)), split(self.to_v(context))
        sim = torch.matmul(q.float(), k.float().transpose(-1, -2)) * self.scale
        attn = sim.softmax(dim=-1).to(v.dtype)
        out = torch.matmul(attn, v).permute(0, 2, 1, 3).reshape(b, n, -1)
        return self.to_out(out)


class GEGLU(nn.Module):
    def __init__(self, din, dout):
        super().__init__()
        self.proj = nn.Linear(din, dout * 2)

    def forward(self, x):
        a, gate = self.proj(x).chunk(2, dim=-1)
        return a * F.gelu(gate)


class FeedForward(nn.Module):
    def __init__(self, dim, mult=4):
        super().__init__()
        self.net = nn.Sequential(GEGLU(dim, dim * mult), nn.Dropout(0.0), nn.Linear(dim * mult, dim))

    def forward(self, x):
        return self.net(x)


class BasicTransformerBlock(nn.Module):
    def __init__(self, dim, heads, dim_head, context_dim):
        super().__init__()
        self.attn1 = CrossAttention(dim, None, heads, dim_head)
        self.ff = FeedForward(dim)
        self.attn2 = CrossAttention(dim, context_dim, heads, dim_head)
        self.norm1, self.norm2, self.norm3 = nn.LayerNorm(dim), nn.LayerNorm(dim), nn.LayerNorm(dim)

    def forward(self, x, context):
        x = self.attn1(self.norm1(x)) + x
        x = self.attn2(self.norm2(x), context) + x
        return self.ff(self.norm3(x)) + x


class SpatialTransformer(nn.Module):
    """GN(eps 1e-6) -> 1x1 -> tokens -> [self-attn, cross-attn, GEGLU-FF] -> 1x1 -> + input  (SURVEY A8b)."""

    def __init__(self, ch, heads, dim_head, depth, context_dim):
        super().__init__()
        inner = heads * dim_head
        self.norm = nn.GroupNorm(32, ch, eps=1e-6, affine=True)
        self.proj_in = nn.Conv2d(ch, inner, 1)
        self.transformer_blocks = nn.ModuleList(
            [BasicTransformerBlock(inner, heads, dim_head, context_dim) for _ in range(depth)])
        self.proj_out = nn.Conv2d(inner, ch, 1)

    def forward(self, x, context):
        b, c, hh, ww = x.shape
        x_in = x
        x = self.proj_in(self.norm(x))
        x = x.reshape(b, -1, hh * ww).transpose(1, 2)
        for blk in self.transformer_blocks:
            x = blk(x, context)
        x = x.transpose(1, 2).reshape(b, -1, hh, ww)
        return self.proj_out(x) + x_in


class _Encoder(nn.Module):
    """time_embed + input_blocks + middle_block shared by ControlNet and the UNet (yaml:54-67 / 71-84)."""

    def __init__(self, in_channels, model_channels, attention_resolutions, num_res_blocks, channel_mult,
                 num_heads, transformer_depth, context_dim):
        super().__init__()
        mc = model_channels
        self.model_channels = mc
        ted = mc * 4
        self.time_embed = nn.Sequential(nn.Linear(mc, ted), nn.SiLU(), nn.Linear(ted, ted))
        self.input_blocks = nn.ModuleList([EmbedSequential(nn.Conv2d(in_channels, mc, 3, padding=1))])
        self.block_chans = [mc]
        ch, ds = mc, 1

        def st(c):
            return SpatialTransformer(c, num_heads, c // num_heads, transformer_depth, context_dim)

        self._st = st
        for level, mult in enumerate(channel_mult):
            for _ in range(num_res_blocks):
                layers = [ResBlock(ch, ted, mult * mc)]
                ch = mult * mc
                if ds in attention_resolutions:
                    layers.append(st(ch))
                self.input_blocks.append(EmbedSequential(*layers))
                self.block_chans.append(ch)
            if level != len(channel_mult) - 1:
                self.input_blocks.append(EmbedSequential(Downsample(ch)))
                self.block_chans.append(ch)
                ds *= 2
        self.middle_block = EmbedSequential(ResBlock(ch, ted, ch), st(ch), ResBlock(ch, ted, ch))
        self._ch, self._ds, self._ted = ch, ds, ted


class ControlNet(_Encoder):
    """cldm.cldm.ControlNet as configured by yaml:52-67.  forward -> list of 13 residuals."""

    def __init__(self, image_size=32, in_channels=4, hint_channels=6, model_channels=320,
                 attention_resolutions=(4, 2, 1), num_res_blocks=2, channel_mult=(1, 2, 4, 4), num_heads=8,
                 use_spatial_transformer=True, transformer_depth=1, context_dim=768, use_checkpoint=False,
                 legacy=False, hint_widths=(16, 16, 32, 32, 96, 96, 256)):
        super().__init__(in_channels, model_channels, attention_resolutions, num_res_blocks, channel_mult,
                         num_heads, transformer_depth, context_dim)
        w = hint_widths
        hb, cin = [], hint_channels
        strides = (1, 1, 2, 1, 2, 1, 2)
        for cout, s in zip(w, strides):
            hb += [nn.Conv2d(cin, cout, 3, stride=s, padding=1), nn.SiLU()]
            cin = cout
        hb.append(nn.Conv2d(cin, model_channels, 3, padding=1))
        self.input_hint_block = EmbedSequential(*hb)
        self.zero_convs = nn.ModuleList([EmbedSequential(nn.Conv2d(c, c, 1)) for c in self.block_chans])
        self.middle_block_out = EmbedSequential(nn.Conv2d(self._ch, self._ch, 1))

    def forward(self, x, hint, timesteps, context, **kwargs):
        emb = self.time_embed(timestep_embedding(timesteps, self.model_channels))
        guided_hint = self.input_hint_block(hint, emb, context)
        outs, h = [], x
        for blk, zc in zip(self.input_blocks, self.zero_convs):
            h = blk(h, emb, context)
            if guided_hint is not None:
                h = h + guided_hint
                guided_hint = None
            outs.append(zc(h, emb, context))
        h = self.middle_block(h, emb, context)
        outs.append(self.middle_block_out(h, emb, context))
        return outs


class ControlledUnetModel(_Encoder):
    """cldm.cldm.ControlledUnetModel as configured by yaml:69-84."""

    def __init__(self, image_size=32, in_channels=4, out_channels=4, model_channels=320,
                 attention_resolutions=(4, 2, 1), num_res_blocks=2, channel_mult=(1, 2, 4, 4), num_heads=8,
                 use_spatial_transformer=True, transformer_depth=1, context_dim=768, use_checkpoint=False,
                 legacy=False):
        super().__init__(in_channels, model_channels, attention_resolutions, num_res_blocks, channel_mult,
                         num_heads, transformer_depth, context_dim)
        mc, ch, ds, ted = model_channels, self._ch, self._ds, self._ted
        chans = list(self.block_chans)
        self.output_blocks = nn.ModuleList()
        for level, mult in list(enumerate(channel_mult))[::-1]:
            for i in range(num_res_blocks + 1):
                ich = chans.pop()
                layers = [ResBlock(ch + ich, ted, mc * mult)]
                ch = mc * mult
                if ds in attention_resolutions:
                    layers.append(self._st(ch))
                if level and i == num_res_blocks:
                    layers.append(Upsample(ch))
                    ds //= 2
                self.output_blocks.append(EmbedSequential(*layers))
        self.out = nn.Sequential(GroupNorm32(32, ch), nn.SiLU(), nn.Conv2d(mc, out_channels, 3, padding=1))

    def forward(self, x, timesteps=None, context=None, control=None, only_mid_control=False, **kwargs):
        emb = self.time_embed(timestep_embedding(timesteps, self.model_channels))
        hs, h = [], x
        for blk in self.input_blocks:
            h = blk(h, emb, context)
            hs.append(h)
        h = self.middle_block(h, emb, context)
        if control is not None:
            control = list(control)
            h = h + control.pop()
        for blk in self.output_blocks:
            if only_mid_control or control is None:
                h = torch.cat([h, hs.pop()], dim=1)
            else:
                h = torch.cat([h, hs.pop() + control.pop()], dim=1)
            h = blk(h, emb, context)
        return self.out(h)
