"""TEST-ONLY stand-ins for makeupdiffuse_b200.ops: each function restates the C-ABI contract of include/mkd_b200.h in
plain torch on the CPU so that the HOST logic (buffer/slot plumbing, weight repacking, call order) can be tested
without a GPU.  Never imported by the product."""
import math

import torch
import torch.nn.functional as F

from makeupdiffuse_b200 import _lib as L


def device_ok(device=0):
    pass


def groupnorm_workspace_bytes(N, groups=32):
    return N * 64 * groups * 8


def ddim_update(x, eps, x_prev, *, sqrt_one_minus_at, sqrt_at, sqrt_a_prev, dir_coef, sigma_t=0.0, temperature=1.0,
                noise=None, pred_x0=None, cfg_scale=None, peer_ptrs=None):
    assert peer_ptrs is None, "peer stores need real GPUs"
    f = lambda v: torch.tensor(v, dtype=torch.float32)  # noqa: E731
    e = eps
    if cfg_scale is not None:
        eu, ec = eps.chunk(2)
        e = eu + f(cfg_scale) * (ec - eu)
    p0 = (x - f(sqrt_one_minus_at) * e) / f(sqrt_at)
    xp = f(sqrt_a_prev) * p0 + f(dir_coef) * e
    if noise is not None:
        xp = xp + f(sigma_t) * noise * f(temperature)
    x_prev.copy_(xp)
    if pred_x0 is not None:
        pred_x0.copy_(p0)


def nchw_to_nhwc(src, dst2d):
    N, C, H, W = src.shape
    dst2d.copy_(src.permute(0, 2, 3, 1).reshape(N * H * W, C))


def nhwc_to_nchw(src2d, dst):
    N, C, H, W = dst.shape
    dst.copy_(src2d.float().reshape(N, H, W, C).permute(0, 3, 1, 2))


def timestep_embedding(t, out, max_period=10000.0):
    B, dim = out.shape
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(half, dtype=torch.float32) / half)
    args = t[:, None].float() * freqs[None]
    out.copy_(torch.cat([torch.cos(args), torch.sin(args)], -1))


def silu(x, y):
    y.copy_(F.silu(x.float()))


def geglu(x2d, y2d):
    inner = y2d.shape[1]
    y2d.copy_(x2d[:, :inner].float() * F.gelu(x2d[:, inner:].float()))


def add(a2d, b2d, y2d):
    y2d.copy_(a2d.float() + b2d.float())


def _by_parts(fn, x2d, y2d, N, gamma, beta, eps, silu, workspace, groups):
    h, C = x2d.shape[0] // 2, x2d.shape[1]
    for g in (0, 1):
        fn(x2d[g * h:(g + 1) * h], y2d[g * h:(g + 1) * h], N // 2, gamma.reshape(2, C)[g], beta.reshape(2, C)[g], eps, silu, workspace, groups)


def groupnorm(x2d, y2d, N, gamma, beta, eps, silu, workspace, groups=32, wgroups=1):
    if wgroups == 2:
        return _by_parts(groupnorm, x2d, y2d, N, gamma, beta, eps, silu, workspace, groups)
    M, C = x2d.shape
    xr = x2d.float().reshape(N, M // N, C).permute(0, 2, 1)
    r = F.group_norm(xr, groups, gamma, beta, eps)
    if silu:
        r = F.silu(r)
    y2d.copy_(r.permute(0, 2, 1).reshape(M, C))


def groupnorm_apply(x2d, y2d, N, gamma, beta, eps, silu, stats, groups=32, wgroups=1):
    if wgroups == 2:
        h, C = x2d.shape[0] // 2, x2d.shape[1]
        for g in (0, 1):
            groupnorm_apply(x2d[g * h:(g + 1) * h], y2d[g * h:(g + 1) * h], N // 2, gamma.reshape(2, C)[g], beta.reshape(2, C)[g], eps, silu,
                            stats[g * (stats.shape[0] // 2):(g + 1) * (stats.shape[0] // 2)], groups)
        return
    """mkd_groupnorm_apply: statistics come from the producer's per-tile (sum, sumsq) partials, not from x"""
    M, C = x2d.shape
    tiles = (M // N) // 128
    assert stats.shape == (M // 128, C, 2)
    tot = stats.float().reshape(N, tiles, C, 2).sum(1)                      # [N, C, 2]
    cg = C // groups
    gsum = tot.reshape(N, groups, cg, 2).sum(2)                             # [N, groups, 2]
    cnt = cg * (M // N)
    mean = gsum[..., 0] / cnt
    var = (gsum[..., 1] / cnt - mean * mean).clamp_min(0)
    rstd = torch.rsqrt(var + eps)
    mean_c = mean.repeat_interleave(cg, 1)[:, None, :]
    rstd_c = rstd.repeat_interleave(cg, 1)[:, None, :]
    r = (x2d.float().reshape(N, M // N, C) - mean_c) * rstd_c * gamma + beta
    if silu:
        r = F.silu(r)
    y2d.copy_(r.reshape(M, C))


def softmax_rows(x2d, y2d, scale=1.0):
    y2d.copy_(torch.softmax(x2d.float() * scale, -1))


def layernorm(x2d, y2d, gamma, beta, eps=1e-5, wgroups=1):
    if wgroups == 2:
        h, C = x2d.shape[0] // 2, x2d.shape[1]
        for g in (0, 1):
            layernorm(x2d[g * h:(g + 1) * h], y2d[g * h:(g + 1) * h], gamma.reshape(2, C)[g], beta.reshape(2, C)[g], eps)
        return
    y2d.copy_(F.layer_norm(x2d.float(), (x2d.shape[1],), gamma, beta, eps))


def conv2d(x2d, w, y2d, *, N, H, W, R=1, S=1, stride=1, pad=0, upsample=False, bias=None, emb=None, residual=None,
           alpha=1.0, act=L.ACT_NONE, geglu_block=0, path=L.PATH_AUTO, workspace=None, y32=None, stats=None,
           pad_hi_extra=0, x2=None, wgroups=1, gn=None):
    if wgroups == 2:
        return conv2d_grouped(x2d, w, y2d, N=N, H=H, W=W, R=R, S=S, stride=stride, pad=pad, upsample=upsample, bias=bias, emb=emb,
                              residual=residual, alpha=alpha, act=act, geglu_block=geglu_block, path=path, workspace=workspace,
                              y32=y32, stats=stats, pad_hi_extra=pad_hi_extra, x2=x2, gn=gn)
    C = x2d.shape[1]
    K = w.shape[0]
    C2 = 0 if x2 is None else x2.shape[1]
    assert x2d.shape[0] == N * H * W and w.numel() == K * (R * S * C + C2)
    w2 = None
    if x2 is not None:  # fused second 1x1 term (mkd_conv_desc.x2): weight rows [R][S][C] + C2 columns
        assert stride == 1 and not upsample and act != L.ACT_GEGLU
        w, w2 = w.reshape(K, -1)[:, :R * S * C], w.reshape(K, -1)[:, R * S * C:]
    xr = x2d.float().reshape(N, H, W, C).permute(0, 3, 1, 2)
    if upsample:
        xr = F.interpolate(xr, scale_factor=2, mode="nearest")
    if pad_hi_extra:
        xr = F.pad(xr, (0, pad_hi_extra, 0, pad_hi_extra))
    acc = F.conv2d(xr, w.float().reshape(K, R, S, C).permute(0, 3, 1, 2), None if bias is None else bias.float(),
                   stride=stride, padding=pad)
    Mo = acc.shape[0] * acc.shape[2] * acc.shape[3]
    if w2 is not None:
        acc = acc + (x2.float() @ w2.float().t()).reshape(N, H, W, K).permute(0, 3, 1, 2)
    if emb is not None:
        acc = acc + emb.float()[:, :, None, None]
    acc = (acc * alpha).permute(0, 2, 3, 1).reshape(Mo, K)
    if act == L.ACT_GEGLU:
        gb = geglu_block
        Ko = K // 2
        assert emb is None and residual is None and alpha == 1.0 and Ko % gb == 0
        a = acc.reshape(Mo, Ko // gb, 2, gb)
        acc = (a[:, :, 0] * F.gelu(a[:, :, 1])).reshape(Mo, Ko)
    if residual is not None:
        acc = acc + residual.float()
    if act == L.ACT_SILU:
        acc = F.silu(acc)
    if stats is not None:  # per-128-row-tile column (sum, sumsq) of the stored values (mkd_conv_desc.stats)
        assert act == L.ACT_NONE and C >= 64 and Mo % 128 == 0 and stats.shape == (Mo // 128, K, 2)
        t = acc.reshape(Mo // 128, 128, K)
        stats[..., 0] = t.sum(1)
        stats[..., 1] = (t * t).sum(1)
    assert y2d is not None or y32 is not None
    for out in (y2d, y32):
        if out is not None:
            assert out.shape == acc.shape, (out.shape, acc.shape)
            out.copy_(acc)
    if gn is not None:  # GroupNorm tail (mkd_conv_desc.gn_y): GroupNorm(+SiLU) of the UNROUNDED result, per image
        assert act == L.ACT_NONE and stats is None and stride == 1 and not upsample
        groupnorm(acc, gn["y"], N, gn["gamma"], gn["beta"], gn["eps"], gn["silu"], None, gn.get("groups", 32))


def conv2d_grouped(x2d, w, y2d, *, N, **kw):
    """contract of mkd_conv_desc.wgroups = 2: part g of the rows with weight rows / biases [g K, (g + 1) K)"""
    K = w.shape[0] // 2
    out = y2d if y2d is not None else kw.get("y32")
    rx, ro = x2d.shape[0] // 2, out.shape[0] // 2

    def part(t, g, n):
        return None if t is None else t[g * n:(g + 1) * n]
    for g in (0, 1):
        k2 = dict(kw)
        k2["bias"] = part(kw.get("bias"), g, K)
        for name in ("residual", "y32"):
            k2[name] = part(kw.get(name), g, ro)
        k2["emb"] = part(kw.get("emb"), g, N // 2)
        k2["x2"] = part(kw.get("x2"), g, rx)
        if kw.get("stats") is not None:
            k2["stats"] = part(kw["stats"], g, kw["stats"].shape[0] // 2)
        if kw.get("gn") is not None:
            k2["gn"] = dict(kw["gn"], y=part(kw["gn"]["y"], g, ro), gamma=part(kw["gn"]["gamma"], g, K), beta=part(kw["gn"]["beta"], g, K))
        conv2d(part(x2d, g, rx), w[g * K:(g + 1) * K], part(y2d, g, ro), N=N // 2, **k2)


def conv2d_supported(x2d, w, y2d, **kw):
    """the contract of the x2 term: stride 1, C2 % 64 == 0, K % 160 == 0 (the CTA-pair kernel's tiles); the GroupNorm tail rides
    only in split-K launches, which the reduced test networks never are (tests that exercise its plumbing force the answer)"""
    if kw.get("gn") is not None:
        return False
    x2 = kw.get("x2")
    return x2 is None or (x2.shape[1] % 64 == 0 and w.shape[0] % 160 == 0 and x2d.shape[1] % 64 == 0)


def attention(q2d, k2d, v2d, o2d, *, B, heads, Nq, Nkv, d, scale):
    sp = lambda t, n: t.float().reshape(B, n, heads, d).permute(0, 2, 1, 3)  # noqa: E731
    o = F.scaled_dot_product_attention(sp(q2d, Nq), sp(k2d, Nkv), sp(v2d, Nkv), scale=scale)
    o2d.copy_(o.permute(0, 2, 1, 3).reshape(B * Nq, heads * d))


def attention_causal(q2d, k2d, v2d, o2d, *, B, heads, N, d, scale):
    sp = lambda t: t.float().reshape(B, N, heads, d).permute(0, 2, 1, 3)  # noqa: E731
    o = F.scaled_dot_product_attention(sp(q2d), sp(k2d), sp(v2d), scale=scale, is_causal=True)
    o2d.copy_(o.permute(0, 2, 1, 3).reshape(B * N, heads * d))


def embed_tokens(ids, tok_emb, pos_emb, out2d):
    B, T = ids.shape
    out2d.copy_((tok_emb[ids] + pos_emb[:T]).reshape(B * T, -1))


def image_grid_u8(images, nrow, padding=2, clamp=True, rescale=True):
    """the reference's sequence itself (diffusion_makeup.py:340-355)"""
    import numpy as np
    import torchvision
    x = images.detach().cpu()
    if clamp:
        x = torch.clamp(x, -1.0, 1.0)
    grid = torchvision.utils.make_grid(x, nrow=nrow, padding=padding)
    if rescale:
        grid = (grid + 1.0) / 2.0
    grid = grid.transpose(0, 1).transpose(1, 2).squeeze(-1).numpy()
    return torch.from_numpy((grid * 255).astype(np.uint8))


ALL = ["image_grid_u8", "attention_causal", "embed_tokens", "device_ok", "groupnorm_workspace_bytes", "ddim_update", "nchw_to_nhwc", "nhwc_to_nchw", "timestep_embedding",
       "silu", "geglu", "add", "groupnorm", "groupnorm_apply", "layernorm", "softmax_rows", "conv2d", "conv2d_supported", "conv2d_grouped", "attention"]
