"""Per-kernel parity on a real B200: every C-ABI entry point against a plain PyTorch fp32 statement of the same op
(the reference executes these as PyTorch library calls).  All calls go through the C-ABI (ops.py -> ctypes)."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from makeupdiffuse_b200 import _lib as L  # noqa: E402
from makeupdiffuse_b200 import ops  # noqa: E402

DEV = "cuda"
BF, F32 = torch.bfloat16, torch.float32


def rel(a, b):
    a, b = a.float(), b.float()
    return float((a - b).norm() / b.norm().clamp_min(1e-20))


def tol(dt):
    return 6e-3 if dt == BF else 2e-5


def rnd(*shape, dt=F32, seed=0, scale=1.0):
    g = torch.Generator(device=DEV).manual_seed(seed)
    return (torch.randn(*shape, device=DEV, generator=g) * scale).to(dt)


def test_device_is_b200():
    ops.device_ok(0)
    assert torch.cuda.get_device_capability(0) == (10, 0)


# ---- DDIM update: bit-exact against the reference's op-by-op fp32 arithmetic (cddim.py:39-40,56-78) --------------
@pytest.mark.parametrize("cfg", [False, True])
@pytest.mark.parametrize("sigma", [0.0, 0.37])
@pytest.mark.parametrize("n", [1, 4096 * 16 + 3])
def test_ddim_update_bit_exact(cfg, sigma, n):
    x, e_c, e_u, nz = rnd(n, seed=1), rnd(n, seed=2), rnd(n, seed=3), rnd(n, seed=4)
    a_t, a_prev, scale, temp = 0.2314, 0.3127, 9.0, 0.9
    full = lambda v: torch.full((1,), v, device=DEV)  # noqa: E731
    ta, tp, ts, t1 = full(a_t), full(a_prev), full(sigma), full(math.sqrt(1 - a_t))
    e = e_u + scale * (e_c - e_u) if cfg else e_c
    pred = (x - t1 * e) / ta.sqrt()
    ref = tp.sqrt() * pred + (1.0 - tp - ts ** 2).sqrt() * e + ts * nz * temp
    out, p0 = torch.empty_like(x), torch.empty_like(x)
    ops.ddim_update(x, torch.cat([e_u, e_c]) if cfg else e_c, out, sqrt_one_minus_at=float(t1), sqrt_at=float(ta.sqrt()),
                    sqrt_a_prev=float(tp.sqrt()), dir_coef=float((1.0 - tp - ts ** 2).sqrt()), sigma_t=sigma,
                    temperature=temp, noise=nz if sigma else None, pred_x0=p0, cfg_scale=scale if cfg else None)
    if not sigma:
        ref = tp.sqrt() * pred + (1.0 - tp - ts ** 2).sqrt() * e
    assert torch.equal(p0, pred)
    assert torch.equal(out, ref)


@pytest.mark.parametrize("dt", [BF, F32])
def test_layout_roundtrip(dt):
    x = rnd(3, 6, 20, 12)
    buf = torch.zeros(3 * 20 * 12, 8, device=DEV, dtype=dt)
    ops.nchw_to_nhwc(x, buf[:, :6])
    ref = x.permute(0, 2, 3, 1).reshape(-1, 6).to(dt)
    assert torch.equal(buf[:, :6], ref) and float(buf[:, 6:].abs().max()) == 0
    back = torch.empty_like(x)
    ops.nhwc_to_nchw(buf[:, :6], back)
    assert torch.equal(back, x.to(dt).float())


@pytest.mark.parametrize("dt", [BF, F32])
def test_timestep_embedding(dt):
    t = torch.tensor([0, 1, 21, 501, 981, 999], device=DEV)
    out = torch.empty(6, 320, device=DEV, dtype=dt)
    ops.timestep_embedding(t, out)
    half = 160
    freqs = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=F32, device=DEV) / half)
    args = t[:, None].float() * freqs[None]
    ref = torch.cat([torch.cos(args), torch.sin(args)], -1)
    assert float((out.float() - ref).abs().max()) < (1e-2 if dt == BF else 2e-4)


@pytest.mark.parametrize("dt", [BF, F32])
def test_silu_geglu_add(dt):
    x = rnd(37, 256, dt=dt)
    y = torch.empty_like(x)
    ops.silu(x, y)
    assert rel(y, F.silu(x.float())) < tol(dt)
    wide = rnd(50, 2 * 128 + 16, dt=dt, seed=3)
    out = torch.zeros(50, 136, device=DEV, dtype=dt)
    ops.geglu(wide[:, :256], out[:, :128])
    ref = wide[:, :128].float() * F.gelu(wide[:, 128:256].float())
    assert rel(out[:, :128], ref) < tol(dt) and float(out[:, 128:].abs().max()) == 0
    a, b = rnd(50, 64, dt=dt, seed=5), rnd(50, 64, dt=dt, seed=6)
    ops.add(a, b, a)
    # a was overwritten in place: recompute the reference from fresh copies
    a0 = rnd(50, 64, dt=dt, seed=5)
    assert rel(a, a0.float() + b.float()) < tol(dt)


@pytest.mark.parametrize("dt", [BF, F32, (F32, BF)])
@pytest.mark.parametrize("N,HW,C,ld,silu,eps", [(2, 1024, 320, 320, True, 1e-5), (3, 256, 960, 960, True, 1e-5),
                                                 (2, 64, 1920, 1920, False, 1e-6), (16, 16, 2560, 2560, True, 1e-5),
                                                 (2, 1024, 320, 640, True, 1e-5), (1, 64, 64, 64, True, 1e-5),
                                                 (5, 4096, 320, 320, False, 1e-6), (2, 100, 128, 136, True, 1e-5),
                                                 (16, 1024, 320, 320, True, 1e-5), (16, 256, 1920, 1920, True, 1e-5),
                                                 (16, 64, 1280, 2560, False, 1e-6), (12, 4096, 320, 320, True, 1e-5),
                                                 (16, 64, 2560, 2560, True, 1e-5), (3, 64, 256, 512, True, 1e-5),
                                                 (2, 16, 1280, 1280, False, 1e-6)])
def test_groupnorm(dt, N, HW, C, ld, silu, eps):
    dt, dto = dt if isinstance(dt, tuple) else (dt, dt)  # (fp32 trunk in, bf16 operand out) is the bf16 path's common case
    buf = rnd(N * HW, ld, dt=dt, seed=1) * 1.7 + 0.4
    x = buf[:, ld - C:]  # channel slice of a wider (concat) buffer
    gamma, beta = 1 + 0.2 * rnd(C, seed=2), 0.1 * rnd(C, seed=3)
    y = torch.empty(N * HW, C, device=DEV, dtype=dto)
    ws = torch.empty(ops.groupnorm_workspace_bytes(N) // 4, device=DEV)
    ops.groupnorm(x, y, N, gamma, beta, eps, silu, ws)
    xr = x.float().reshape(N, HW, C).permute(0, 2, 1)
    ref = F.group_norm(xr, 32, gamma, beta, eps)
    ref = (F.silu(ref) if silu else ref).permute(0, 2, 1).reshape(N * HW, C)
    assert rel(y, ref) < tol(dto)


@pytest.mark.parametrize("dt", [(F32, BF), F32, BF])
@pytest.mark.parametrize("M,C,scale", [(1024, 1024, 512 ** -0.5), (77, 256, 0.125), (5, 2048, 1.0), (300, 64, 3.0)])
def test_softmax_rows(dt, M, C, scale):
    dt, dto = dt if isinstance(dt, tuple) else (dt, dt)
    x = rnd(M, C, dt=dt, seed=1) * 4
    y = torch.empty(M, C, device=DEV, dtype=dto)
    ops.softmax_rows(x, y, scale)
    ref = torch.softmax(x.float() * scale, -1)
    assert rel(y, ref) < tol(dto)
    assert float((y.float().sum(-1) - 1).abs().max()) < (2e-2 if dto == BF else 1e-5)


@pytest.mark.parametrize("dt", [BF, F32, (F32, BF)])
@pytest.mark.parametrize("M,C", [(1024, 320), (77, 640), (300, 1280), (5, 64)])
def test_layernorm(dt, M, C):
    dt, dto = dt if isinstance(dt, tuple) else (dt, dt)
    x = rnd(M, C, dt=dt, seed=1) * 2 + 0.3
    g, b = 1 + 0.2 * rnd(C, seed=2), 0.1 * rnd(C, seed=3)
    y = torch.empty(M, C, device=DEV, dtype=dto)
    ops.layernorm(x, y, g, b)
    assert rel(y, F.layer_norm(x.float(), (C,), g, b)) < tol(dto)


# ---- convolution / GEMM family -----------------------------------------------------------------------------------
def conv_case(dt, path, N, H, W, C, K, R=1, stride=1, upsample=False, bias=True, emb=False, residual=False,
              inplace=False, alpha=1.0, act=L.ACT_NONE, ldx_extra=0, ldy_extra=0, ldy_pad=0, workspace=False, seed=0,
              y32="", res32=False, gb=None, C2=0):
    """y32: "" (activation-dtype output only), "both" (+ fp32 copy) or "only" (fp32 copy only); res32: fp32 residual;
    C2: channels of a fused second 1x1 term over another input (mkd_conv_desc.x2), 0 = none"""
    M = N * H * W
    xb = rnd(M, C + ldx_extra, dt=dt, seed=seed + 1)
    x = xb[:, ldx_extra:]
    w_oihw = rnd(K, C, R, R, seed=seed + 2, scale=1.0 / math.sqrt(C * R * R)).to(dt)
    w = w_oihw.permute(0, 2, 3, 1).contiguous()
    x2 = w2 = None
    if C2:  # the second input is a channel slice of a wider buffer, like a decoder concat
        x2 = rnd(M, C2 + 64, dt=dt, seed=seed + 7)[:, 64:]
        w2 = rnd(K, C2, seed=seed + 8, scale=1.0 / math.sqrt(C2)).to(dt)
        w = torch.cat([w.reshape(K, -1), w2], 1).contiguous()
    b = 0.5 * rnd(K, seed=seed + 3) if bias else None
    e = rnd(N, K + 8, dt=dt, seed=seed + 4)[:, 8:] if emb else None
    Hi, Wi = (2 * H, 2 * W) if upsample else (H, W)
    P, Q = (Hi + 2 * (R // 2) - R) // stride + 1, (Wi + 2 * (R // 2) - R) // stride + 1
    Mo = N * P * Q
    Ko = K // 2 if act == L.ACT_GEGLU else K
    yb = rnd(Mo, Ko + ldy_extra + ldy_pad, dt=dt, seed=seed + 5)
    y = yb[:, ldy_extra:ldy_extra + Ko]
    res = y if inplace else (rnd(Mo, Ko, dt=F32 if res32 else dt, seed=seed + 6) if residual else None)
    out32 = torch.full((Mo, Ko + 8), 3.0, device=DEV)[:, :Ko] if y32 else None
    if y32 == "only":
        assert not inplace
        y = None
    res_ref = None if res is None else res.float().clone()
    if gb is None:
        gb = 80 if act == L.ACT_GEGLU and Ko % 80 == 0 else 16
    ws = torch.empty(64 << 20, dtype=torch.uint8, device=DEV) if workspace else None
    kw = dict(N=N, H=H, W=W, R=R, S=R, stride=stride, pad=R // 2, upsample=upsample, bias=b, emb=e, residual=res,
              alpha=alpha, act=act, geglu_block=gb, path=path, workspace=ws, y32=out32, x2=x2)
    want = min(path, L.PATH_TCGEN05)  # the two forced tensor-core kernels both report PATH_TCGEN05
    assert ops.conv2d_path(x, w, y, **kw) == (want if path else ops.conv2d_path(x, w, y, **kw))
    ops.conv2d(x, w, y, **kw)
    # fp32 reference on the same (rounded) inputs
    xr = x.float().reshape(N, H, W, C).permute(0, 3, 1, 2)
    if upsample:
        xr = F.interpolate(xr, scale_factor=2, mode="nearest")
    wr = w_oihw.float()
    if act == L.ACT_GEGLU:  # undo the [gb value | gb gate] row blocking
        wv = wr.reshape(Ko // gb, 2, gb, C, R, R)
        wr = torch.cat([wv[:, 0].reshape(Ko, C, R, R), wv[:, 1].reshape(Ko, C, R, R)], 0)
        bv = b.reshape(Ko // gb, 2, gb)
        br = torch.cat([bv[:, 0].reshape(Ko), bv[:, 1].reshape(Ko)], 0)
    else:
        br = b
    acc = F.conv2d(xr.double(), wr.double(), None if br is None else br.double(), stride=stride, padding=R // 2)
    if C2:
        acc = acc + (x2.double() @ w2.double().t()).reshape(N, H, W, K).permute(0, 3, 1, 2)
    if emb:
        acc = acc + e.double()[:, :, None, None]
    acc = acc * alpha
    acc = acc.permute(0, 2, 3, 1).reshape(Mo, K)
    if res_ref is not None:
        acc = acc + res_ref.double()
    if act == L.ACT_SILU:
        acc = F.silu(acc)
    if act == L.ACT_GEGLU:
        acc = acc[:, :Ko] * F.gelu(acc[:, Ko:])
    if out32 is not None:
        e32 = rel(out32, acc.float())
        assert e32 < (2e-5 if dt == F32 else 3e-3 if res32 else 6e-3), f"fp32 copy rel err {e32}"  # operands are still bf16
        if y is None:
            return e32
        assert torch.equal(y, out32.to(dt))  # the bf16 output is the rounding of the fp32 copy
    err = rel(y, acc.float())
    assert err < tol(dt), f"rel err {err}"
    if ldy_extra:  # the kernel must not touch the columns outside its slice
        assert torch.equal(yb[:, :ldy_extra], rnd(Mo, Ko + ldy_extra + ldy_pad, dt=dt, seed=seed + 5)[:, :ldy_extra])
    if ldy_pad:
        assert torch.equal(yb[:, ldy_extra + Ko:], rnd(Mo, Ko + ldy_extra + ldy_pad, dt=dt, seed=seed + 5)[:, ldy_extra + Ko:])
    return err


GENERIC_CASES = [
    dict(N=2, H=8, W=8, C=4, K=64, R=3),                                   # conv_in
    dict(N=2, H=8, W=8, C=64, K=4, R=3),                                   # out conv
    dict(N=1, H=16, W=16, C=6, K=16, R=3, act=L.ACT_SILU),                 # hint conv 0
    dict(N=2, H=16, W=16, C=32, K=96, R=3, stride=2, act=L.ACT_SILU),      # hint stride-2
    dict(N=2, H=8, W=8, C=64, K=64, R=3, stride=2),                        # Downsample
    dict(N=2, H=4, W=4, C=64, K=64, R=3, upsample=True),                   # Upsample
    dict(N=3, H=8, W=8, C=64, K=128, R=3, emb=True),                       # ResBlock conv1 (+emb)
    dict(N=2, H=8, W=8, C=128, K=128, R=3, residual=True, ldx_extra=64, ldy_extra=64),  # ResBlock conv2 (+skip), slices
    dict(N=2, H=8, W=8, C=64, K=64, R=1, inplace=True, alpha=0.7),         # zero-conv injection
    dict(N=1, H=1, W=77, C=64, K=128, R=1, bias=False),                    # linear, no bias
    dict(N=1, H=1, W=100, C=64, K=512, R=1, act=L.ACT_GEGLU),              # GEGLU
    dict(N=1, H=1, W=3, C=1280, K=320, R=1, act=L.ACT_SILU),               # tiny-M linear (time embed)
    dict(N=2, H=8, W=8, C=64, K=64, R=3, stride=2, y32="both"),            # Downsample, bf16 + fp32 outputs
    dict(N=2, H=8, W=8, C=64, K=128, R=3, emb=True, y32="only"),           # ResBlock h kept in fp32
    dict(N=2, H=8, W=8, C=128, K=128, R=3, residual=True, res32=True, y32="both"),  # fp32 trunk residual
]


@pytest.mark.parametrize("dt", [BF, F32])
@pytest.mark.parametrize("case", GENERIC_CASES)
def test_conv_generic(dt, case):
    conv_case(dt, L.PATH_GENERIC, **case)


TC_CASES = [
    dict(N=1, H=1, W=128, C=64, K=32, R=1, bias=False),                    # one MMA group
    dict(N=1, H=1, W=128, C=128, K=160, R=1),                              # 2 k-blocks, BN=160
    dict(N=1, H=1, W=1024, C=320, K=320, R=1),                             # attention projection
    dict(N=1, H=1, W=1000, C=320, K=960, R=1, bias=False),                 # ragged M (qkv)
    dict(N=1, H=1, W=77 * 3, C=768, K=640, R=1, bias=False),               # cross k/v projection
    dict(N=1, H=1, W=16, C=1280, K=1280, R=1, act=L.ACT_SILU),             # time-embed linear, tiny M
    dict(N=1, H=1, W=300, C=320, K=2560, R=1, act=L.ACT_GEGLU),            # fused GEGLU
    dict(N=1, H=1, W=256, C=1280, K=320, R=1, residual=True),              # FF out + residual
    dict(N=1, H=1, W=200, C=640, K=640, R=1, inplace=True, alpha=0.5),     # in-place residual (injection)
    dict(N=2, H=32, W=32, C=320, K=320, R=3, emb=True),                    # ResBlock conv1 @32x32
    dict(N=2, H=32, W=32, C=320, K=320, R=3, residual=True),               # ResBlock conv2 @32x32
    dict(N=3, H=16, W=16, C=640, K=640, R=3, emb=True),                    # @16x16
    dict(N=3, H=8, W=8, C=1280, K=1280, R=3, residual=True),               # @8x8, box spans 2 images, ragged N
    dict(N=5, H=4, W=4, C=1280, K=1280, R=3, emb=True),                    # @4x4, box spans 8 images
    dict(N=1, H=64, W=64, C=320, K=320, R=3, emb=True),                    # 512^2 level: box = 2 rows of 64
    dict(N=2, H=16, W=16, C=960, K=640, R=3, ldx_extra=320, ldy_extra=640, residual=True),  # concat slices in/out
    dict(N=2, H=32, W=32, C=320, K=4, R=3, ldy_pad=4),                     # `out` conv: 4 channels in an 8-wide buffer
    dict(N=16, H=4, W=4, C=2560, K=1280, R=3, emb=True, workspace=True),   # split-K (deep level)
    dict(N=4, H=8, W=8, C=1280, K=1280, R=1, workspace=True),              # split-K plain
    dict(N=1, H=1, W=64, C=1280, K=10240, R=1, act=L.ACT_GEGLU, workspace=True),  # split-K + GEGLU
    dict(N=2, H=32, W=32, C=320, K=320, R=1, residual=True),               # proj_out 1x1 conv
    dict(N=2, H=32, W=32, C=320, K=320, R=3, stride=2, workspace=True, y32="both"),   # Downsample via im2col + GEMM
    dict(N=3, H=8, W=8, C=1280, K=1280, R=3, stride=2, workspace=True),               # deep Downsample (+ split-K)
    dict(N=2, H=8, W=8, C=1280, K=1280, R=3, upsample=True, workspace=True),          # Upsample 8 -> 16
    dict(N=2, H=16, W=16, C=640, K=640, R=3, upsample=True, workspace=True, ldx_extra=64, ldy_extra=640),  # 16 -> 32 into a slot
    dict(N=2, H=32, W=32, C=320, K=320, R=3, emb=True, y32="only"),        # ResBlock h -> fp32 only
    dict(N=2, H=16, W=16, C=640, K=640, R=3, residual=True, res32=True, y32="both", ldy_extra=640),  # trunk: slot + fp32
    dict(N=1, H=1, W=300, C=1280, K=320, R=1, residual=True, res32=True),  # ff2: fp32 stream in, bf16 operand out
    dict(N=8, H=4, W=4, C=1280, K=1280, R=3, residual=True, res32=True, y32="both", workspace=True),  # split-K + fp32
    dict(N=1, H=256, W=256, C=64, K=64, R=3, act=L.ACT_SILU),              # hint block (channel-padded) @256^2: box = half a row
    dict(N=2, H=256, W=256, C=64, K=64, R=3, stride=2, act=L.ACT_SILU, workspace=True),  # hint stride-2 @256^2
    dict(N=2, H=64, W=64, C=64, K=128, R=3, stride=2, act=L.ACT_SILU, workspace=True),   # hint 32(+pad) -> 96(+pad)
    dict(N=2, H=32, W=32, C=256, K=320, R=3),                              # last hint conv
    dict(N=4, H=64, W=64, C=128, K=256, R=3, residual=True, res32=True, y32="both"),  # VAE widths: N tile 128
    dict(N=8, H=64, W=64, C=64, K=512, R=1),                               # N tile 256
    dict(N=8, H=64, W=64, C=128, K=256, R=3, emb=True, y32="only"),        # N tile 256, 3x3
    dict(N=1, H=1, W=1024, C=512, K=1024, R=1, bias=False, y32="only"),    # VAE attention S = Q K^T (weights = keys)
]


@pytest.mark.parametrize("case", TC_CASES)
def test_conv_tcgen05(case):
    conv_case(BF, L.PATH_TCGEN05, **case)


# ---- the CTA-pair kernel (cta_group::2, TMA-store epilogue), forced: every operand set / tile plan it implements ----
PAIR_CASES = [
    dict(N=1, H=1, W=256, C=128, K=160, R=1),                                # one pair, one unit, two k-blocks
    dict(N=1, H=1, W=1024, C=320, K=320, R=1),                               # bias-only GEMM, bf16 out
    dict(N=1, H=1, W=1024, C=320, K=320, R=1, residual=True, res32=True, y32="both"),   # attention out-projection: fp32 stream in / out
    dict(N=1, H=1, W=1024, C=320, K=320, R=1, residual=True, res32=True, y32="only"),   # fp32 stream updated in place (no bf16 copy)
    dict(N=1, H=1, W=1000, C=320, K=960, R=1, bias=False),                   # qkv, ragged M (TMA clips the last tile)
    dict(N=1, H=1, W=384, C=320, K=320, R=1),                                # odd number of 128-row tiles: the last pair's peer CTA idles
    dict(N=1, H=1, W=200, C=640, K=640, R=1, inplace=True, alpha=0.5),       # zero-conv injection: bf16 residual == output
    dict(N=1, H=1, W=2048, C=640, K=640, R=1, inplace=True, alpha=0.5, ldy_extra=640),  # ... into a concat slot
    dict(N=1, H=1, W=256, C=320, K=320, R=1, act=L.ACT_SILU),                # SiLU epilogue
    dict(N=1, H=1, W=512, C=1280, K=320, R=1, residual=True, res32=True),    # ff2: fp32 stream in, bf16 operand out
    dict(N=2, H=32, W=32, C=320, K=320, R=3, emb=True, y32="only"),          # ResBlock conv1: timestep embedding, fp32 h
    dict(N=2, H=32, W=32, C=320, K=320, R=3, residual=True, res32=True, y32="both", ldy_extra=320),  # conv2 into a slot + fp32
    dict(N=4, H=16, W=16, C=640, K=640, R=3, residual=True),                 # 16x16: box = 8 rows
    dict(N=3, H=8, W=8, C=1280, K=1280, R=3, residual=True),                 # 8x8: box spans 2 images, ragged N
    dict(N=5, H=4, W=4, C=1280, K=1280, R=3, emb=True),                      # 4x4: box spans 8 images, one partial tile
    dict(N=1, H=64, W=64, C=320, K=320, R=3, emb=True),                      # 512^2 level: box = 2 rows of 64
    dict(N=2, H=16, W=16, C=960, K=640, R=3, ldx_extra=320, ldy_extra=640, residual=True),  # concat slices in / out
    dict(N=16, H=32, W=32, C=320, K=320, R=3, emb=True, y32="only"),         # two N = 160 sub-tiles per A tile (64 units)
    dict(N=16, H=32, W=32, C=960, K=320, R=3, residual=True, res32=True, y32="both"),   # ... deep K, fp32 in / out
    dict(N=1, H=1, W=512, C=320, K=2560, R=1, act=L.ACT_GEGLU, gb=128),      # fused GEGLU, [128 value | 128 gate] blocks
    dict(N=1, H=1, W=300, C=640, K=5120, R=1, act=L.ACT_GEGLU, gb=128),      # ... ragged M
    dict(N=1, H=1, W=66000, C=64, K=256, R=1, act=L.ACT_GEGLU, gb=128),      # more than 2^16 rows (batch 64+ at 32x32)
    dict(N=1, H=1, W=16384, C=1280, K=320, R=1, residual=True, res32=True),  # ff2 at batch 16: 3 panel slots, two sub-tiles (slot-order race)
    dict(N=16, H=8, W=8, C=1280, K=1280, R=3, emb=True, y32="only", workspace=True),    # split-K (8x8 level)
    dict(N=16, H=4, W=4, C=2560, K=1280, R=3, emb=True, workspace=True),     # split-K (4x4 level, K = 23040)
    dict(N=16, H=8, W=8, C=5120, K=1280, R=1, residual=True, res32=True, workspace=True),  # split-K ff2
    dict(N=8, H=4, W=4, C=1280, K=1280, R=3, residual=True, res32=True, y32="both", workspace=True),  # split-K + fp32
    # stride-2 Downsample convs read in place through an element-strided tensor map (no im2col, no workspace)
    dict(N=16, H=32, W=32, C=320, K=320, R=3, stride=2, y32="both"),         # 32x32 -> 16x16 at batch 16
    dict(N=3, H=16, W=16, C=640, K=640, R=3, stride=2),                      # 16x16 -> 8x8: box spans 2 images, ragged N
    dict(N=16, H=8, W=8, C=1280, K=1280, R=3, stride=2, workspace=True),     # 8x8 -> 4x4: split-K
    dict(N=2, H=64, W=64, C=320, K=320, R=3, stride=2, y32="only"),          # 512^2 level: 64x64 -> 32x32
    # ResBlock out conv + 1x1 skip projection of the block input as one contraction (second activation map, x2)
    dict(N=16, H=32, W=32, C=320, K=320, R=3, C2=960, y32="only"),           # decoder, 32x32: cat 960 -> 320
    dict(N=16, H=16, W=16, C=640, K=640, R=3, C2=320, y32="both"),           # encoder 320 -> 640
    dict(N=16, H=8, W=8, C=1280, K=1280, R=3, C2=2560, workspace=True),      # 8x8 level: split-K, a split starts inside x2's blocks
    dict(N=16, H=4, W=4, C=1280, K=1280, R=3, C2=2560, y32="only", workspace=True),  # 4x4 level
    dict(N=3, H=16, W=16, C=640, K=640, R=3, C2=1920, y32="only"),           # ragged N (box spans 2 images)
    dict(N=1, H=1, W=4096, C=640, K=320, R=1, C2=320),                       # two 1x1 terms (2-D maps)
]


@pytest.mark.parametrize("case", PAIR_CASES)
def test_conv_pair(case):
    conv_case(BF, L.PATH_TCGEN05_PAIR, **case)


def test_conv_pair_matches_single_on_hot_shapes():
    """the step's hot shapes at batch 16 through AUTO (pair kernel) and through the forced single-CTA kernel"""
    for c in (dict(N=16, H=32, W=32, C=320, K=320, R=3, emb=True, y32="only"),
              dict(N=16, H=16, W=16, C=640, K=640, R=3, residual=True, res32=True, y32="both"),
              dict(N=16, H=16, W=16, C=640, K=640, R=3, upsample=True, workspace=True),
              dict(N=16, H=32, W=32, C=320, K=320, R=3, stride=2, workspace=True, y32="both"),
              dict(N=1, H=1, W=16384, C=320, K=960, R=1, bias=False)):
        e_auto = conv_case(BF, L.PATH_AUTO, **c)
        e_single = conv_case(BF, L.PATH_TCGEN05_SINGLE, **c)
        assert abs(e_auto - e_single) < 1e-3


def test_conv_auto_dispatch():
    """hot shapes go to the tcgen05 kernel, odd ones to the generic kernel"""
    def path(**c):
        M = c["N"] * c["H"] * c["W"]
        x = torch.empty(M, c["C"], device=DEV, dtype=BF)
        R = c.get("R", 1)
        w = torch.empty(c["K"], R, R, c["C"], device=DEV, dtype=BF)
        s = c.get("stride", 1)
        y = torch.empty(M // (s * s), c["K"], device=DEV, dtype=BF)
        ws = torch.empty(64 << 20, dtype=torch.uint8, device=DEV) if c.get("ws") else None
        return ops.conv2d_path(x, w, y, N=c["N"], H=c["H"], W=c["W"], R=R, S=R, stride=s, pad=R // 2, workspace=ws)
    assert path(N=16, H=32, W=32, C=320, K=320, R=3) == L.PATH_TCGEN05
    assert path(N=16, H=1, W=1024, C=320, K=960) == L.PATH_TCGEN05
    assert path(N=16, H=32, W=32, C=4, K=320, R=3) == L.PATH_GENERIC
    assert path(N=16, H=32, W=32, C=320, K=320, R=3, stride=2, ws=True) == L.PATH_TCGEN05
    assert path(N=16, H=32, W=32, C=320, K=320, R=3, stride=2) == L.PATH_TCGEN05  # CTA-pair kernel: strided map, no workspace
    assert path(N=16, H=32, W=32, C=64, K=64, R=3, stride=2) == L.PATH_GENERIC  # single-CTA kernel: no workspace to materialise into


def test_conv_x2_term_only_where_the_pair_kernel_runs():
    """a fused second term on a shape the CTA-pair kernel declines is reported as unsupported (the caller then issues two
    launches), never silently routed to a kernel that would ignore x2"""
    def ok(N, H, W, C, K, C2, **kw):
        x, x2 = rnd(N * H * W, C, dt=BF), rnd(N * H * W, C2, dt=BF)
        w = rnd(K, 9 * C + C2, dt=BF)
        y = torch.empty(N * H * W, K, device=DEV, dtype=BF)
        return ops.conv2d_supported(x, w, y, N=N, H=H, W=W, R=3, S=3, pad=1, x2=x2, **kw)
    assert ok(16, 32, 32, 320, 320, 640)
    assert not ok(1, 8, 8, 64, 64, 64)            # K % 160 != 0: single-CTA kernel territory
    assert not ok(16, 32, 32, 320, 320, 72)       # C2 % 64 != 0
    assert not ok(1, 8, 8, 320, 320, 640)         # too few units for a cluster launch, no workspace to split along K
    with pytest.raises(RuntimeError):
        x, x2 = rnd(64, 64, dt=BF), rnd(64, 64, dt=BF)
        ops.conv2d(x, rnd(64, 9 * 64 + 64, dt=BF), torch.empty(64, 64, device=DEV, dtype=BF), N=1, H=8, W=8, R=3, S=3, pad=1, x2=x2)


def test_conv_rejects_bad_descriptors():
    x = torch.empty(64, 64, device=DEV, dtype=BF)
    w = torch.empty(64, 64, device=DEV, dtype=BF)
    with pytest.raises(RuntimeError, match="tcgen05|generic|stride"):
        ops.conv2d(x, w, x, N=1, H=8, W=8, stride=2, path=L.PATH_TCGEN05)
    with pytest.raises(ValueError):
        ops.conv2d(x, w, x[:, :32], N=1, H=8, W=8)  # 32-column view cannot hold K=64 output channels


# ---- GroupNorm statistics fused into the producing epilogue + the one-pass apply kernel ------------------------------
STATS_CASES = [
    dict(N=2, H=32, W=32, C=320, K=320, R=3, emb=True),                      # ResBlock conv1 -> h (fp32 only) + stats
    dict(N=3, H=16, W=16, C=640, K=640, R=3, residual=True, slot=(640, 1280)),  # conv2 into a concat slot + stats slice
    dict(N=2, H=32, W=32, C=320, K=320, R=1, inplace=True),                  # zero-conv injection re-emitting stats
    dict(N=1, H=64, W=64, C=320, K=320, R=1, residual=True),                 # 512^2 level
    dict(N=2, H=16, W=16, C=1280, K=640, R=3, upsample=False, emb=True),     # C_in != C_out
    dict(N=2, H=32, W=32, C=320, K=320, R=3, stride=2),                      # Downsample (im2col + GEMM), 16x16 out
    dict(N=2, H=8, W=8, C=640, K=640, R=3, up=True),                         # Upsample 8 -> 16
    dict(N=8, H=64, W=64, C=128, K=256, R=3, residual=True),                 # N tile 256 (VAE decoder widths)
    dict(N=4, H=64, W=64, C=256, K=128, R=3),                                # N tile 128
    dict(N=1, H=256, W=256, C=128, K=128, R=3),                              # 512 row tiles per sample
]


PAIR_STATS_CASES = [
    dict(N=2, H=32, W=32, C=320, K=320, R=3, emb=True),                      # conv1 -> h + stats
    dict(N=4, H=16, W=16, C=640, K=640, R=3, residual=True, slot=(640, 1280)),  # conv2 into a concat slot + stats slice
    dict(N=2, H=32, W=32, C=320, K=320, R=1, inplace=True),                  # zero-conv injection re-emitting stats
    dict(N=1, H=64, W=64, C=320, K=320, R=1, residual=True),                 # proj_out, 512^2 level
    dict(N=16, H=32, W=32, C=320, K=320, R=3, residual=True),                # two sub-tiles per A tile
    dict(N=3, H=16, W=16, C=1280, K=640, R=3, emb=True),                     # odd tile count
    dict(N=16, H=32, W=32, C=320, K=320, R=3, stride=2),                     # Downsample in place (element-strided map) + stats
    dict(N=16, H=32, W=32, C=320, K=320, R=3, C2=640),                       # out conv + skip projection + stats
]


@pytest.mark.parametrize("case", PAIR_STATS_CASES)
def test_conv_pair_stats_and_groupnorm_apply(case):
    test_conv_stats_and_groupnorm_apply(case, path=L.PATH_TCGEN05_PAIR)


@pytest.mark.parametrize("case", STATS_CASES)
def test_conv_stats_and_groupnorm_apply(case, path=L.PATH_AUTO):
    """conv2d(stats=) must emit, per 128-row tile and channel, the (sum, sumsq) of exactly the fp32 values it stored;
    groupnorm_apply on those statistics must equal F.group_norm of the stored tensor."""
    N, H, W, C, K, R = (case[k] for k in "NHWCKR")
    stride, up = case.get("stride", 1), case.get("up", False)
    Ho, Wo = (2 * H, 2 * W) if up else (H // stride, W // stride)
    M, Mo = N * H * W, N * Ho * Wo
    x = rnd(M, C, dt=BF, seed=1)
    w = (rnd(K, R, R, C, dt=F32, seed=2) / math.sqrt(C * R * R)).to(BF)
    b = rnd(K, seed=3)
    x2 = None
    if case.get("C2"):  # fused second 1x1 term
        x2 = rnd(M, case["C2"], dt=BF, seed=9)
        w = torch.cat([w.reshape(K, -1), (rnd(K, case["C2"], seed=10) / math.sqrt(case["C2"])).to(BF)], 1).contiguous()
    e = rnd(N, K, dt=BF, seed=4) if case.get("emb") else None
    off, width = case.get("slot", (0, K))
    yb = rnd(Mo, width, dt=BF, seed=5)
    y = yb[:, off:off + K]
    y32 = torch.empty(Mo, K, device=DEV)
    res = y if case.get("inplace") else (rnd(Mo, K, dt=F32, seed=6) if case.get("residual") else None)
    stb = torch.full((Mo // 128, width, 2), -7.0, device=DEV)
    st = stb[:, off:off + K, :]
    ws = torch.empty(64 << 20, dtype=torch.uint8, device=DEV)
    ops.conv2d(x, w, y, N=N, H=H, W=W, R=R, S=R, stride=stride, pad=R // 2, upsample=up, bias=b, emb=e, residual=res,
               alpha=0.5 if case.get("inplace") else 1.0, workspace=ws, y32=y32, stats=st, path=path, x2=x2)
    t = y32.reshape(Mo // 128, 128, K)
    ref = torch.stack([t.sum(1), (t * t).sum(1)], -1)
    assert rel(st[..., 0], ref[..., 0]) < 1e-5 and rel(st[..., 1], ref[..., 1]) < 1e-5
    if off:  # columns outside the slice untouched
        assert bool((stb[:, :off] == -7.0).all())
    # determinism: a second run writes bit-identical statistics
    st1 = st.clone()
    if not case.get("inplace"):
        ops.conv2d(x, w, y, N=N, H=H, W=W, R=R, S=R, stride=stride, pad=R // 2, upsample=up, bias=b, emb=e, residual=res,
                   workspace=ws, y32=y32, stats=st, path=path, x2=x2)
        assert torch.equal(st, st1)
    for silu in (True, False):
        for src in (y32, y):
            g, bt = rnd(K, seed=7) + 1.0, rnd(K, seed=8)
            out = torch.empty(Mo, K, device=DEV, dtype=BF)
            ops.groupnorm_apply(src, out, Nn := N, g, bt, 1e-5, silu, st)
            r = F.group_norm(y32.reshape(N, Ho * Wo, K).permute(0, 2, 1), 32, g, bt, 1e-5)
            r = (F.silu(r) if silu else r).permute(0, 2, 1).reshape(Mo, K)
            assert rel(out, r) < 6e-3, (silu, src.dtype, rel(out, r))


def test_conv_stats_rejected_on_generic_path():
    x = rnd(256, 4, dt=BF)
    w = rnd(64, 3, 3, 4, dt=BF)
    y = torch.empty(256, 64, device=DEV, dtype=BF)
    with pytest.raises(RuntimeError, match="statistics|stats"):
        ops.conv2d(x, w, y, N=1, H=16, W=16, R=3, S=3, pad=1, stats=torch.zeros(2, 64, 2, device=DEV))


# ---- attention -------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dt", [BF, F32])
@pytest.mark.parametrize("B,heads,Nq,Nkv,d", [(2, 8, 1024, 1024, 40), (2, 8, 256, 256, 80), (3, 8, 64, 64, 160),
                                               (2, 8, 16, 16, 160), (2, 8, 1024, 77, 40), (2, 8, 64, 77, 160),
                                               (1, 4, 100, 77, 16), (1, 4, 64, 64, 32), (2, 2, 200, 130, 64),
                                               (1, 8, 4096, 4096, 40), (2, 8, 300, 300, 8), (1, 3, 513, 257, 96),
                                               (1, 2, 129, 128, 128), (2, 1, 256, 256, 512),  # last: the VAE decoder's single 512-wide head (SIMT kernel)
                                               # split-KV kernel (head dims <= 64): one half only, odd halves, ragged halves
                                               (2, 4, 128, 40, 40), (2, 4, 130, 64, 48), (1, 8, 256, 192, 40), (1, 2, 70, 321, 64),
                                               (1, 8, 1024, 1000, 40)])
def test_attention(dt, B, heads, Nq, Nkv, d):
    C = heads * d
    self_attn = Nq == Nkv
    if self_attn:
        qkv = rnd(B * Nq, 3 * C, dt=dt, seed=1)
        q, k, v = qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:]
    else:
        q = rnd(B * Nq, C, dt=dt, seed=1)
        kv = rnd(B * Nkv, 2 * C, dt=dt, seed=2)
        k, v = kv[:, :C], kv[:, C:]
    o = torch.empty(B * Nq, C, device=DEV, dtype=dt)
    ops.attention(q, k, v, o, B=B, heads=heads, Nq=Nq, Nkv=Nkv, d=d, scale=d ** -0.5)
    sp = lambda t, n: t.float().reshape(B, n, heads, d).permute(0, 2, 1, 3)  # noqa: E731
    ref = F.scaled_dot_product_attention(sp(q, Nq), sp(k, Nkv), sp(v, Nkv)).permute(0, 2, 1, 3).reshape(B * Nq, C)
    assert rel(o, ref) < (1e-2 if dt == BF else 2e-5)


# ---- weight groups (mkd_conv_desc.wgroups / the norms' wgroups): the same layer of two networks on two stacked batches --------
WG_CASES = [
    dict(N=32, H=32, W=32, C=320, K=320, R=3, emb=True, y32="both", stats=True),      # ResBlock conv1, 32x32 level
    dict(N=32, H=32, W=32, C=320, K=320, R=1, res32=True, y32="only"),                # attention out-projection into the fp32 stream
    dict(N=32, H=16, W=16, C=640, K=5120, R=1, act="geglu"),                          # FF1 GEGLU
    dict(N=32, H=4, W=4, C=1280, K=1280, R=3, emb=True, y32="only", workspace=True),  # 4x4 level: split-K + reducer
    dict(N=32, H=16, W=16, C=640, K=640, R=3, C2=320, y32="both", stats=True),        # out conv + fused skip projection
    dict(N=32, H=32, W=32, C=320, K=320, R=3, stride=2, y32="both"),                  # Downsample in place
    dict(N=32, H=8, W=8, C=1280, K=3840, R=1),                                        # q/k/v at the 8x8 level
    dict(N=2, H=16, W=16, C=640, K=640, R=1),                                         # 2 x 4 units: still one launch (not two single-CTA ones)
    dict(N=32, H=4, W=4, C=1280, K=1280, R=1, res32=True, y32="only"),                # middle block's attention out-projection: 2 x 8 units
    dict(N=32, H=4, W=4, C=1280, K=1280, R=1, y32="both"),                            # middle block's proj_in
    dict(N=2, H=8, W=8, C=640, K=640, R=1, fallback=True),                            # parts that are not whole tile pairs: two launches
]


@pytest.mark.parametrize("case", WG_CASES)
def test_conv_weight_groups(case):
    """conv2d_grouped == the two networks' launches on their own halves (same kernel, so the same arithmetic up to the split-K
    plan), and the grouped launch really is ONE launch where the kernel takes it"""
    N, H, W, C, K, R = (case[k] for k in "NHWCKR")
    stride, C2 = case.get("stride", 1), case.get("C2", 0)
    geglu = case.get("act") == "geglu"
    M, Mo = N * H * W, N * (H // stride) * (W // stride)
    Ko = K // 2 if geglu else K
    x = rnd(M, C, dt=BF, seed=1)
    w = (rnd(2 * K, R * R * C + C2, seed=2) / math.sqrt(C * R * R + C2)).to(BF)
    b = 0.5 * rnd(2 * K, seed=3)
    kw = dict(H=H, W=W, R=R, S=R, stride=stride, pad=R // 2, bias=b)
    if case.get("emb"):
        kw["emb"] = rnd(N, K, dt=BF, seed=4)
    if C2:
        kw["x2"] = rnd(M, C2, dt=BF, seed=5)
    if case.get("workspace"):
        kw["workspace"] = torch.empty(64 << 20, dtype=torch.uint8, device=DEV)
    if geglu:
        kw.update(act=L.ACT_GEGLU, geglu_block=128)
    res = rnd(Mo, Ko, dt=F32, seed=6) if case.get("res32") else None
    outs = []
    for grouped in (True, False):
        y = None if case.get("y32") == "only" else torch.zeros(Mo, Ko, device=DEV, dtype=BF)
        y32 = torch.zeros(Mo, Ko, device=DEV) if case.get("y32") else None
        st = torch.zeros(Mo // 128, Ko, 2, device=DEV) if case.get("stats") else None
        k2 = dict(kw, y32=y32, stats=st, residual=None if res is None else res.clone())
        n0 = L.load().mkd_launch_count()
        if grouped:
            ops.conv2d_grouped(x, w, y, N=N, **k2)
        else:
            for g in (0, 1):
                part = lambda t, n: None if t is None else t[g * n:(g + 1) * n]  # noqa: E731
                k3 = dict(k2, bias=part(b, K), emb=part(k2.get("emb"), N // 2), x2=part(k2.get("x2"), M // 2), y32=part(y32, Mo // 2),
                          residual=part(k2["residual"], Mo // 2), stats=part(st, Mo // 256))
                ops.conv2d(part(x, M // 2), w[g * K:(g + 1) * K], part(y, Mo // 2), N=N // 2, **k3)
        launches = L.load().mkd_launch_count() - n0
        outs.append((y, y32, st, launches))
    (y1, y321, st1, l1), (y0, y320, st0, l0) = outs
    if not case.get("fallback"):
        assert l1 * 2 == l0, (l1, l0)       # one launch (+ one reducer) instead of two (+ two)
    else:
        assert l1 == l0
    for a, r in ((y1, y0), (y321, y320)):
        if a is not None:
            assert rel(a, r.float()) < (3e-3 if case.get("workspace") else 1e-6), rel(a, r.float())
            assert float(r.float().abs().max()) > 0
    if st1 is not None:
        assert rel(st1, st0) < 1e-5


GN_TAIL_CASES = [
    dict(N=16, H=4, W=4, C=1280, K=1280, emb=True, y32="only"),                 # decoder ResBlock conv1 at the 4x4 level (9 splits)
    dict(N=16, H=8, W=8, C=2560, K=1280, emb=True, y32="only"),                 # ... at the 8x8 level (2 splits), 64 x 5 vectors per block
    dict(N=32, H=4, W=4, C=1280, K=1280, emb=True, y32="only", wgroups=2),      # stacked trunk: gamma / beta per network
    dict(N=16, H=8, W=8, C=1280, K=1280, C2=2560, y32="both", silu=False, eps=1e-6),  # out conv + skip projection -> SpatialTransformer norm
    dict(N=16, H=32, W=32, C=320, K=320, emb=True, y32="only", declined=True),   # no split-K at this level: the caller keeps its GroupNorm launch
]


@pytest.mark.parametrize("case", GN_TAIL_CASES)
def test_conv_groupnorm_tail(case):
    """mkd_conv_desc.gn_y: the split-K reducer also writes GroupNorm(+SiLU) of the layer's output == the conv followed by
    mkd_groupnorm on its fp32 output; y / y32 unchanged; one launch fewer; shapes that do not split are declined"""
    N, H, W, C, K = (case[k] for k in "NHWCK")
    wg, C2 = case.get("wgroups", 1), case.get("C2", 0)
    M = N * H * W
    x = rnd(M, C, dt=BF, seed=1)
    w = (rnd(wg * K, 9 * C + C2, seed=2) / math.sqrt(9 * C + C2)).to(BF)
    b = 0.5 * rnd(wg * K, seed=3)
    gamma, beta = 1.0 + 0.3 * rnd(wg * K, seed=7), 0.2 * rnd(wg * K, seed=8)
    kw = dict(N=N, H=H, W=W, R=3, S=3, pad=1, bias=b, workspace=torch.empty(64 << 20, dtype=torch.uint8, device=DEV), wgroups=wg)
    if case.get("emb"):
        kw["emb"] = rnd(N, K, dt=BF, seed=4)
    if C2:
        kw["x2"] = rnd(M, C2, dt=BF, seed=5)
    silu, eps = case.get("silu", True), case.get("eps", 1e-5)

    def outs():
        return (torch.zeros(M, K, device=DEV, dtype=BF) if case["y32"] == "both" else None), torch.zeros(M, K, device=DEV)
    y, y32 = outs()
    g = torch.zeros(M, K, device=DEV, dtype=BF)
    gn = dict(y=g, gamma=gamma, beta=beta, eps=eps, silu=silu)
    ok = ops.conv2d_supported(x, w, y, y32=y32, gn=gn, **kw)
    assert ok == (not case.get("declined"))
    if not ok:
        with pytest.raises(RuntimeError):
            ops.conv2d(x, w, y, y32=y32, gn=gn, **kw)
        return
    n0 = L.load().mkd_launch_count()
    ops.conv2d(x, w, y, y32=y32, gn=gn, **kw)
    fused_launches = L.load().mkd_launch_count() - n0
    yr, y32r = outs()
    gr = torch.zeros(M, K, device=DEV, dtype=BF)
    n0 = L.load().mkd_launch_count()
    ops.conv2d(x, w, yr, y32=y32r, **kw)
    ops.groupnorm(y32r, gr, N, gamma, beta, eps, silu, torch.empty(ops.groupnorm_workspace_bytes(N), dtype=torch.uint8, device=DEV), wgroups=wg)
    assert L.load().mkd_launch_count() - n0 > fused_launches == 2          # GEMM + reducer; the GroupNorm launch(es) are gone
    assert torch.equal(y32, y32r) and (y is None or torch.equal(y, yr))    # the layer's own outputs: same sums in the same order
    assert float(gr.float().abs().max()) > 0.5
    assert rel(g, gr.float()) < 4e-3, rel(g, gr.float())                    # both round to bf16 once; statistics in fp32 either way
    # against torch on the fp32 output
    ref = torch.empty_like(y32r)
    for p in range(wg):
        rows = slice(p * M // wg, (p + 1) * M // wg)
        t = F.group_norm(y32r[rows].reshape(N // wg, H * W, K).permute(0, 2, 1), 32, gamma[p * K:(p + 1) * K], beta[p * K:(p + 1) * K], eps)
        ref[rows] = (F.silu(t) if silu else t).permute(0, 2, 1).reshape(-1, K)
    assert rel(g, ref) < 4e-3, rel(g, ref)


@pytest.mark.parametrize("N,HW,C", [(32, 1024, 320), (32, 256, 640), (32, 64, 1280), (32, 16, 2560), (4, 100, 128)])
def test_norm_weight_groups(N, HW, C):
    """GroupNorm (all three kernels), GroupNorm-apply and LayerNorm with wgroups = 2 == the two halves with their own gamma / beta,
    bit for bit"""
    x = rnd(N * HW, C, dt=F32, seed=1) * 1.5 + 0.2
    g2, b2 = 1 + 0.2 * rnd(2, C, seed=2), 0.1 * rnd(2, C, seed=3)
    ws = torch.empty(ops.groupnorm_workspace_bytes(N) // 4, device=DEV)
    h = N * HW // 2

    def both(fn):
        a, r = torch.empty(N * HW, C, device=DEV, dtype=BF), torch.empty(N * HW, C, device=DEV, dtype=BF)
        fn(x, a, g2, b2, 2, N)
        for g in (0, 1):
            fn(x[g * h:(g + 1) * h], r[g * h:(g + 1) * h], g2[g].contiguous(), b2[g].contiguous(), 1, N // 2)
        assert torch.equal(a, r)
    both(lambda xx, yy, ga, be, wg, n: ops.groupnorm(xx, yy, n, ga, be, 1e-5, True, ws, wgroups=wg))
    if C <= 1280:
        both(lambda xx, yy, ga, be, wg, n: ops.layernorm(xx, yy, ga, be, wgroups=wg))
    if HW % 128 == 0:
        t = x.reshape(N * HW // 128, 128, C)
        st = torch.stack([t.sum(1), (t * t).sum(1)], -1).contiguous()
        both(lambda xx, yy, ga, be, wg, n: ops.groupnorm_apply(xx, yy, n, ga, be, 1e-5, True,
                                                               st[:xx.shape[0] // 128] if xx.data_ptr() == x.data_ptr() else st[h // 128:], wgroups=wg))
