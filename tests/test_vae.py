"""First-stage (VAE) decoder — SURVEY.md §8(f) rank 1: oracle structural pins, host logic over the torch fakes (CPU),
and parity of the CUDA path against the oracle on a B200 (-m gpu)."""
import pytest
import torch

import fake_ops
from makeupdiffuse_b200 import B200ControlLDM, B200FirstStageDecoder, ops
from oracle import seeded_state_dict
from oracle.vae import OracleFirstStageDecoder, decode_first_stage

TINY_DD = dict(ch=64, ch_mult=(1, 2), num_res_blocks=1)   # reduced widths / depth, same op sequence


def rel(a, b):
    return float((a.float() - b.float()).norm() / b.float().norm())


def psnr(a, b):
    import math
    peak = float(b.abs().max())
    return 10 * math.log10(peak * peak / float(((a.float() - b.float()) ** 2).mean()))


def test_oracle_decoder_matches_the_public_sd_vae_structure():
    m = OracleFirstStageDecoder()
    assert sum(p.numel() for p in m.decoder.parameters()) == 49_490_179      # public SD-1.x VAE decoder
    assert sum(p.numel() for p in m.post_quant_conv.parameters()) == 20
    keys = set(m.state_dict())
    for k in ("post_quant_conv.weight", "decoder.conv_in.weight", "decoder.mid.attn_1.proj_out.bias",
              "decoder.up.3.upsample.conv.weight", "decoder.up.1.block.0.nin_shortcut.weight",
              "decoder.up.0.block.2.conv2.bias", "decoder.norm_out.weight", "decoder.conv_out.weight"):
        assert k in keys, k
    assert "decoder.up.0.upsample.conv.weight" not in keys                    # the full-resolution level has no upsample
    assert m.state_dict()["decoder.up.1.block.0.conv1.weight"].shape == (256, 512, 3, 3)
    # the B200 module expects exactly the same keys and shapes
    b = B200FirstStageDecoder()
    assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == b.upstream_shapes()
    z = torch.randn(1, 4, 8, 8)
    with torch.no_grad():
        assert m.decode(z).shape == (1, 3, 64, 64)


@pytest.fixture()
def faked(monkeypatch):
    for name in fake_ops.ALL:
        monkeypatch.setattr(ops, name, getattr(fake_ops, name))


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-5), (torch.bfloat16, 2e-2)])
def test_decoder_plumbing_on_cpu_fakes(faked, dtype, tol):
    o = OracleFirstStageDecoder(ddconfig=TINY_DD).eval()
    sd = seeded_state_dict(o, 0, prefix="first_stage_model.")
    m = B200FirstStageDecoder(ddconfig=TINY_DD, dtype=dtype).load_state_dict(sd, device="cpu")
    z = torch.randn(2, 4, 16, 16, generator=torch.Generator().manual_seed(3))
    with torch.no_grad():
        ref = o.decode(z)
    out = m.decode(z)
    assert out.shape == ref.shape == (2, 3, 32, 32) and out.dtype == torch.float32
    assert rel(out, ref) < tol, rel(out, ref)
    # decode_first_stage on the LDM object applies 1 / scale_factor (makeups.py:260-262)
    ldm = B200ControlLDM({"model_channels": 64, "num_heads": 4, "context_dim": 64},
                         {"model_channels": 64, "num_heads": 4, "context_dim": 64}, dtype=dtype, device="cpu")
    with pytest.raises(RuntimeError):
        ldm.decode_first_stage(z)
    ldm.attach_first_stage_decoder(m)
    with torch.no_grad():
        assert rel(ldm.decode_first_stage(z), decode_first_stage(o, z)) < tol
    with pytest.raises(KeyError):
        B200FirstStageDecoder(ddconfig=TINY_DD, dtype=dtype).load_state_dict({}, device="cpu")


@pytest.mark.gpu
@pytest.mark.parametrize("dd,B,h", [(TINY_DD, 2, 16), (None, 1, 32)])
def test_decoder_parity_on_b200(dd, B, h):
    """tiny widths (fast) and the yaml-sized decoder at 256^2 (49.5 M parameters): fp32 check mode <= 1e-4 rel-L2,
    bf16 image PSNR >= 30 dB against the fp32 oracle on the same latents and weights"""
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    with torch.device("cuda"):
        o = OracleFirstStageDecoder(ddconfig=dd).eval()
    sd = seeded_state_dict(o, 0, prefix="first_stage_model.")
    z = torch.randn(B, 4, h, h, device="cuda", generator=torch.Generator(device="cuda").manual_seed(5))
    with torch.no_grad():
        ref = o.decode(z)
    f32 = B200FirstStageDecoder(ddconfig=dd, dtype=torch.float32).load_state_dict(sd).decode(z)
    b16 = B200FirstStageDecoder(ddconfig=dd, dtype=torch.bfloat16).load_state_dict(sd).decode(z)
    r32, r16, p16 = rel(f32, ref), rel(b16, ref), psnr(b16, ref)
    print(f"VAE decode {tuple(ref.shape)}: rel-L2 fp32-check {r32:.2e}  bf16 {r16:.2e}  bf16 PSNR {p16:.1f} dB")
    assert r32 < 1e-4 and r16 < 2e-2 and p16 > 30, (r32, r16, p16)
    again = B200FirstStageDecoder(ddconfig=dd, dtype=torch.bfloat16).load_state_dict(sd).decode(z)
    assert torch.equal(again, b16)  # deterministic


# ---- encode side: the x_p entry ----------------------------------------------------------------
from makeupdiffuse_b200 import B200FirstStageEncoder  # noqa: E402
from oracle.vae import OracleFirstStageEncoder, get_z  # noqa: E402

TINY_ENC = dict(ch=64, ch_mult=(1, 2), num_res_blocks=1)


def test_oracle_encoder_matches_the_public_sd_vae_structure():
    m = OracleFirstStageEncoder()
    assert sum(p.numel() for p in m.encoder.parameters()) == 34_163_592      # public SD-1.x VAE encoder
    assert sum(p.numel() for p in m.quant_conv.parameters()) == 72
    keys = set(m.state_dict())
    for k in ("quant_conv.weight", "encoder.conv_in.weight", "encoder.down.0.downsample.conv.weight",
              "encoder.down.1.block.0.nin_shortcut.weight", "encoder.mid.attn_1.q.weight", "encoder.conv_out.bias"):
        assert k in keys, k
    assert "encoder.down.3.downsample.conv.weight" not in keys                # the last level has no downsample
    assert m.state_dict()["encoder.conv_out.weight"].shape == (8, 512, 3, 3)  # double_z: mean and logvar
    assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == B200FirstStageEncoder().upstream_shapes()
    # Downsample = pad bottom / right by one, stride-2 conv without padding: an 8x reduction overall
    with torch.no_grad():
        mean, logvar = m.encode(torch.randn(1, 3, 64, 64))
    assert mean.shape == logvar.shape == (1, 4, 8, 8) and float(logvar.max()) <= 20.0


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-5), (torch.bfloat16, 2e-2)])
def test_encoder_plumbing_on_cpu_fakes(faked, dtype, tol):
    o = OracleFirstStageEncoder(ddconfig=TINY_ENC).eval()
    sd = seeded_state_dict(o, 0, prefix="first_stage_model.")
    m = B200FirstStageEncoder(ddconfig=TINY_ENC, dtype=dtype).load_state_dict(sd, device="cpu")
    g = torch.Generator().manual_seed(4)
    x = torch.rand(2, 3, 32, 32, generator=g) * 2 - 1
    with torch.no_grad():
        rm, rl = o.encode(x)
    mean, logvar = m.encode(x)
    assert mean.shape == rm.shape == (2, 4, 16, 16) and mean.dtype == torch.float32
    assert rel(mean, rm) < tol and rel(logvar, rl) < tol, (rel(mean, rm), rel(logvar, rl))
    ldm = B200ControlLDM({"model_channels": 64, "num_heads": 4, "context_dim": 64},
                         {"model_channels": 64, "num_heads": 4, "context_dim": 64}, dtype=dtype, device="cpu")
    with pytest.raises(RuntimeError):
        ldm.get_z(x)
    ldm.attach_first_stage_encoder(m)
    noise = torch.randn(2, 4, 16, 16, generator=g)
    with torch.no_grad():
        assert rel(ldm.get_z(x, noise), get_z(o, x, noise)) < tol        # makeup_diffuse.py:37-40
    assert ldm.get_z(x).shape == (2, 4, 16, 16)                          # noise drawn internally


@pytest.mark.gpu
@pytest.mark.parametrize("dd,B,hw", [(TINY_ENC, 2, 64), (None, 1, 256)])
def test_encoder_parity_on_b200(dd, B, hw):
    """tiny widths and the yaml-sized encoder on a 256^2 image (34.2 M parameters): posterior moments and the latent
    z = scale_factor * sample against the fp32 oracle; fp32 check mode <= 1e-4, bf16 <= 2e-2 rel-L2"""
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    with torch.device("cuda"):
        o = OracleFirstStageEncoder(ddconfig=dd).eval()
    sd = seeded_state_dict(o, 0, prefix="first_stage_model.")
    g = torch.Generator(device="cuda").manual_seed(6)
    x = torch.rand(B, 3, hw, hw, device="cuda", generator=g) * 2 - 1
    noise = torch.randn(B, 4, hw // (2 ** (len((dd or {}).get("ch_mult", (1, 2, 4, 4))) - 1)),
                        hw // (2 ** (len((dd or {}).get("ch_mult", (1, 2, 4, 4))) - 1)), device="cuda", generator=g)
    with torch.no_grad():
        rm, rl = o.encode(x)
        rz = get_z(o, x, noise)
    res = {}
    for name, dt in (("f32", torch.float32), ("bf16", torch.bfloat16)):
        enc = B200FirstStageEncoder(ddconfig=dd, dtype=dt).load_state_dict(sd)
        mean, logvar = enc.encode(x)
        z = 0.18215 * (mean + torch.exp(0.5 * logvar) * noise)
        res[name] = (rel(mean, rm), rel(logvar, rl), rel(z, rz))
    print(f"VAE encode {tuple(x.shape)}: rel-L2 (mean, logvar, z) fp32-check {res['f32']}  bf16 {res['bf16']}")
    assert max(res["f32"]) < 1e-4 and max(res["bf16"]) < 2e-2, res


@pytest.mark.gpu
def test_x_p_entry_reconstruct_on_b200():
    """the teacher-conditioned flow end to end on the B200 path (diffusion_makeup.py:384-387 + cddim.py:81-100):
    image -> get_z -> q_sample(t) -> reconstruct(t_start) -> decode_first_stage, against the oracle, tiny networks"""
    from makeupdiffuse_b200 import B200DDIMSampler
    from oracle import MKDDIMSampler, OracleControlLDM
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    params = dict(model_channels=64, num_heads=4, context_dim=64)
    dd = dict(ch=64, ch_mult=(1, 2, 4, 4), num_res_blocks=1)
    with torch.device("cuda"):
        ol = OracleControlLDM(control_params=params, unet_params=params).eval()
        oe, od = OracleFirstStageEncoder(ddconfig=dd).eval(), OracleFirstStageDecoder(ddconfig=dd).eval()
    sd = seeded_state_dict(ol, 0)
    sde, sdd = seeded_state_dict(oe, 0, prefix="first_stage_model."), seeded_state_dict(od, 0, prefix="first_stage_model.")
    m = B200ControlLDM(params, params, dtype=torch.float32).load_state_dict(sd)
    m.attach_first_stage_encoder(B200FirstStageEncoder(ddconfig=dd, dtype=torch.float32).load_state_dict(sde))
    m.attach_first_stage_decoder(B200FirstStageDecoder(ddconfig=dd, dtype=torch.float32).load_state_dict(sdd))
    g = torch.Generator(device="cuda").manual_seed(8)
    B, hw, S, t_start = 2, 128, 10, 6
    pgt = torch.rand(B, 3, hw, hw, device="cuda", generator=g) * 2 - 1              # teacher output x_p in [-1, 1]
    cond = {"c_crossattn": [torch.randn(B, 77, 64, device="cuda", generator=g)],
            "c_concat": [torch.rand(B, 6, hw, hw, device="cuda", generator=g)]}
    n1 = torch.randn(B, 4, hw // 8, hw // 8, device="cuda", generator=g)
    n2 = torch.randn(B, 4, hw // 8, hw // 8, device="cuda", generator=g)
    so, sb = MKDDIMSampler(ol), B200DDIMSampler(m, use_cuda_graph=False)
    so.make_schedule(S, ddim_eta=0.0, verbose=False)
    sb.make_schedule(S, ddim_eta=0.0, verbose=False)
    t = torch.full((B,), int(so.ddim_timesteps[t_start - 1]), device="cuda", dtype=torch.long)
    with torch.no_grad():
        z_ref = get_z(oe, pgt, n1)
        x_ref = so.reconstruct(ol.q_sample(z_ref, t, n2), cond, t_start=t_start)
        img_ref = decode_first_stage(od, x_ref)
    z = m.get_z(pgt, n1)
    x = sb.reconstruct(m.q_sample(z, t, n2), cond, t_start=t_start)
    img = m.decode_first_stage(x)
    r = (rel(z, z_ref), rel(x, x_ref), rel(img, img_ref))
    print(f"x_p entry (fp32 check mode): rel-L2 z {r[0]:.2e}, reconstructed latents {r[1]:.2e}, decoded image {r[2]:.2e}")
    assert max(r) < 1e-3, r
