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
