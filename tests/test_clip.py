"""Conditioning producer: the CLIP text encoder behind get_learned_conditioning.

CPU: the oracle restatement against the REAL implementation the reference wraps (transformers.CLIPTextModel, importable
in this image) at the full ViT-L/14 text size; the product's host logic (weight repacking, quick_gelu folding, in-place
residual stream) against the oracle with the kernel contracts restated by tests/fake_ops.py.
GPU: B200FrozenCLIPEmbedder (embedding kernel, tcgen05 GEMMs, causal attention kernel, LayerNorm) against the oracle.
"""
import pytest
import torch

import fake_ops
from makeupdiffuse_b200 import ops
from makeupdiffuse_b200.clip import B200FrozenCLIPEmbedder
from oracle.clip import BOS, CLIP_L_TEXT, EOS, OracleCLIPTextEncoder, empty_prompt_tokens

SMALL = dict(vocab_size=1000, hidden_size=64, num_hidden_layers=2, num_attention_heads=4, intermediate_size=128)


def rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return float((a - b).norm() / b.norm())


def seeded_oracle(seed=0, **cfg):
    torch.manual_seed(seed)
    m = OracleCLIPTextEncoder(**cfg).eval()
    with torch.no_grad():  # biases / LayerNorm parameters away from their 0 / 1 defaults so that every term is exercised
        for n, p in m.named_parameters():
            if n.endswith("bias"):
                p.normal_(0.0, 0.05)
            elif "layer_norm" in n and n.endswith("weight"):
                p.normal_(1.0, 0.1)
    return m


def tokens(B, T, vocab, seed=1):
    g = torch.Generator().manual_seed(seed)
    t = torch.randint(0, vocab - 2, (B, T), generator=g)
    t[:, 0] = min(BOS, vocab - 2)
    t[:, T // 2:] = min(EOS, vocab - 1)  # padded tail, like the tokenizer's max_length padding
    return t


def test_oracle_matches_transformers_clip_text_model_full_size():
    tr = pytest.importorskip("transformers")
    torch.set_num_threads(4)
    cfg = tr.CLIPTextConfig(**CLIP_L_TEXT, hidden_act="quick_gelu")
    torch.manual_seed(0)
    hf = tr.CLIPTextModel(cfg).eval()
    o = OracleCLIPTextEncoder().eval()
    sd = {k: v for k, v in hf.state_dict().items() if not k.endswith("position_ids")}
    o.load_state_dict(sd)
    assert sum(p.numel() for p in o.parameters()) == 123_060_480  # the published ViT-L/14 text tower size
    tok = torch.cat([tokens(1, 77, 49408), empty_prompt_tokens(1)])
    with torch.no_grad():
        ref = hf(input_ids=tok).last_hidden_state
        got = o(tok)
    assert ref.shape == (2, 77, 768)
    assert float((got - ref).abs().max()) < 2e-5 and rel(got, ref) < 1e-5


def test_causality_and_empty_prompt_ids():
    o = seeded_oracle(**SMALL)
    t1 = tokens(1, 20, 1000)
    t2 = t1.clone()
    t2[0, 12:] = 7  # changing later tokens must not change earlier positions
    a, b = o(t1), o(t2)
    assert torch.equal(a[0, :12], b[0, :12]) and not torch.allclose(a[0, 12:], b[0, 12:])
    e = empty_prompt_tokens(3)
    assert e.shape == (3, 77) and e[0, 0] == BOS and bool((e[:, 1:] == EOS).all())


def test_product_host_logic_matches_oracle(monkeypatch):
    for name in fake_ops.ALL:
        monkeypatch.setattr(ops, name, getattr(fake_ops, name))
    o = seeded_oracle(**SMALL)
    enc = B200FrozenCLIPEmbedder(device="cpu", dtype=torch.float32, **SMALL)
    sd = {"cond_stage_model.transformer." + k: v for k, v in o.state_dict().items()}  # checkpoint-style prefix
    enc.load_state_dict(sd)
    tok = tokens(3, 77, 1000)
    assert rel(enc(tok), o(tok)) < 2e-6
    with pytest.raises(KeyError):
        enc.load_state_dict({**sd, "cond_stage_model.transformer.text_model.extra.weight": torch.zeros(1)})
    with pytest.raises(RuntimeError):
        enc.encode(["makeup transfer"])  # no vocabulary in this image: needs tokenize=
    with pytest.raises(IndexError):
        enc(torch.full((1, 77), 1000))


@pytest.mark.gpu
@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-4), (torch.bfloat16, 2e-2)])
def test_clip_text_encoder_parity_on_b200(dtype, tol):
    torch.set_num_threads(8)
    o = seeded_oracle()
    enc = B200FrozenCLIPEmbedder(device="cuda", dtype=dtype).load_state_dict(o.state_dict())
    tok = torch.cat([tokens(3, 77, 49408), empty_prompt_tokens(1)])
    ref = o(tok)
    got = enc(tok.cuda())
    assert got.shape == (4, 77, 768) and got.dtype == torch.float32
    e = rel(got, ref)
    print(f"CLIP text encoder {dtype}: rel-L2 {e:.2e}")
    assert e < tol
    e2 = rel(enc.encode(["", ""]), o(empty_prompt_tokens(2)))
    assert e2 < tol


@pytest.mark.gpu
def test_clip_small_and_short_sequences_on_b200():
    o = seeded_oracle(**SMALL)
    enc = B200FrozenCLIPEmbedder(device="cuda", dtype=torch.float32, **SMALL).load_state_dict(o.state_dict())
    for B, T in ((1, 77), (5, 20), (2, 1)):
        tok = tokens(B, T, 1000, seed=B)
        assert rel(enc(tok.cuda()), o(tok)) < 1e-4


@pytest.mark.gpu
def test_get_learned_conditioning_feeds_apply_model_on_b200(tiny_params):
    """prompt -> cond_stage_model -> c_crossattn, cached per prompt; hint assembly is source-first"""
    from makeupdiffuse_b200 import B200ControlLDM
    cfg = {**SMALL, "vocab_size": 49408}  # the empty prompt's ids are BOS / EOS of the real vocabulary
    o = seeded_oracle(**cfg)
    m = B200ControlLDM(control_params=tiny_params, unet_params=tiny_params, dtype=torch.float32, device="cuda")
    enc = B200FrozenCLIPEmbedder(device="cuda", dtype=torch.float32, **cfg).load_state_dict(o.state_dict())
    m.attach_cond_stage_model(enc)
    uc = m.get_unconditional_conditioning(3)
    assert uc.shape == (3, 77, 64) and rel(uc, o(empty_prompt_tokens(3))) < 1e-4
    assert m.get_unconditional_conditioning(3) is uc  # encoded once
    tok = tokens(2, 77, 1000)
    assert rel(m.get_learned_conditioning(tok.cuda()), o(tok)) < 1e-4
    src, ref = torch.rand(2, 3, 8, 8), torch.rand(2, 3, 8, 8)
    h = B200ControlLDM.assemble_hint(src, ref)
    assert h.shape == (2, 6, 8, 8) and torch.equal(h[:, :3], src) and torch.equal(h[:, 3:], ref)
