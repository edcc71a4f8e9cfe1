"""Known-answer / structural tests that pin the oracle (SURVEY.md §8(c) K1-K7 + the parameter-count pins).

The reference ships no tests or golden vectors for this path ("parity unpinned"); these are the pins that are
mathematically forced by the cited reference code (diffmk/cddim.py, diffmk/makeup_diffuse.py, the yaml).
"""
import numpy as np
import pytest
import torch

from oracle import (ControlNet, ControlledUnetModel, MKDDIMSampler, OracleControlLDM, hash_uniform,
                    linear_beta_alphas_cumprod, seeded_state_dict, timestep_embedding)


def _cond(B, h, cdim, seed=0, hint=True):
    g = torch.Generator().manual_seed(seed)
    c = {"c_crossattn": [torch.randn(B, 77, cdim, generator=g)],
         "c_concat": [torch.rand(B, 6, 8 * h, 8 * h, generator=g)] if hint else None}
    return c, torch.randn(B, 4, h, h, generator=g)


@pytest.fixture(scope="module")
def tiny(tiny_params):
    m = OracleControlLDM(control_params=tiny_params, unet_params=tiny_params).eval()
    seeded_state_dict(m, 0)
    return m


def test_param_counts_and_keys_full_size():
    """yaml:52-84 must enumerate to the public SD-1.5 / ControlNet sizes; hint conv is [16,6,3,3] (train.py:60)."""
    with torch.device("meta"):
        cn, un = ControlNet(), ControlledUnetModel()
    assert sum(p.numel() for p in un.parameters()) == 859_520_964
    assert sum(p.numel() for p in cn.parameters()) == 361_279_552
    sd = cn.state_dict()
    assert tuple(sd["input_hint_block.0.weight"].shape) == (16, 6, 3, 3)
    assert len(un.state_dict()) == 686
    for k in ["time_embed.0.weight", "input_blocks.0.0.weight", "input_blocks.1.0.in_layers.2.weight",
              "input_blocks.1.0.emb_layers.1.weight", "input_blocks.1.0.out_layers.3.weight",
              "input_blocks.3.0.op.weight", "input_blocks.4.0.skip_connection.weight",
              "input_blocks.1.1.transformer_blocks.0.attn1.to_q.weight",
              "input_blocks.1.1.transformer_blocks.0.attn2.to_k.weight",
              "input_blocks.1.1.transformer_blocks.0.ff.net.0.proj.weight",
              "input_blocks.1.1.transformer_blocks.0.ff.net.2.weight", "input_blocks.1.1.proj_out.weight",
              "middle_block.1.norm.weight", "output_blocks.2.1.conv.weight", "output_blocks.5.2.conv.weight",
              "output_blocks.11.1.proj_in.weight", "out.0.weight", "out.2.weight"]:
        assert k in un.state_dict(), k
    for k in ["zero_convs.0.0.weight", "zero_convs.11.0.bias", "middle_block_out.0.weight",
              "input_hint_block.14.weight"]:
        assert k in sd, k
    assert tuple(un.state_dict()["output_blocks.5.0.in_layers.2.weight"].shape) == (1280, 1920, 3, 3)
    assert tuple(sd["input_hint_block.14.weight"].shape) == (320, 256, 3, 3)


def published_manifest(controlnet):
    """Every state-dict key and shape of the SD-1.5 latent UNet / the ControlNet as published,
    enumerated from the table there — NOT from the oracle's module tree, which is what it pins."""
    mc, ctx, emb = 320, 768, 1280
    m = {}

    def lin(k, o, i, bias=True):
        m[k + ".weight"] = (o, i)
        if bias:
            m[k + ".bias"] = (o,)

    def conv(k, o, i, r):
        m[k + ".weight"] = (o, i, r, r)
        m[k + ".bias"] = (o,)

    def norm(k, c):
        m[k + ".weight"] = (c,)
        m[k + ".bias"] = (c,)

    def res(k, ci, co):
        norm(k + ".in_layers.0", ci)
        conv(k + ".in_layers.2", co, ci, 3)
        lin(k + ".emb_layers.1", co, emb)
        norm(k + ".out_layers.0", co)
        conv(k + ".out_layers.3", co, co, 3)
        if ci != co:
            conv(k + ".skip_connection", co, ci, 1)

    def st(k, c):
        norm(k + ".norm", c)
        conv(k + ".proj_in", c, c, 1)
        t = k + ".transformer_blocks.0"
        for n in ("norm1", "norm2", "norm3"):
            norm(f"{t}.{n}", c)
        for a, kv in (("attn1", c), ("attn2", ctx)):
            lin(f"{t}.{a}.to_q", c, c, bias=False)
            lin(f"{t}.{a}.to_k", c, kv, bias=False)
            lin(f"{t}.{a}.to_v", c, kv, bias=False)
            lin(f"{t}.{a}.to_out.0", c, c)
        lin(f"{t}.ff.net.0.proj", 8 * c, c)
        lin(f"{t}.ff.net.2", c, 4 * c)
        conv(k + ".proj_out", c, c, 1)

    lin("time_embed.0", emb, mc)
    lin("time_embed.2", emb, emb)
    conv("input_blocks.0.0", mc, 4, 3)
    enc = [(1, 320, 320, True), (2, 320, 320, True), (3, None, 320, None), (4, 320, 640, True), (5, 640, 640, True),
           (6, None, 640, None), (7, 640, 1280, True), (8, 1280, 1280, True), (9, None, 1280, None), (10, 1280, 1280, False),
           (11, 1280, 1280, False)]
    skip = [320]
    for i, ci, co, attn in enc:
        if ci is None:
            conv(f"input_blocks.{i}.0.op", co, co, 3)
        else:
            res(f"input_blocks.{i}.0", ci, co)
            if attn:
                st(f"input_blocks.{i}.1", co)
        skip.append(co)
    res("middle_block.0", 1280, 1280)
    st("middle_block.1", 1280)
    res("middle_block.2", 1280, 1280)
    if controlnet:
        ch = [6, 16, 16, 32, 32, 96, 96, 256, 320]
        for j in range(8):
            conv(f"input_hint_block.{2 * j}", ch[j + 1], ch[j], 3)
        for j, c in enumerate(skip):
            conv(f"zero_convs.{j}.0", c, c, 1)
        conv("middle_block_out.0", 1280, 1280, 1)
        assert skip == [320, 320, 320, 320, 640, 640, 640, 1280, 1280, 1280, 1280, 1280]
        return m
    dec = [(0, 1280, False, False), (1, 1280, False, False), (2, 1280, False, True), (3, 1280, True, False),
           (4, 1280, True, False), (5, 1280, True, True), (6, 640, True, False), (7, 640, True, False), (8, 640, True, True),
           (9, 320, True, False), (10, 320, True, False), (11, 320, True, False)]
    h, widths = 1280, []
    for i, co, attn, up in dec:
        ci = h + skip.pop()
        widths.append(ci)
        res(f"output_blocks.{i}.0", ci, co)
        if attn:
            st(f"output_blocks.{i}.1", co)
        if up:
            conv(f"output_blocks.{i}.{2 if attn else 1}.conv", co, co, 3)
        h = co
    assert widths == [2560, 2560, 2560, 2560, 2560, 1920, 1920, 1280, 960, 960, 640, 640]  # Appendix A, last line
    norm("out.0", 320)
    conv("out.2", 4, 320, 3)
    return m


def test_full_manifest_of_both_networks_matches_the_published_checkpoints():
    """every key and every shape of the oracle's ControlNet and UNet state dicts == the published manifest (SURVEY.md
    Appendix A): the structural pin of rows A7 / A8, complete rather than sampled"""
    with torch.device("meta"):
        cn, un = ControlNet(), ControlledUnetModel()
    for net, want in ((cn, published_manifest(True)), (un, published_manifest(False))):
        got = {k: tuple(v.shape) for k, v in net.state_dict().items()}
        assert sorted(got) == sorted(want), (sorted(set(got) ^ set(want))[:8])
        assert got == want, [k for k in got if got[k] != want[k]][:8]
    assert len(published_manifest(False)) == 686
    assert sum(int(np.prod(v)) for v in published_manifest(False).values()) == 859_520_964
    assert sum(int(np.prod(v)) for v in published_manifest(True).values()) == 361_279_552


def test_K1_schedule():
    ac = linear_beta_alphas_cumprod().astype(np.float32)
    np.testing.assert_allclose(ac[[0, 1, 981, 999]], [0.999149978, 0.998296022, 0.005775500, 0.004660098],
                               rtol=2e-7)
    m = OracleControlLDM(control_params=dict(model_channels=32, num_heads=2, context_dim=16),
                         unet_params=dict(model_channels=32, num_heads=2, context_dim=16))
    s = MKDDIMSampler(m)
    s.make_schedule(50, ddim_eta=0.0, verbose=False)
    assert list(s.ddim_timesteps[:3]) == [1, 21, 41] and s.ddim_timesteps[-1] == 981 and len(s.ddim_timesteps) == 50
    s.make_schedule(20, ddim_eta=0.0, verbose=False)
    assert list(s.ddim_timesteps[:2]) == [1, 51] and s.ddim_timesteps[-1] == 951
    assert float(s.ddim_alphas_prev[0]) == float(m.alphas_cumprod[0])
    assert float(s.ddim_sigmas.abs().max()) == 0.0


def test_timestep_embedding_layout():
    e = timestep_embedding(torch.tensor([0, 7]), 8)
    assert torch.allclose(e[0], torch.tensor([1., 1, 1, 1, 0, 0, 0, 0]))
    f = torch.exp(-np.log(10000.0) * torch.arange(4) / 4)
    assert torch.allclose(e[1], torch.cat([torch.cos(7 * f), torch.sin(7 * f)]), atol=1e-6)


@pytest.mark.parametrize("S,gain", [(50, 13.152870), (20, 11.068869)])
def test_K2_eps_zero_closed_form(tiny, S, gain, monkeypatch):
    """eps == 0, eta == 0  =>  x_0 = x_T * sqrt(abar_prev[0] / abar[last])  (cddim.py:63,74,78,94)."""
    cond, x = _cond(2, 8, 64)
    monkeypatch.setattr(tiny, "apply_model", lambda x, t, c: torch.zeros_like(x))
    out, _ = MKDDIMSampler(tiny).sample(S, 2, (4, 8, 8), cond, eta=0.0, x_T=x, verbose=False)
    torch.testing.assert_close(out, x * gain, rtol=2e-5, atol=0)


def test_K3_cfg_identities(tiny):
    cond, x = _cond(2, 8, 64)
    uc, _ = _cond(2, 8, 64, seed=1)
    uc["c_concat"] = cond["c_concat"]
    s = MKDDIMSampler(tiny)
    s.make_schedule(50, verbose=False)
    t = torch.full((2,), 981, dtype=torch.long)
    calls = []
    orig = tiny.apply_model
    tiny.apply_model = lambda x, t, c: (calls.append(x.shape[0]), orig(x, t, c))[1]
    try:
        with torch.no_grad():
            torch.manual_seed(0)
            a, _ = s.denoising_step(x, cond, t, 49, unconditional_guidance_scale=1.0, unconditional_conditioning=uc)
            b, _ = s.denoising_step(x, cond, t, 49, unconditional_guidance_scale=9.0, unconditional_conditioning=None)
            assert calls == [2, 2]  # single un-doubled call (cddim.py:15-16)
            torch.testing.assert_close(a, b)
            c9, _ = s.denoising_step(x, cond, t, 49, unconditional_guidance_scale=9.0, unconditional_conditioning=cond)
            assert calls[-1] == 4
            torch.testing.assert_close(c9, a, rtol=1e-4, atol=1e-4)  # uc == c => independent of scale
            # batching order is [uncond; cond]
            e_c, e_u = orig(x, t, cond), orig(x, t, uc)
            g, _ = s.denoising_step(x, cond, t, 49, unconditional_guidance_scale=3.0, unconditional_conditioning=uc)
            e = e_u + 3.0 * (e_c - e_u)
            a_t, a_p, s1 = float(s.ddim_alphas[49]), float(s.ddim_alphas_prev[49]), float(s.ddim_sqrt_one_minus_alphas[49])
            ref = np.sqrt(a_p) * (x - s1 * e) / np.sqrt(a_t) + np.sqrt(1 - a_p) * e
            torch.testing.assert_close(g, ref, rtol=1e-4, atol=1e-4)
    finally:
        tiny.apply_model = orig


def test_K4_zero_convs_make_hint_irrelevant(tiny_params):
    m = OracleControlLDM(control_params=tiny_params, unet_params=tiny_params).eval()
    seeded_state_dict(m, 0)
    with torch.no_grad():
        for zc in list(m.control_model.zero_convs) + [m.control_model.middle_block_out]:
            for p in zc.parameters():
                p.zero_()
        cond, x = _cond(1, 8, 64)
        t = torch.tensor([501])
        a = m.apply_model(x, t, cond)
        b = m.apply_model(x, t, {"c_crossattn": cond["c_crossattn"], "c_concat": None})
    torch.testing.assert_close(a, b, rtol=0, atol=0)


def test_K5_eta0_independent_of_rng_but_consumes_it(tiny):
    cond, x = _cond(1, 8, 64)
    s = MKDDIMSampler(tiny)
    torch.manual_seed(1)
    a, _ = s.sample(4, 1, (4, 8, 8), cond, eta=0.0, x_T=x, verbose=False)
    after_a = torch.rand(1)
    torch.manual_seed(2)
    b, _ = s.sample(4, 1, (4, 8, 8), cond, eta=0.0, x_T=x, verbose=False)
    torch.testing.assert_close(a, b, rtol=0, atol=0)
    torch.manual_seed(1)
    for _ in range(4):
        torch.randn(1, 4, 8, 8)
    assert torch.equal(after_a, torch.rand(1))  # one randn(x.shape) per step (cddim.py:75)


def test_K6_reconstruct_equals_sample(tiny):
    cond, x = _cond(2, 8, 64)
    s = MKDDIMSampler(tiny)
    a, _ = s.sample(10, 2, (4, 8, 8), cond, eta=0.0, x_T=x, verbose=False)
    with torch.no_grad():
        b = s.reconstruct(x, cond, t_start=10)
        c = s.reconstruct(x, cond, t_start=4)  # truncated: only the 4 smallest timesteps
    torch.testing.assert_close(a, b, rtol=0, atol=0)
    assert not torch.allclose(a, c)


def test_K7_control_scales_identity_and_dataflow(tiny):
    """makeup_diffuse.py:157-168: ControlNet -> x scale -> UNet; mid residual is consumed first (pop())."""
    cond, x = _cond(1, 8, 64)
    t = torch.tensor([301])
    with torch.no_grad():
        ctrl = tiny.control_model(x=x, hint=cond["c_concat"][0], timesteps=t, context=cond["c_crossattn"][0])
        assert len(ctrl) == 13
        shapes = [(c.shape[1], c.shape[2]) for c in ctrl]
        assert shapes == [(64, 8)] * 3 + [(64, 4)] + [(128, 4)] * 2 + [(128, 2)] + [(256, 2)] * 2 + [(256, 1)] * 4
        e = tiny.model.diffusion_model(x=x, timesteps=t, context=cond["c_crossattn"][0], control=ctrl)
        torch.testing.assert_close(e, tiny.apply_model(x, t, cond), rtol=0, atol=0)
        eps, x0 = tiny.apply_model(x, t, cond, return_all=True)
        a = float(tiny.alphas_cumprod[301])
        torch.testing.assert_close(x0, (x - np.sqrt(1 - a) * eps) / np.sqrt(a), rtol=1e-4, atol=1e-4)


def test_hash_init_is_device_independent_and_nonzero(tiny):
    u = hash_uniform((1000,), 0, 123)
    assert u.min() >= -1 and u.max() < 1 and abs(float(u.mean())) < 0.1 and 0.5 < float(u.std()) < 0.65
    assert torch.equal(u, hash_uniform((1000,), 0, 123))
    assert not torch.equal(u, hash_uniform((1000,), 1, 123))
    assert float(tiny.control_model.zero_convs[3][0].weight.abs().max()) > 0
    assert float(tiny.model.diffusion_model.out[2].weight.abs().max()) > 0
