"""Harness I/O: pair list, log_results, save_local — diffusion_makeup.py:326-411,
datasets.py:728-784.  CPU: file formats and the host logic of save_local over the reference's own torchvision / numpy
sequence (tests/fake_ops.py).  GPU: the grid kernel byte-for-byte against that sequence, and log_results end to end
(x_p -> get_z -> q_sample -> one-step x_0 / DDIM / guided DDIM -> decode) against the oracle pipeline."""
import os

import numpy as np
import pytest
import torch

import fake_ops
from makeupdiffuse_b200 import harness, ops


def test_pair_list_and_test_pairs_files(tmp_path):
    p = tmp_path / "test_0412.txt"
    p.write_text("non-makeup/xfsy_0106.png makeup/vFG112.png\nnon-makeup/vSYYZ639.png makeup/XMY-074.png\n\n")
    src, ref = harness.read_pair_list(str(p))
    assert src == ["non-makeup/xfsy_0106.png", "non-makeup/vSYYZ639.png"] and ref == ["makeup/vFG112.png", "makeup/XMY-074.png"]
    names = harness.pair_basenames(src, ref)
    assert names == ["xfsy_0106&vFG112", "vSYYZ639&XMY-074"]  # datasets.py:760-764
    rows = harness.test_pair_rows(3, names)
    assert rows[1] == ["0003-2", "non-makeup/vSYYZ639.png", "makeup/XMY-074.png"]  # diffusion_makeup.py:376-381
    out = tmp_path / "pairs.txt"
    harness.write_test_pairs(str(out), rows)
    assert out.read_text() == "0003-1 non-makeup/xfsy_0106.png makeup/vFG112.png\n0003-2 non-makeup/vSYYZ639.png makeup/XMY-074.png\n"


def test_save_local_writes_reference_grids(tmp_path, monkeypatch):
    monkeypatch.setattr(ops, "image_grid_u8", fake_ops.image_grid_u8)
    from PIL import Image
    g = torch.Generator().manual_seed(0)
    images = {"samples": torch.rand(3, 3, 16, 16, generator=g) * 2.4 - 1.2, "control_src": torch.rand(3, 3, 16, 16, generator=g) * 2 - 1}
    grids = harness.save_local(images, 7, str(tmp_path / "run"))
    for k in images:
        f = tmp_path / "run" / f"{k}_0007.png"
        assert f.exists()
        assert np.array_equal(np.asarray(Image.open(f)), grids[k])
        assert grids[k].shape == (2 * 18 + 2, 2 * 18 + 2, 3)  # nrow = number of panels (2): 3 images -> 2 x 2 cells, padding 2
        assert grids[k][0, 0, 0] == 127  # pad_value 0 -> (0 + 1) / 2 * 255 truncated


def test_conditioning_text_panels():
    """log["conditioning"] (diffusion_makeup.py:372; upstream ldm.util.log_txt_as_img): white canvases with the image names in
    black from the top-left corner, wrapped every int(40 * width / 256) characters, [B, 3, H, W] in [-1, 1]"""
    names = ["xfsy_0106&vFG112", "", "a" * 95]
    t = harness.txt_panels((256, 256), names, size=16)
    assert t.shape == (3, 3, 256, 256) and t.dtype == torch.float32
    assert float(t.max()) == 1.0 and float(t.min()) == -1.0
    assert bool((t[1] == 1.0).all())                               # empty caption: blank white canvas
    ink = lambda x: (x[0] < 0.0).nonzero()                          # noqa: E731  (dark pixels of the first channel: [row, col])
    one, three = ink(t[0]), ink(t[2])
    assert 0 < one.shape[0] and int(one[:, 0].max()) < 24           # one line of 16-px text at the top
    assert int(three[:, 0].max()) > 2 * int(one[:, 0].max())        # 95 characters wrap into 3 lines of 40
    assert int(three[:, 0].max()) < 80 and int(one[:, 1].min()) < 8  # starts at the left edge
    w = harness.txt_panels((128, 64), ["b" * 45], size=10)          # other canvas: wrap width scales with the canvas width (20)
    assert w.shape == (1, 3, 64, 128) and int(ink(w[0])[:, 0].max()) > 20


def reference_grid(x, nrow, padding, clamp, rescale):
    return fake_ops.image_grid_u8(x, nrow, padding, clamp, rescale).numpy()


@pytest.mark.gpu
@pytest.mark.parametrize("N,C,H,W,nrow,pad", [(5, 3, 32, 24, 3, 2), (1, 3, 8, 8, 4, 2), (7, 1, 16, 16, 7, 0), (6, 3, 256, 256, 6, 2), (4, 3, 5, 3, 2, 1)])
def test_image_grid_kernel_is_byte_identical(N, C, H, W, nrow, pad):
    g = torch.Generator().manual_seed(N * 100 + H)
    x = torch.rand(N, C, H, W, generator=g) * 2.6 - 1.3  # beyond [-1, 1]: exercises the clamp
    x.view(-1)[:4] = torch.tensor([1.0, -1.0, 0.0, 0.999999])
    for clamp, rescale in ((True, True), (True, False)):
        xx = x if rescale else x.abs().clamp(max=1.0)  # without rescale only [0, 1] values are meaningful as uint8
        got = ops.image_grid_u8(xx.cuda(), nrow, pad, clamp, rescale).cpu().numpy()
        ref = reference_grid(xx, nrow, pad, clamp, rescale)
        assert got.shape == ref.shape and np.array_equal(got, ref)


@pytest.mark.gpu
def test_log_results_end_to_end_against_oracle(tmp_path):
    from makeupdiffuse_b200 import (B200ControlLDM, B200DDIMSampler, B200FirstStageDecoder, B200FirstStageEncoder,
                                    B200FrozenCLIPEmbedder)
    from oracle import MKDDIMSampler, OracleControlLDM, seeded_state_dict
    from oracle.clip import empty_prompt_tokens
    from oracle.vae import OracleFirstStageDecoder, OracleFirstStageEncoder, decode_first_stage, get_z
    from test_clip import SMALL, seeded_oracle, tokens
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    params = dict(model_channels=64, num_heads=4, context_dim=64)
    dd = dict(ch=64, ch_mult=(1, 2, 4, 4), num_res_blocks=1)
    ccfg = {**SMALL, "vocab_size": 49408}
    with torch.device("cuda"):
        ol = OracleControlLDM(control_params=params, unet_params=params).eval()
        oe, od = OracleFirstStageEncoder(ddconfig=dd).eval(), OracleFirstStageDecoder(ddconfig=dd).eval()
    oc = seeded_oracle(**ccfg).cuda()
    sd = seeded_state_dict(ol, 0)
    sde, sdd = seeded_state_dict(oe, 0, prefix="first_stage_model."), seeded_state_dict(od, 0, prefix="first_stage_model.")
    m = B200ControlLDM(params, params, dtype=torch.float32).load_state_dict(sd)
    m.attach_first_stage_encoder(B200FirstStageEncoder(ddconfig=dd, dtype=torch.float32).load_state_dict(sde))
    m.attach_first_stage_decoder(B200FirstStageDecoder(ddconfig=dd, dtype=torch.float32).load_state_dict(sdd))
    m.attach_cond_stage_model(B200FrozenCLIPEmbedder(device="cuda", dtype=torch.float32, **ccfg).load_state_dict(oc.state_dict()))
    B, hw, S, scale, t_min = 2, 64, 6, 9.0, 400
    g = torch.Generator().manual_seed(5)
    tok = tokens(B, 77, 49408)
    batch = {"pgt_sr": torch.rand(B, 3, hw, hw, generator=g) * 2 - 1, "src_img": torch.rand(B, 3, hw, hw, generator=g),
             "ref_img": torch.rand(B, 3, hw, hw, generator=g), "tokens": tok.cuda(), "img_name": ["a&b", "c&d"]}
    pairs = []
    gg = torch.Generator(device="cuda").manual_seed(11)
    log = harness.log_results(m, batch, 2, ddim_steps=S, unconditional_guidance_scale=scale, t_min=t_min, test_pairs=pairs,
                              sampler=B200DDIMSampler(m, use_cuda_graph=False), generator=gg)
    assert pairs == [["0002-1", "non-makeup/a.png", "makeup/b.png"], ["0002-2", "non-makeup/c.png", "makeup/d.png"]]
    assert list(log) == ["reconstruction", "control_src", "control_ref", "conditioning", "ground_truth", "sample_ddmp", "samples",
                         "samples_cfg_scale_9.00"]  # the reference's panels in the reference's order (diffusion_makeup.py:369-408)
    assert log["conditioning"].shape == (B, 3, 256, 256) and float(log["conditioning"].min()) == -1.0
    # the same flow on the oracle, same random draws in the same order (zn, t, noise; x_T of each DDIM run)
    gg = torch.Generator(device="cuda").manual_seed(11)
    dev = "cuda"
    pgt = batch["pgt_sr"].to(dev)
    c_cat = torch.cat((batch["src_img"], batch["ref_img"]), 1).to(dev)
    with torch.no_grad():
        c = oc(tok.cuda())
        zn = torch.randn(B, 4, hw // 8, hw // 8, device=dev, generator=gg)
        z = get_z(oe, pgt, zn)
        ref = {"reconstruction": decode_first_stage(od, z), "control_src": c_cat[:, :3] * 2 - 1, "control_ref": c_cat[:, 3:] * 2 - 1,
               "ground_truth": pgt}
        t = torch.randint(t_min, 1000, (B,), device=dev, generator=gg).long()
        noise = torch.randn(z.shape, device=dev, generator=gg)
        cond = {"c_concat": [c_cat], "c_crossattn": [c]}
        x_noisy = ol.q_sample(z, t, noise)
        eps = ol.apply_model(x_noisy, t, cond)
        ref["sample_ddmp"] = decode_first_stage(od, ol.predict_start_from_noise(x_noisy, t, eps))
    assert int(t.min()) >= t_min
    for k in ("reconstruction", "control_src", "control_ref", "ground_truth", "sample_ddmp"):
        e = float((log[k] - ref[k]).norm() / ref[k].norm())
        print(f"log_results[{k}]: rel-L2 {e:.2e}")
        assert e < 1e-3, (k, e)
    # the two DDIM panels draw x_T from torch's global CUDA generator inside sample(): replay with the same seed
    so, sb = MKDDIMSampler(ol), B200DDIMSampler(m, use_cuda_graph=False)
    with torch.no_grad():
        uc = {"c_concat": [c_cat], "c_crossattn": [oc(empty_prompt_tokens(B).cuda())]}
        for kw, key in (({}, "samples"), (dict(unconditional_guidance_scale=scale, unconditional_conditioning=uc), "samples_cfg")):
            x_T = torch.randn(B, 4, hw // 8, hw // 8, device=dev, generator=gg)
            xo, _ = so.sample(S, B, (4, hw // 8, hw // 8), cond, eta=0.0, x_T=x_T, verbose=False, **kw)
            kb = dict(kw)
            if kb:
                kb["unconditional_conditioning"] = {"c_concat": [c_cat], "c_crossattn": [m.get_unconditional_conditioning(B)]}
            xb, _ = sb.sample(S, B, (4, hw // 8, hw // 8), {"c_concat": [c_cat], "c_crossattn": [m.get_learned_conditioning(tok.cuda())]},
                              eta=0.0, x_T=x_T, verbose=False, **kb)
            e = float((m.decode_first_stage(xb) - decode_first_stage(od, xo)).norm() / decode_first_stage(od, xo).norm())
            print(f"harness DDIM panel {key}: rel-L2 {e:.2e}")
            assert e < 1e-3
    assert log["samples"].shape == (B, 3, hw, hw) and torch.isfinite(log["samples_cfg_scale_9.00"]).all()
    grids = harness.save_local(log, 2, str(tmp_path))
    assert len(grids) == 8 and os.path.exists(tmp_path / "samples_0002.png") and os.path.exists(tmp_path / "conditioning_0002.png")
    for k, v in log.items():
        assert np.array_equal(grids[k], reference_grid(v.float().cpu(), 8, 2, True, True))
