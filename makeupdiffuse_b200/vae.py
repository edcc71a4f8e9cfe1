"""``B200FirstStageDecoder``: the VAE decoder the reference runs right after sampling.

Reference surface: ``decode_first_stage(z)`` = ``first_stage_model.decode(1 / scale_factor * z)``
(``diffmk/makeups.py:260-262``; ``diffmk/diffusion_makeup.py:389,396,409``), ``first_stage_model`` = upstream
``AutoencoderKL`` with the ``ddconfig`` of ``diffmodels/base_diffusion_makeup.yaml:86-105``.  Same constructor kwargs
(``embed_dim``, ``ddconfig``), ``load_state_dict`` takes the upstream ``first_stage_model.*`` keys (only
``post_quant_conv`` and ``decoder.*`` are consumed), ``decode(z)`` returns NCHW fp32 images like the reference.

Everything runs through the same C-ABI kernels as the denoiser: 3x3 / 1x1 convolutions on the tcgen05 implicit-GEMM
kernel (the 4 latent channels are zero-padded to 64 for ``conv_in``; nearest x2 upsampling is materialised by the
library ahead of the conv), GroupNorm(+swish) as one streaming pass from statistics fused into the producing epilogue
(up to 512 row tiles per sample, i.e. 256^2 maps; the two-phase kernel above that), the residual trunk in fp32 (``nets.Act``).  The single-head
512-wide attention of ``mid.attn_1`` exceeds the tensor-core attention kernel's head-dim range and runs on the SIMT
kernel.  No PyTorch compute fallback.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib as L
from . import ops
from .nets import Act

_MAX_STAT_TILES = 512  # groupnorm_apply sums a sample's tile partials in its prologue (512 tiles = a 256^2 map)


class _FirstStageNet(nn.Module):
    """buffers, weight slots and the layer kernels shared by the decoder and the encoder"""

    def __init__(self, dtype):
        super().__init__()
        self.dtype = dtype
        self.w: dict[str, torch.Tensor] = {}
        self._bufs: dict = {}
        self._loaded = False

    @property
    def _hi(self):
        return self.dtype != torch.float32

    def _put(self, name, t, act_dtype=False):
        self.w[name] = t.detach().to(device=self._device, dtype=self.dtype if act_dtype else torch.float32).contiguous()

    @staticmethod
    def _krsc(w):
        return w.permute(0, 2, 3, 1).contiguous()

    def _put_res(self, g, key, cin, cout):
        for n in ("norm1", "norm2"):
            self._put(f"{key}.{n}.g", g(f"{key}.{n}.weight")); self._put(f"{key}.{n}.b", g(f"{key}.{n}.bias"))
        for n in ("conv1", "conv2") + (("nin_shortcut",) if cin != cout else ()):
            self._put(f"{key}.{n}.w", self._krsc(g(f"{key}.{n}.weight")), True); self._put(f"{key}.{n}.b", g(f"{key}.{n}.bias"))

    def _put_attn(self, g, key):
        self._put(key + ".norm.g", g(key + ".norm.weight")); self._put(key + ".norm.b", g(key + ".norm.bias"))
        self._put(key + ".qkv.w", torch.cat([self._krsc(g(key + f".{n}.weight")) for n in "qkv"], 0), True)
        self._put(key + ".qkv.b", torch.cat([g(key + f".{n}.bias") for n in "qkv"], 0))
        self._put(key + ".po.w", self._krsc(g(key + ".proj_out.weight")), True); self._put(key + ".po.b", g(key + ".proj_out.bias"))

    @staticmethod
    def _res_shapes(out, key, ci, co):
        out.update({key + ".norm1.weight": (ci,), key + ".norm1.bias": (ci,), key + ".conv1.weight": (co, ci, 3, 3),
                    key + ".conv1.bias": (co,), key + ".norm2.weight": (co,), key + ".norm2.bias": (co,),
                    key + ".conv2.weight": (co, co, 3, 3), key + ".conv2.bias": (co,)})
        if ci != co:
            out.update({key + ".nin_shortcut.weight": (co, ci, 1, 1), key + ".nin_shortcut.bias": (co,)})

    @staticmethod
    def _attn_shapes(out, key, c):
        out.update({key + ".norm.weight": (c,), key + ".norm.bias": (c,)})
        for n in ("q", "k", "v", "proj_out"):
            out.update({key + f".{n}.weight": (c, c, 1, 1), key + f".{n}.bias": (c,)})

    def _check_keys(self, sd, prefix, strict, what):
        shapes = self.upstream_shapes()
        missing = [k for k in shapes if prefix + k not in sd]
        if missing and strict:
            raise KeyError(f"missing first-stage {what} keys: {missing[:5]} ...")
        for k, shp in shapes.items():
            if tuple(sd[prefix + k].shape) != tuple(shp):
                raise ValueError(f"{prefix + k}: shape {tuple(sd[prefix + k].shape)} != expected {shp}")


    # ---- buffers ---------------------------------------------------------------------------------------------------
    def _buf(self, name, rows, cols, dtype=None, zero=False):
        key = (name, rows, cols, dtype)
        b = self._bufs.get(key)
        if b is None:
            b = (torch.zeros if zero else torch.empty)(rows, cols, dtype=dtype or self.dtype, device=self._device)
            self._bufs[key] = b
        return b

    def _stats_ok(self, HW):
        return self._hi and HW % 128 == 0 and HW // 128 <= _MAX_STAT_TILES

    def _act(self, name, rows, cols, HW, lo, hi):
        """trunk tensor with the requested forms (fp32 check mode: one fp32 buffer) and, where possible, statistics"""
        if not self._hi:
            return Act(self._buf(name, rows, cols))
        a = Act(self._buf(name, rows, cols) if lo else None, self._buf(name + "32", rows, cols, torch.float32) if hi else None)
        if hi and self._stats_ok(HW):
            a.st = self._buf("st_" + name, rows // 128, cols * 2, torch.float32).view(rows // 128, cols, 2)
        return a

    def _gn(self, x, y, N, g, b, silu):
        if x.st is not None:
            ops.groupnorm_apply(x.src(), y, N, g, b, 1e-6, silu, x.st)
        else:
            ws = self._buf("gn_ws", 1, ops.groupnorm_workspace_bytes(N) // 4, torch.float32)
            ops.groupnorm(x.src(), y, N, g, b, 1e-6, silu, ws)

    def _conv(self, x, key, y, N, H, W, R, **kw):
        ops.conv2d(x, self.w[key + ".w"], y, N=N, H=H, W=W, R=R, S=R, pad=R // 2, bias=self.w[key + ".b"],
                   workspace=self._ws, **kw)

    @staticmethod
    def _needs(layer):
        """(lo, hi): forms a layer reads of its input"""
        if layer is None:           # norm_out
            return False, True
        if layer[0] == "res":
            return layer[2] != layer[3], True
        if layer[0] == "attn":
            return False, True
        return True, False          # upsample / downsample conv: tensor-core operand only

    # ---- layers -----------------------------------------------------------------------------------------------------
    def _res(self, layer, x, y, N, H, W):
        _, key, cin, cout = layer
        M = N * H * W
        t1 = self._buf("gn_a", M, cin)
        self._gn(x, t1, N, self.w[key + ".norm1.g"], self.w[key + ".norm1.b"], True)
        h = self._act("res_h", M, cout, H * W, lo=False, hi=True) if self._hi else Act(self._buf("res_h", M, cout))
        self._conv(t1, key + ".conv1", h.lo, N, H, W, 3, y32=h.hi, stats=h.st)
        t2 = self._buf("gn_b", M, cout)
        self._gn(h, t2, N, self.w[key + ".norm2.g"], self.w[key + ".norm2.b"], True)
        if cin != cout:
            s = self._act("res_sk", M, cout, 1, lo=False, hi=True) if self._hi else Act(self._buf("res_sk", M, cout))
            self._conv(x.lo, key + ".nin_shortcut", s.lo, N, H, W, 1, y32=s.hi)
            sk = s.src()
        else:
            sk = x.src()
        self._conv(t2, key + ".conv2", y.lo, N, H, W, 3, residual=sk, y32=y.hi, stats=y.st)

    def _attn(self, layer, x, y, N, H, W):
        _, key, c = layer
        M = N * H * W
        n = self._buf("at_n", M, c)
        self._gn(x, n, N, self.w[key + ".norm.g"], self.w[key + ".norm.b"], False)
        HW = H * W
        att = self._buf("at_o", M, c)
        if self._hi and c % 64 == 0 and HW % 64 == 0 and HW <= 2048:
            # One 512-wide head is outside the flash kernel's head-dim range, but per image it is just two GEMMs around
            # a row softmax, all on the tensor-core kernel with the projections' outputs used in place as operands:
            #   S = Q K^T      x = q_b [HW, C],  "weights" = k_b [HW, C]             -> fp32 [HW, HW]
            #   P = softmax(S / sqrt C)                                               -> bf16 [HW, HW]
            #   O = P V + b_v  x = P,            "weights" = V^T_b = W_v n_b^T [C, HW] (the v projection with the roles of
            #                                    activation and weight swapped, so no transpose pass exists; the v bias
            #                                    is added after P V, exact because the rows of P sum to 1)
            q, k = self._buf("at_q", M, c), self._buf("at_k", M, c)
            wq, wk, wv = (self.w[key + ".qkv.w"][i * c:(i + 1) * c] for i in range(3))
            bq, bk, bv = (self.w[key + ".qkv.b"][i * c:(i + 1) * c] for i in range(3))
            ops.conv2d(n, wq, q, N=1, H=1, W=M, bias=bq, workspace=self._ws)
            ops.conv2d(n, wk, k, N=1, H=1, W=M, bias=bk, workspace=self._ws)
            vt = self._buf("at_vt", c, HW)
            s32 = self._buf("at_s", HW, HW, torch.float32)
            p = self._buf("at_p", HW, HW)
            wv2d = wv.reshape(c, c)
            for b in range(N):
                rows = slice(b * HW, (b + 1) * HW)
                ops.conv2d(wv2d, n[rows], vt, N=1, H=1, W=c, workspace=self._ws)
                ops.conv2d(q[rows], k[rows], None, N=1, H=1, W=HW, y32=s32, workspace=self._ws)
                ops.softmax_rows(s32, p, scale=float(c) ** -0.5)
                ops.conv2d(p, vt, att[rows], N=1, H=1, W=HW, bias=bv, workspace=self._ws)
        else:
            qkv = self._buf("at_qkv", M, 3 * c)
            self._conv(n, key + ".qkv", qkv, N, H, W, 1)
            ops.attention(qkv[:, :c], qkv[:, c:2 * c], qkv[:, 2 * c:], att, B=N, heads=1, Nq=HW, Nkv=HW, d=c,
                          scale=float(c) ** -0.5)
        self._conv(att, key + ".po", y.lo, N, H, W, 1, residual=x.src(), y32=y.hi, stats=y.st)


class B200FirstStageDecoder(_FirstStageNet):
    def __init__(self, embed_dim=4, ddconfig=None, dtype=torch.bfloat16, **unused):
        super().__init__(dtype)
        dd = dict(ch=128, out_ch=3, ch_mult=(1, 2, 4, 4), num_res_blocks=2, z_channels=4)
        dd.update(ddconfig or {})
        if list(dd.get("attn_resolutions", [])):
            raise NotImplementedError("attn_resolutions other than [] (the yaml's value) are not implemented")
        self.embed_dim, self.dd = embed_dim, dd
        ch, mult, nrb = dd["ch"], tuple(dd["ch_mult"]), dd["num_res_blocks"]
        self.zc, self.out_ch = dd["z_channels"], dd["out_ch"]
        cin = ch * mult[-1]
        self.block_in = cin
        # program: ("res", key, cin, cout) | ("attn", key, c) | ("up", key, c), in execution order
        self.layers = [("res", "decoder.mid.block_1", cin, cin), ("attn", "decoder.mid.attn_1", cin),
                       ("res", "decoder.mid.block_2", cin, cin)]
        for lvl in reversed(range(len(mult))):
            cout = ch * mult[lvl]
            for i in range(nrb + 1):
                self.layers.append(("res", f"decoder.up.{lvl}.block.{i}", cin, cout))
                cin = cout
            if lvl != 0:
                self.layers.append(("up", f"decoder.up.{lvl}.upsample.conv", cin))
        self.c_last = cin

    # ---- parameters ----------------------------------------------------------------------------------------------
    def upstream_shapes(self) -> dict:
        out = {"post_quant_conv.weight": (self.zc, self.embed_dim, 1, 1), "post_quant_conv.bias": (self.zc,),
               "decoder.conv_in.weight": (self.block_in, self.zc, 3, 3), "decoder.conv_in.bias": (self.block_in,),
               "decoder.norm_out.weight": (self.c_last,), "decoder.norm_out.bias": (self.c_last,),
               "decoder.conv_out.weight": (self.out_ch, self.c_last, 3, 3), "decoder.conv_out.bias": (self.out_ch,)}
        for layer in self.layers:
            kind, key = layer[0], layer[1]
            if kind == "res":
                ci, co = layer[2], layer[3]
                out.update({key + ".norm1.weight": (ci,), key + ".norm1.bias": (ci,), key + ".conv1.weight": (co, ci, 3, 3),
                            key + ".conv1.bias": (co,), key + ".norm2.weight": (co,), key + ".norm2.bias": (co,),
                            key + ".conv2.weight": (co, co, 3, 3), key + ".conv2.bias": (co,)})
                if ci != co:
                    out.update({key + ".nin_shortcut.weight": (co, ci, 1, 1), key + ".nin_shortcut.bias": (co,)})
            elif kind == "attn":
                c = layer[2]
                out.update({key + ".norm.weight": (c,), key + ".norm.bias": (c,)})
                for n in ("q", "k", "v", "proj_out"):
                    out.update({key + f".{n}.weight": (c, c, 1, 1), key + f".{n}.bias": (c,)})
            else:
                c = layer[2]
                out.update({key + ".weight": (c, c, 3, 3), key + ".bias": (c,)})
        return out

    def load_state_dict(self, sd, strict=True, prefix="first_stage_model.", device="cuda"):  # noqa: D401
        """Repack the upstream-keyed decoder weights (fp32 OIHW) into KRSC bf16 / fp32 vectors.  Encoder, quant_conv and
        loss keys of a full ``first_stage_model`` state dict are ignored; with strict=True every decoder key must be there."""
        self._device = torch.device(device)
        krsc = lambda w: w.permute(0, 2, 3, 1).contiguous()  # noqa: E731
        shapes = self.upstream_shapes()
        missing = [k for k in shapes if prefix + k not in sd]
        if missing and strict:
            raise KeyError(f"missing first-stage decoder keys: {missing[:5]} ...")
        g = lambda k: sd[prefix + k]  # noqa: E731
        for k, shp in shapes.items():
            if tuple(g(k).shape) != tuple(shp):
                raise ValueError(f"{prefix + k}: shape {tuple(g(k).shape)} != expected {shp}")
        self._put("pq.w", krsc(g("post_quant_conv.weight")), True); self._put("pq.b", g("post_quant_conv.bias"))
        w_in = krsc(g("decoder.conv_in.weight"))
        if self._hi:  # 4 latent channels -> 64 zero-padded input channels: conv_in runs on the tensor cores
            wp = torch.zeros(w_in.shape[0], 3, 3, 64, dtype=w_in.dtype, device=w_in.device)
            wp[..., :self.zc] = w_in
            w_in = wp
        self._put("conv_in.w", w_in, True); self._put("conv_in.b", g("decoder.conv_in.bias"))
        for layer in self.layers:
            kind, key = layer[0], layer[1]
            if kind == "res":
                for n in ("norm1", "norm2"):
                    self._put(f"{key}.{n}.g", g(f"{key}.{n}.weight")); self._put(f"{key}.{n}.b", g(f"{key}.{n}.bias"))
                for n in ("conv1", "conv2") + (("nin_shortcut",) if layer[2] != layer[3] else ()):
                    self._put(f"{key}.{n}.w", krsc(g(f"{key}.{n}.weight")), True); self._put(f"{key}.{n}.b", g(f"{key}.{n}.bias"))
            elif kind == "attn":
                self._put(key + ".norm.g", g(key + ".norm.weight")); self._put(key + ".norm.b", g(key + ".norm.bias"))
                self._put(key + ".qkv.w", torch.cat([krsc(g(key + f".{n}.weight")) for n in "qkv"], 0), True)
                self._put(key + ".qkv.b", torch.cat([g(key + f".{n}.bias") for n in "qkv"], 0))
                self._put(key + ".po.w", krsc(g(key + ".proj_out.weight")), True); self._put(key + ".po.b", g(key + ".proj_out.bias"))
            else:
                self._put(key + ".w", krsc(g(key + ".weight")), True); self._put(key + ".b", g(key + ".bias"))
        self._put("norm_out.g", g("decoder.norm_out.weight")); self._put("norm_out.b", g("decoder.norm_out.bias"))
        self._put("conv_out.w", krsc(g("decoder.conv_out.weight")), True); self._put("conv_out.b", g("decoder.conv_out.bias"))
        self._loaded = True
        return self

    # ---- first_stage_model.decode -------------------------------------------------------------------------------------
    @torch.no_grad()
    def decode(self, z):
        """z: [B, embed_dim, h, w] fp32 latents ALREADY divided by scale_factor.  Returns [B, out_ch, 8h, 8w] fp32."""
        assert self._loaded, "load_state_dict() first"
        N, Cz, H, W = z.shape
        if Cz != self.embed_dim:
            raise ValueError(f"expected {self.embed_dim} latent channels, got {Cz}")
        ups = [l for l in self.layers if l[0] == "up"]
        # workspace: the library materialises the x2-upsampled input of every Upsample conv here
        need, hh = 1 << 20, H * W
        for l in ups:
            hh *= 4
            need = max(need, N * hh * l[2] * 2 + (2 << 20))
        self._ws = self._buf("ws", 1, -(-need // 4), torch.float32)
        zin = self._buf("z_in", N * H * W, 8)
        ops.nchw_to_nhwc(z.float().contiguous(), zin[:, :Cz])
        cpad = 64 if self._hi else self.zc
        pq = self._buf("pq", N * H * W, max(cpad, 8), zero=True)  # pad columns stay zero
        ops.conv2d(zin[:, :Cz], self.w["pq.w"], pq[:, :self.zc], N=N, H=H, W=W, R=1, S=1, pad=0, bias=self.w["pq.b"])
        lo, hi = self._needs(self.layers[0])
        cur = self._act("t0", N * H * W, self.block_in, H * W, lo, hi)
        self._conv(pq[:, :cpad], "conv_in", cur.lo, N, H, W, 3, y32=cur.hi, stats=cur.st)
        for i, layer in enumerate(self.layers):
            nxt = self.layers[i + 1] if i + 1 < len(self.layers) else None
            lo, hi = self._needs(nxt)
            kind = layer[0]
            if kind == "up":
                H, W = 2 * H, 2 * W
            cout = layer[3] if kind == "res" else layer[2]
            out = self._act(f"t{(i + 1) % 2}", N * H * W, cout, H * W, lo, hi)
            if kind == "res":
                self._res(layer, cur, out, N, H, W)
            elif kind == "attn":
                self._attn(layer, cur, out, N, H, W)
            else:
                self._conv(cur.lo, layer[1], out.lo, N, H // 2, W // 2, 3, upsample=True, y32=out.hi, stats=out.st)
            cur = out
        M = N * H * W
        t = self._buf("gn_a", M, self.c_last)
        self._gn(cur, t, N, self.w["norm_out.g"], self.w["norm_out.b"], True)
        img = self._buf("img", M, 8)
        self._conv(t, "conv_out", img[:, :self.out_ch], N, H, W, 3)
        out = torch.empty(N, self.out_ch, H, W, dtype=torch.float32, device=z.device)
        ops.nhwc_to_nchw(img[:, :self.out_ch], out)
        return out

    forward = decode


class B200FirstStageEncoder(_FirstStageNet):
    """Encode side of the first stage — the x_p entry of the path: ``get_z`` =
    ``scale_factor * encode_first_stage(x).sample()`` (``diffmk/makeup_diffuse.py:37-40``), feeding ``q_sample`` and
    ``reconstruct`` (``diffmk/diffusion_makeup.py:384-387``, ``diffmk/cddim.py:81-84``).  ``encode(x)`` returns the
    posterior moments (mean, logvar clamped to [-30, 20]) as NCHW fp32, like upstream's DiagonalGaussianDistribution.

    Same kernels as the decoder.  The upstream Downsample (zero-pad bottom / right by one, 3x3 stride-2 conv without
    padding) is ``mkd_conv_desc.pad = 0, pad_hi_extra = 1``; the 8 moment channels leave the tensor-core kernel in fp32
    and ``quant_conv`` runs in fp32 on them, so the latent is not rounded to bf16 on its way out."""

    def __init__(self, embed_dim=4, ddconfig=None, dtype=torch.bfloat16, **unused):
        super().__init__(dtype)
        dd = dict(ch=128, ch_mult=(1, 2, 4, 4), num_res_blocks=2, in_channels=3, z_channels=4, double_z=True)
        dd.update(ddconfig or {})
        if list(dd.get("attn_resolutions", [])):
            raise NotImplementedError("attn_resolutions other than [] (the yaml's value) are not implemented")
        self.embed_dim, self.dd = embed_dim, dd
        ch, mult, nrb = dd["ch"], tuple(dd["ch_mult"]), dd["num_res_blocks"]
        self.in_ch, self.ch = dd["in_channels"], ch
        self.mom = (2 if dd["double_z"] else 1) * dd["z_channels"]
        in_mult = (1,) + mult
        self.layers = []
        cin = ch
        for lvl in range(len(mult)):
            cin, cout = ch * in_mult[lvl], ch * mult[lvl]
            for i in range(nrb):
                self.layers.append(("res", f"encoder.down.{lvl}.block.{i}", cin, cout))
                cin = cout
            if lvl != len(mult) - 1:
                self.layers.append(("down", f"encoder.down.{lvl}.downsample.conv", cin))
        self.layers += [("res", "encoder.mid.block_1", cin, cin), ("attn", "encoder.mid.attn_1", cin),
                        ("res", "encoder.mid.block_2", cin, cin)]
        self.c_last = cin

    def upstream_shapes(self) -> dict:
        out = {"quant_conv.weight": (2 * self.embed_dim, self.mom, 1, 1), "quant_conv.bias": (2 * self.embed_dim,),
               "encoder.conv_in.weight": (self.ch, self.in_ch, 3, 3), "encoder.conv_in.bias": (self.ch,),
               "encoder.norm_out.weight": (self.c_last,), "encoder.norm_out.bias": (self.c_last,),
               "encoder.conv_out.weight": (self.mom, self.c_last, 3, 3), "encoder.conv_out.bias": (self.mom,)}
        for layer in self.layers:
            if layer[0] == "res":
                self._res_shapes(out, layer[1], layer[2], layer[3])
            elif layer[0] == "attn":
                self._attn_shapes(out, layer[1], layer[2])
            else:
                out.update({layer[1] + ".weight": (layer[2], layer[2], 3, 3), layer[1] + ".bias": (layer[2],)})
        return out

    def load_state_dict(self, sd, strict=True, prefix="first_stage_model.", device="cuda"):  # noqa: D401
        """upstream ``first_stage_model.*`` keys; only ``encoder.*`` and ``quant_conv`` are consumed"""
        self._device = torch.device(device)
        self._check_keys(sd, prefix, strict, "encoder")
        g = lambda k: sd[prefix + k]  # noqa: E731
        w_in = self._krsc(g("encoder.conv_in.weight"))
        if self._hi:  # 3 image channels -> 64 zero-padded input channels: conv_in runs on the tensor cores
            wp = torch.zeros(w_in.shape[0], 3, 3, 64, dtype=w_in.dtype, device=w_in.device)
            wp[..., :self.in_ch] = w_in
            w_in = wp
        self._put("conv_in.w", w_in, True); self._put("conv_in.b", g("encoder.conv_in.bias"))
        for layer in self.layers:
            if layer[0] == "res":
                self._put_res(g, layer[1], layer[2], layer[3])
            elif layer[0] == "attn":
                self._put_attn(g, layer[1])
            else:
                self._put(layer[1] + ".w", self._krsc(g(layer[1] + ".weight")), True); self._put(layer[1] + ".b", g(layer[1] + ".bias"))
        self._put("norm_out.g", g("encoder.norm_out.weight")); self._put("norm_out.b", g("encoder.norm_out.bias"))
        self._put("conv_out.w", self._krsc(g("encoder.conv_out.weight")), True); self._put("conv_out.b", g("encoder.conv_out.bias"))
        self._put("quant.w", self._krsc(g("quant_conv.weight"))); self._put("quant.b", g("quant_conv.bias"))  # fp32
        self._loaded = True
        return self

    @torch.no_grad()
    def encode(self, x):
        """x: [B, 3, H, W] fp32 images in [-1, 1].  Returns (mean, logvar), each [B, embed_dim, H/8, W/8] fp32."""
        assert self._loaded, "load_state_dict() first"
        N, Cx, H, W = x.shape
        if Cx != self.in_ch:
            raise ValueError(f"expected {self.in_ch} image channels, got {Cx}")
        # workspace: the library materialises the im2col matrix of every stride-2 Downsample conv here
        need, hh = 1 << 20, H * W
        for l in self.layers:
            if l[0] == "down":
                hh //= 4
                need = max(need, N * hh * 9 * l[2] * 2 + (2 << 20))
        self._ws = self._buf("ws", 1, -(-need // 4), torch.float32)
        cpad = 64 if self._hi else self.in_ch
        xin = self._buf("x_in", N * H * W, max(cpad, 8), zero=True)  # pad columns stay zero
        ops.nchw_to_nhwc(x.float().contiguous(), xin[:, :Cx])
        lo, hi = self._needs(self.layers[0])
        cur = self._act("t0", N * H * W, self.ch, H * W, lo, hi)
        self._conv(xin[:, :cpad], "conv_in", cur.lo, N, H, W, 3, y32=cur.hi, stats=cur.st)
        for i, layer in enumerate(self.layers):
            nxt = self.layers[i + 1] if i + 1 < len(self.layers) else None
            lo, hi = self._needs(nxt)
            kind = layer[0]
            Ho, Wo = (H // 2, W // 2) if kind == "down" else (H, W)
            cout = layer[3] if kind == "res" else layer[2]
            out = self._act(f"t{(i + 1) % 2}", N * Ho * Wo, cout, Ho * Wo, lo, hi)
            if kind == "res":
                self._res(layer, cur, out, N, H, W)
            elif kind == "attn":
                self._attn(layer, cur, out, N, H, W)
            else:
                ops.conv2d(cur.lo, self.w[layer[1] + ".w"], out.lo, N=N, H=H, W=W, R=3, S=3, stride=2, pad=0, pad_hi_extra=1,
                           bias=self.w[layer[1] + ".b"], workspace=self._ws, y32=out.hi, stats=out.st)
            cur, H, W = out, Ho, Wo
        M = N * H * W
        t = self._buf("gn_a", M, self.c_last)
        self._gn(cur, t, N, self.w["norm_out.g"], self.w["norm_out.b"], True)
        mom = self._buf("mom", M, 8 * -(-self.mom // 8), torch.float32)
        if self._hi:
            self._conv(t, "conv_out", None, N, H, W, 3, y32=mom[:, :self.mom])
        else:
            self._conv(t, "conv_out", mom[:, :self.mom], N, H, W, 3)
        q = self._buf("quant", M, 8 * -(-2 * self.embed_dim // 8), torch.float32)
        ops.conv2d(mom[:, :self.mom], self.w["quant.w"], q[:, :2 * self.embed_dim], N=N, H=H, W=W, R=1, S=1, pad=0,
                   bias=self.w["quant.b"])  # fp32 in, fp32 weights, fp32 out (generic kernel: 8 channels)
        out = torch.empty(N, 2 * self.embed_dim, H, W, dtype=torch.float32, device=x.device)
        ops.nhwc_to_nchw(q[:, :2 * self.embed_dim], out)
        mean, logvar = out[:, :self.embed_dim], out[:, self.embed_dim:]
        return mean.contiguous(), logvar.clamp(-30.0, 20.0).contiguous()

    forward = encode
