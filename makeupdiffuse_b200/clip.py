"""B200 text encoder of the conditioning path: ``FrozenCLIPEmbedder`` on the library's kernels.

Reference: yaml:109-110 instantiates ``ldm.modules.encoders.modules.FrozenCLIPEmbedder`` = HuggingFace ``CLIPTextModel``
(ViT-L/14 text tower, ``last_hidden_state``); ``get_learned_conditioning`` turns the constant prompt
``'makeup transfer'`` (``datasets.py:772``) into ``c_crossattn`` ``[B, 77, 768]`` (``makeup_controlnet.py:20``) and
``get_unconditional_conditioning`` the empty prompt (``diffusion_makeup.py:399-402``).  It runs once per prompt, not per
step, so the point here is the same parity bar on device, not speed: token + position embedding (``mkd_embed_tokens``),
12 pre-LN blocks — LayerNorm, fused q/k/v GEMM, causal attention (``mkd_attention_causal``), out-projection GEMM adding
into the fp32 stream in place, LayerNorm, fc1 GEMM with quick_gelu fused, fc2 GEMM adding in place — and the final
LayerNorm.  quick_gelu(x) = x sigmoid(1.702 x) = silu(1.702 x) / 1.702: the fc1 epilogue runs ``alpha = 1.702`` + SiLU and
the 1 / 1.702 is folded into fc2's weights when they are repacked.

Same module / state-dict names as HuggingFace (any prefix before ``text_model.`` is ignored, so
``cond_stage_model.transformer.text_model...`` checkpoint keys load).  The tokenizer's vocabulary files are not in this
image: ``encode(text)`` needs a caller-supplied ``tokenize`` callable, except for the empty prompt whose ids are fixed
(``[BOS, EOS, EOS, ...]``).  No CPU fallback.
"""
from __future__ import annotations

import torch

from . import _lib as L
from . import ops

CLIP_L_TEXT = dict(vocab_size=49408, hidden_size=768, num_hidden_layers=12, num_attention_heads=12,
                   intermediate_size=3072, max_position_embeddings=77, layer_norm_eps=1e-5)
BOS, EOS = 49406, 49407
_QG = 1.702


class B200FrozenCLIPEmbedder:
    def __init__(self, device="cuda", dtype=torch.bfloat16, max_length=77, tokenize=None, **cfg):
        self.cfg = {**CLIP_L_TEXT, **cfg}
        self.device, self.dtype, self.max_length, self.tokenize = torch.device(device), dtype, max_length, tokenize
        self.w = None
        ops.device_ok(self.device.index or 0)

    # ---- weights ------------------------------------------------------------------------------------------------
    def load_state_dict(self, sd):
        """HuggingFace CLIPTextModel keys; repacked once: q/k/v fused [3C, C], fc2 pre-scaled by 1 / 1.702"""
        sd = {k[k.index("text_model."):]: v for k, v in sd.items() if "text_model." in k}
        c = self.cfg
        dev, wd = self.device, self.dtype
        f32 = lambda t: t.detach().to(dev, torch.float32).contiguous()  # noqa: E731
        mat = lambda t: t.detach().to(dev, torch.float32).to(wd).contiguous()  # noqa: E731
        w = {"tok": f32(sd["text_model.embeddings.token_embedding.weight"]),
             "pos": f32(sd["text_model.embeddings.position_embedding.weight"]),
             "lnf.g": f32(sd["text_model.final_layer_norm.weight"]), "lnf.b": f32(sd["text_model.final_layer_norm.bias"])}
        used = 4
        for i in range(c["num_hidden_layers"]):
            p = f"text_model.encoder.layers.{i}."
            a = p + "self_attn."
            w[f"{i}.qkv.w"] = mat(torch.cat([sd[a + "q_proj.weight"], sd[a + "k_proj.weight"], sd[a + "v_proj.weight"]], 0))
            w[f"{i}.qkv.b"] = f32(torch.cat([sd[a + "q_proj.bias"], sd[a + "k_proj.bias"], sd[a + "v_proj.bias"]], 0))
            w[f"{i}.o.w"], w[f"{i}.o.b"] = mat(sd[a + "out_proj.weight"]), f32(sd[a + "out_proj.bias"])
            w[f"{i}.fc1.w"], w[f"{i}.fc1.b"] = mat(sd[p + "mlp.fc1.weight"]), f32(sd[p + "mlp.fc1.bias"])
            w[f"{i}.fc2.w"] = mat(sd[p + "mlp.fc2.weight"].detach().to(torch.float32) / _QG)
            w[f"{i}.fc2.b"] = f32(sd[p + "mlp.fc2.bias"])
            for n in ("layer_norm1", "layer_norm2"):
                w[f"{i}.{n}.g"], w[f"{i}.{n}.b"] = f32(sd[p + n + ".weight"]), f32(sd[p + n + ".bias"])
            used += 16
        extra = [k for k in sd if not k.endswith("position_ids")]
        if len(extra) != used:
            raise KeyError(f"unexpected CLIP text-model keys: {len(extra)} tensors, {used} consumed")
        assert w["tok"].shape == (c["vocab_size"], c["hidden_size"]) and w["pos"].shape[0] >= self.max_length
        self.w = w
        return self

    # ---- FrozenCLIPEmbedder surface ----------------------------------------------------------------------------
    def empty_prompt_tokens(self, batch):
        t = torch.full((batch, self.max_length), EOS, dtype=torch.long, device=self.device)
        t[:, 0] = BOS
        return t

    def encode(self, text):
        if isinstance(text, str):
            text = [text]
        if all(t == "" for t in text):
            return self(self.empty_prompt_tokens(len(text)))
        if self.tokenize is None:
            raise RuntimeError("the CLIP BPE vocabulary is not available here: construct B200FrozenCLIPEmbedder with "
                               "tokenize=<callable: list[str] -> int64 [B, 77]> or call it with token ids")
        return self(torch.as_tensor(self.tokenize(text), dtype=torch.long, device=self.device))

    @torch.no_grad()
    def __call__(self, tokens):
        return self.forward(tokens)

    @torch.no_grad()
    def forward(self, tokens):
        """tokens [B, T <= 77] int64 (device) -> last_hidden_state [B, T, C] fp32"""
        if self.w is None:
            raise RuntimeError("load_state_dict() first")
        c, w, dev = self.cfg, self.w, self.device
        tokens = tokens.to(dev).contiguous()
        B, T = tokens.shape
        C, H, I = c["hidden_size"], c["num_attention_heads"], c["intermediate_size"]
        M, hd, eps = B * T, C // H, c["layer_norm_eps"]
        hi = self.dtype != torch.float32          # bf16 path: the residual stream stays fp32, GEMM operands are bf16
        act = self.dtype
        xs = torch.empty(M, C, device=dev, dtype=torch.float32)
        ln = torch.empty(M, C, device=dev, dtype=act)
        qkv = torch.empty(M, 3 * C, device=dev, dtype=act)
        att = torch.empty(M, C, device=dev, dtype=act)
        h = torch.empty(M, I, device=dev, dtype=act)
        ws = torch.empty(64 << 20, device=dev, dtype=torch.uint8)
        lin = dict(N=1, H=1, W=M, workspace=ws)
        into_xs = dict(residual=xs, y32=xs) if hi else dict(residual=xs)   # x += ... in place
        y_xs = None if hi else xs
        ops.embed_tokens(tokens, w["tok"], w["pos"], xs)
        for i in range(c["num_hidden_layers"]):
            ops.layernorm(xs, ln, w[f"{i}.layer_norm1.g"], w[f"{i}.layer_norm1.b"], eps)
            ops.conv2d(ln, w[f"{i}.qkv.w"], qkv, bias=w[f"{i}.qkv.b"], **lin)
            ops.attention_causal(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], att, B=B, heads=H, N=T, d=hd, scale=hd ** -0.5)
            ops.conv2d(att, w[f"{i}.o.w"], y_xs, bias=w[f"{i}.o.b"], **into_xs, **lin)
            ops.layernorm(xs, ln, w[f"{i}.layer_norm2.g"], w[f"{i}.layer_norm2.b"], eps)
            ops.conv2d(ln, w[f"{i}.fc1.w"], h, bias=w[f"{i}.fc1.b"], alpha=_QG, act=L.ACT_SILU, **lin)  # silu(1.702 x)
            ops.conv2d(h, w[f"{i}.fc2.w"], y_xs, bias=w[f"{i}.fc2.b"], **into_xs, **lin)                # (W2 / 1.702) . + x
        out = torch.empty(M, C, device=dev, dtype=torch.float32)
        ops.layernorm(xs, out, w["lnf.g"], w["lnf.b"], eps)
        return out.view(B, T, C)
