"""Oracle of the conditioning producer: the text encoder behind ``get_learned_conditioning``.

The reference instantiates ``ldm.modules.encoders.modules.FrozenCLIPEmbedder`` (yaml:109-110), i.e. HuggingFace's
``CLIPTextModel`` of openai/clip-vit-large-patch14 (``layer="last"`` -> ``last_hidden_state``), and feeds its output as
``c_crossattn`` (``makeup_controlnet.py:20``: "[bs, 77, 768]"; ``makeup_teacher.py:33-42``).  This is a plain fp32
restatement of that module with HuggingFace's parameter names, so a ``CLIPTextModel.state_dict()`` loads unchanged.

PINNED: ``transformers`` (5.5, the library the reference wraps) is importable in this image, so
``tests/test_clip.py`` holds this restatement to ``transformers.CLIPTextModel`` itself on identical seeded weights
(max abs difference ~1e-6 at the full ViT-L/14 text size).

Test infrastructure (see oracle/__init__.py).
"""
from __future__ import annotations

import torch
import torch.nn as nn

CLIP_L_TEXT = dict(vocab_size=49408, hidden_size=768, num_hidden_layers=12, num_attention_heads=12,
                   intermediate_size=3072, max_position_embeddings=77, layer_norm_eps=1e-5)
BOS, EOS = 49406, 49407  # <|startoftext|>, <|endoftext|> (also the pad token of the CLIP tokenizer)


def empty_prompt_tokens(batch: int, length: int = 77) -> torch.Tensor:
    """token ids of the prompt "" padded to max_length: [BOS, EOS, EOS, ...] — what get_unconditional_conditioning
    encodes (diffusion_makeup.py:399-402)"""
    t = torch.full((batch, length), EOS, dtype=torch.long)
    t[:, 0] = BOS
    return t


class _Attn(nn.Module):
    def __init__(self, c, heads):
        super().__init__()
        self.heads, self.scale = heads, (c // heads) ** -0.5
        self.q_proj, self.k_proj, self.v_proj, self.out_proj = (nn.Linear(c, c) for _ in range(4))

    def forward(self, x):
        B, T, C = x.shape
        sp = lambda t: t.view(B, T, self.heads, C // self.heads).transpose(1, 2)  # noqa: E731
        q, k, v = sp(self.q_proj(x) * self.scale), sp(self.k_proj(x)), sp(self.v_proj(x))
        s = q @ k.transpose(-1, -2)
        s = s + torch.full((T, T), float("-inf"), device=x.device).triu(1)  # causal: key j visible iff j <= i
        o = torch.softmax(s, dim=-1) @ v
        return self.out_proj(o.transpose(1, 2).reshape(B, T, C))


class _MLP(nn.Module):
    def __init__(self, c, inner):
        super().__init__()
        self.fc1, self.fc2 = nn.Linear(c, inner), nn.Linear(inner, c)

    def forward(self, x):
        h = self.fc1(x)
        return self.fc2(h * torch.sigmoid(1.702 * h))  # quick_gelu


class _Layer(nn.Module):
    def __init__(self, c, heads, inner, eps):
        super().__init__()
        self.self_attn = _Attn(c, heads)
        self.layer_norm1 = nn.LayerNorm(c, eps=eps)
        self.mlp = _MLP(c, inner)
        self.layer_norm2 = nn.LayerNorm(c, eps=eps)

    def forward(self, x):
        x = x + self.self_attn(self.layer_norm1(x))
        return x + self.mlp(self.layer_norm2(x))


class _Embeddings(nn.Module):
    def __init__(self, vocab, c, max_len):
        super().__init__()
        self.token_embedding = nn.Embedding(vocab, c)
        self.position_embedding = nn.Embedding(max_len, c)

    def forward(self, ids):
        return self.token_embedding(ids) + self.position_embedding.weight[: ids.shape[1]]


class _Encoder(nn.Module):
    def __init__(self, n, c, heads, inner, eps):
        super().__init__()
        self.layers = nn.ModuleList([_Layer(c, heads, inner, eps) for _ in range(n)])


class _TextModel(nn.Module):
    def __init__(self, vocab_size, hidden_size, num_hidden_layers, num_attention_heads, intermediate_size,
                 max_position_embeddings, layer_norm_eps):
        super().__init__()
        self.embeddings = _Embeddings(vocab_size, hidden_size, max_position_embeddings)
        self.encoder = _Encoder(num_hidden_layers, hidden_size, num_attention_heads, intermediate_size, layer_norm_eps)
        self.final_layer_norm = nn.LayerNorm(hidden_size, eps=layer_norm_eps)


class OracleCLIPTextEncoder(nn.Module):
    """``forward(tokens [B, T] int64) -> last_hidden_state [B, T, C]`` (FrozenCLIPEmbedder.forward after tokenisation)"""

    def __init__(self, **cfg):
        super().__init__()
        self.cfg = {**CLIP_L_TEXT, **cfg}
        self.text_model = _TextModel(**self.cfg)

    @torch.no_grad()
    def forward(self, tokens):
        tm = self.text_model
        x = tm.embeddings(tokens)
        for layer in tm.encoder.layers:
            x = layer(x)
        return tm.final_layer_norm(x)
