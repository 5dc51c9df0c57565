"""Modality encoders with the reference's constructor surface and state_dict names (encoders.py:17-283).

The modules hold parameters/buffers (so checkpoints of the reference load unchanged); the arithmetic of all five
encoder types — EmbeddedSequenceEncoder (CMU), TabularEncoder (TCGA), and SequenceEncoder / SparseTabularEncoder /
PatchEncoder (no shipped config uses them) — runs in the fused CUDA path driven by
mca_paper_b200.engine.Engine.encode(), which writes tokens straight into the packed [B, N, 512] buffer.
"""
from __future__ import annotations

import math
from typing import Optional

import torch
from torch import nn


class TokenEncoder(nn.Module):
    """nn.Embedding holder (encoders.py:17-37): padding_idx row is zero and gets no gradient, max_norm=1.0 rows are
    renormalised in place at every forward (done by mca_embedding_renorm)."""

    def __init__(self, num_embeddings: int, embedding_dim: int, padding_idx: Optional[int] = None,
                 max_norm: Optional[float] = 1.0, **kwargs):
        super().__init__()
        self.num_embeddings = num_embeddings
        self.embedding = nn.Embedding(num_embeddings, embedding_dim, padding_idx=padding_idx, max_norm=max_norm)


class ContinuousValueEncoder(nn.Module):
    """Parameter holder for encoders.py:40-72."""

    def __init__(self, d_model: int, dropout: float = 0.1, max_value: int = 512, padding_value=0.0, **kwargs):
        super().__init__()
        self.linear1 = nn.Linear(1, d_model)
        self.linear2 = nn.Linear(d_model, d_model)
        self.norm = nn.LayerNorm(d_model)
        self.max_value = max_value
        self.padding_value = padding_value


class PositionalEncoder(nn.Module):
    """Sinusoidal table as a persistent buffer `pe` (encoders.py:123-135)."""

    def __init__(self, d_model: int, dropout: float = 0.1, max_len: int = 2048, **kwargs):
        super().__init__()
        position = torch.arange(max_len).unsqueeze(1)
        div_term = torch.exp(torch.arange(0, d_model, 2) * (-math.log(10000.0) / d_model))
        pe = torch.zeros(max_len, d_model)
        pe[:, 0::2] = torch.sin(position * div_term)
        pe[:, 1::2] = torch.cos(position * div_term)
        self.register_buffer("pe", pe)


class _EncoderBase(nn.Module):
    kind = ""

    def _spec(self):
        raise NotImplementedError

    def forward(self, batch):
        """`(tokens [B, L, 512], attention_mask)` like the reference encoders.  Inside MCA.forward the engine runs the
        same kernels and writes the tokens straight into the packed [B, N, 512] buffer; this free-standing call goes
        through an encoders-only engine built around the module (standalone.encoder_forward)."""
        from .standalone import encoder_forward

        return encoder_forward(self, batch)


class EmbeddedSequenceEncoder(_EncoderBase):
    """encoders.py:169-214.  token_encoder = Sequential(LayerNorm(in), Linear(in, 512), LayerNorm(512))."""
    kind = "EmbeddedSequenceEncoder"

    def __init__(self, input_size=128, embedding_dim=512, padding_idx=0, dropout=0.0, max_tokens=1024, **kwargs):
        super().__init__()
        self.input_size, self.embedding_dim, self.max_tokens = input_size, embedding_dim, max_tokens
        self.token_encoder = nn.Sequential(nn.LayerNorm(input_size), nn.Linear(input_size, embedding_dim),
                                           nn.LayerNorm(embedding_dim))
        self.positional_encoder = PositionalEncoder(embedding_dim, dropout, max_tokens)

    def _spec(self):
        return {"type": self.kind, "input_size": self.input_size, "max_tokens": self.max_tokens}


class TabularEncoder(_EncoderBase):
    """encoders.py:75-96 (padding_idx=-1 flows into both the embedding and the value encoder's pad test)."""
    kind = "TabularEncoder"

    def __init__(self, num_embeddings=128, embedding_dim=512, padding_idx=-1, dropout=0.0, max_value=10000, **kwargs):
        super().__init__()
        self.num_embeddings, self.padding_idx, self.max_value = num_embeddings, padding_idx, max_value
        self.register_buffer("index", torch.arange(num_embeddings))
        self.token_encoder = TokenEncoder(num_embeddings, embedding_dim, padding_idx)
        self.value_encoder = ContinuousValueEncoder(embedding_dim, dropout, max_value, padding_idx)

    def _spec(self):
        return {"type": self.kind, "num_embeddings": self.num_embeddings, "max_tokens": self.num_embeddings,
                "max_value": self.max_value, "padding_idx": self.padding_idx}


class SparseTabularEncoder(_EncoderBase):
    """encoders.py:100-120: batch {'indices': int64 [B,L], 'data': f32 [B,L], 'attention_mask'}; tokens =
    Embedding(indices) + ContinuousValueEncoder(data) (its pad test is `data == padding_idx`)."""
    kind = "SparseTabularEncoder"

    def __init__(self, num_embeddings=36602, embedding_dim=512, padding_idx=0, dropout=0.0, max_value=10000,
                 max_tokens=1024, **kwargs):
        super().__init__()
        self.num_embeddings, self.padding_idx, self.max_value, self.max_tokens = num_embeddings, padding_idx, max_value, max_tokens
        self.token_encoder = TokenEncoder(num_embeddings, embedding_dim, padding_idx)
        self.value_encoder = ContinuousValueEncoder(embedding_dim, dropout, max_value, padding_idx)

    def _spec(self):
        return {"type": self.kind, "num_embeddings": self.num_embeddings, "max_tokens": self.max_tokens,
                "max_value": self.max_value, "padding_idx": self.padding_idx}


class SequenceEncoder(_EncoderBase):
    """encoders.py:145-166: batch {'tokens': int64 [B,L], 'attention_mask'}; tokens = Embedding(tokens) + sinusoidal PE."""
    kind = "SequenceEncoder"

    def __init__(self, num_embeddings=36602, embedding_dim=512, padding_idx=0, dropout=0.0, max_tokens=1024, **kwargs):
        super().__init__()
        self.num_embeddings, self.padding_idx, self.max_tokens = num_embeddings, padding_idx, max_tokens
        self.token_encoder = TokenEncoder(num_embeddings, embedding_dim, padding_idx)
        self.positional_encoder = PositionalEncoder(embedding_dim, dropout, max_tokens)

    def _spec(self):
        return {"type": self.kind, "num_embeddings": self.num_embeddings, "max_tokens": self.max_tokens,
                "padding_idx": self.padding_idx}


class PatchEncoder(_EncoderBase):
    """encoders.py:217-274, "matrix" mode (the only mode whose forward works in the reference: `image` / `video`
    never set `self.layer`, which its attention-mask line needs): values [B,H,W] -> patches 'b (h p1) (w p2) ->
    b (h w) (p1 p2)' -> LayerNorm -> Linear -> LayerNorm, + learned position embedding, dropout; the pad mask
    all(patch == pad_token) is computed by the encoder itself."""
    kind = "PatchEncoder"

    def __init__(self, patch_size=(16, 16), mode="matrix", num_channels=0, embedding_dim=512, max_tokens=1024,
                 dropout: float = 0.1, attn_mask=True, pad_token=-10000, **kwargs):
        super().__init__()
        assert mode in ["matrix", "image", "video"]
        if mode != "matrix":
            raise NotImplementedError("PatchEncoder modes 'image' / 'video' cannot run in the reference either "
                                      "(encoders.py:252-259 never define self.layer used at :273)")
        assert len(patch_size) == 2
        input_dim = 1
        for p in patch_size:
            input_dim *= p
        self.patch_size, self.mode, self.pad_token = tuple(patch_size), mode, -10000  # encoders.py:234 ignores the argument
        self.input_dim, self.max_tokens, self.p_drop, self.attn_mask = input_dim, max_tokens, float(dropout), attn_mask
        self.batch_to_tokens = nn.Sequential(nn.Identity(), nn.LayerNorm(input_dim), nn.Linear(input_dim, embedding_dim),
                                             nn.LayerNorm(embedding_dim))
        self.register_buffer("index", torch.arange(max_tokens))
        self.embedding = nn.Embedding(max_tokens, embedding_dim)

    def _spec(self):
        return {"type": self.kind, "patch_size": self.patch_size, "max_tokens": self.max_tokens, "dropout": self.p_drop}


encoders_dict = {
    "SequenceEncoder": SequenceEncoder,
    "TabularEncoder": TabularEncoder,
    "SparseTabularEncoder": SparseTabularEncoder,
    "PatchEncoder": PatchEncoder,
    "EmbeddedSequenceEncoder": EmbeddedSequenceEncoder,
}
