"""BERT-base sequence classifier on the sm_100a layers (BASELINE.json config 3: the large
delta-reduce that stresses the NVLink roofline -- ~109.5 M parameters, 219 MB in bf16).

Standard post-LN encoder (embeddings -> 12 x [self-attention, FFN] -> pooler -> classifier) with
the usual parameter names (``bert.embeddings.word_embeddings.weight``,
``bert.encoder.layer.N.attention.self.query.weight`` ... are folded into one packed
``attention.qkv`` projection here for a single GEMM; ``load_hf_state_dict`` maps a stock
Hugging-Face ``BertForSequenceClassification`` state_dict onto it).  Every matmul is the tcgen05
GEMM (GELU fused in the epilogue), attention is four strided-batched GEMMs + the softmax kernel on
the packed QKV buffer, LayerNorm fuses the residual add.  Dropout is omitted (p = 0): the
reference has none and synthetic-shard benchmarking does not want the noise.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import torch
from torch import nn

from ..ops import nn as bnn
from .base import FederatedModule


@dataclass
class BertConfig:
    vocab_size: int = 30522
    hidden_size: int = 768
    num_hidden_layers: int = 12
    num_attention_heads: int = 12
    intermediate_size: int = 3072
    max_position_embeddings: int = 512
    type_vocab_size: int = 2
    layer_norm_eps: float = 1e-12
    num_labels: int = 2


class BertEmbeddings(nn.Module):
    def __init__(self, c: BertConfig):
        super().__init__()
        self.word_embeddings = bnn.Embedding(c.vocab_size, c.hidden_size)
        self.position_embeddings = bnn.Embedding(c.max_position_embeddings, c.hidden_size)
        self.token_type_embeddings = bnn.Embedding(c.type_vocab_size, c.hidden_size)
        self.LayerNorm = bnn.LayerNorm(c.hidden_size, c.layer_norm_eps)

    def forward(self, ids, pos_ids, type_ids):
        w = self.word_embeddings(ids)
        p = self.position_embeddings(pos_ids)
        t = self.token_type_embeddings(type_ids)
        if w.is_cuda:
            from ..ops import functional as F
            pt = _Add.apply(p, t)
            return self.LayerNorm(w, pt)
        return self.LayerNorm(w + p + t)


class _Add(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b):
        from ..ops import functional as F
        return F.add(a, b)

    @staticmethod
    def backward(ctx, g):
        return g, g


class BertLayer(nn.Module):
    def __init__(self, c: BertConfig):
        super().__init__()
        self.H, self.dh = c.num_attention_heads, c.hidden_size // c.num_attention_heads
        self.qkv = bnn.Linear(c.hidden_size, 3 * c.hidden_size)
        self.attn_out = bnn.Linear(c.hidden_size, c.hidden_size)
        self.attn_ln = bnn.LayerNorm(c.hidden_size, c.layer_norm_eps)
        self.ffn_in = bnn.Linear(c.hidden_size, c.intermediate_size, act="gelu")
        self.ffn_out = bnn.Linear(c.intermediate_size, c.hidden_size)
        self.ffn_ln = bnn.LayerNorm(c.hidden_size, c.layer_norm_eps)

    def forward(self, x, B, S):
        a = bnn.attention(self.qkv(x), B, S, self.H, self.dh)
        x = self.attn_ln(self.attn_out(a), x)
        return self.ffn_ln(self.ffn_out(self.ffn_in(x)), x)


class BertForSequenceClassification(FederatedModule):
    name = "bert_base"
    loss_kind = "ce"
    default_lr = 0.01
    default_batch_size = 32

    def __init__(self, config: Optional[BertConfig] = None, name: Optional[str] = None):
        super().__init__()
        self.config = c = config or BertConfig()
        if name:
            self.name = name
        self.embeddings = BertEmbeddings(c)
        self.layers = nn.ModuleList([BertLayer(c) for _ in range(c.num_hidden_layers)])
        self.pooler = bnn.Linear(c.hidden_size, c.hidden_size)
        self.classifier = bnn.Linear(c.hidden_size, c.num_labels, out_fp32=True)
        self._static = {}
        for m in self.modules():
            if isinstance(m, bnn.Linear):
                nn.init.normal_(m.weight, std=0.02)
                if m.bias is not None:
                    nn.init.zeros_(m.bias)

    def _ids(self, B, S, device):
        key = (B, S, str(device))
        if key not in self._static:
            pos = torch.arange(S, device=device).repeat(B)
            self._static[key] = (pos, torch.zeros(B * S, dtype=torch.long, device=device))
        return self._static[key]

    def forward(self, input_ids):
        """``input_ids``: ``[B, S]`` int64 -> logits ``[B, num_labels]`` (fp32)."""
        B, S = input_ids.shape
        pos, typ = self._ids(B, S, input_ids.device)
        x = self.embeddings(input_ids, pos, typ)
        for layer in self.layers:
            x = layer(x, B, S)
        first = x.view(B, S, -1)[:, 0].contiguous()
        pooled = torch.tanh(self.pooler(first).float())
        if pooled.is_cuda:
            pooled = pooled.to(torch.bfloat16)
        return self.classifier(pooled)


def bert_base(num_labels: int = 2, **kw) -> BertForSequenceClassification:
    return BertForSequenceClassification(BertConfig(num_labels=num_labels, **kw))


def bert_tiny(num_labels: int = 2) -> BertForSequenceClassification:
    """2-layer, 128-wide model for tests."""
    return BertForSequenceClassification(BertConfig(vocab_size=1024, hidden_size=128, num_hidden_layers=2,
                                                    num_attention_heads=2, intermediate_size=512,
                                                    max_position_embeddings=128, num_labels=num_labels), name="bert_tiny")
