"""BERT-base sequence classifier on the sm_100a layers (BASELINE.json config 3: the large
delta-reduce that stresses the NVLink roofline -- ~109.5 M parameters, 219 MB in bf16).

Standard post-LN encoder (embeddings -> 12 x [self-attention, FFN] -> pooler -> classifier) with
the usual parameter names (``bert.embeddings.word_embeddings.weight``,
``bert.encoder.layer.N.attention.self.query.weight`` ... are folded into one packed
``attention.qkv`` projection here for a single GEMM; ``load_hf_state_dict`` / ``hf_state_dict`` map a stock
Hugging-Face ``BertForSequenceClassification`` state_dict onto it and back, so HF checkpoints load and our
checkpoints stay loadable by HF -- tests/test_bert_hf_compat.py).  Every matmul is the tcgen05
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

    def forward(self, x, B, S, mask_bias=None):
        a = bnn.attention(self.qkv(x), B, S, self.H, self.dh, mask_bias=mask_bias)
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

    def forward(self, input_ids, attention_mask=None, token_type_ids=None):
        """``input_ids``: ``[B, S]`` int64 -> logits ``[B, num_labels]`` (fp32).  ``attention_mask`` (``[B, S]``, 1 =
        attend, 0 = padding) and ``token_type_ids`` follow the Hugging-Face call convention; without a mask the
        attention core takes its unmasked fast path."""
        B, S = input_ids.shape
        pos, typ = self._ids(B, S, input_ids.device)
        if token_type_ids is not None:
            typ = token_type_ids.reshape(-1)
        mask_bias = None
        if attention_mask is not None:
            mask_bias = (1.0 - attention_mask.to(torch.float32)) * -30000.0      # additive, finite in bf16
        x = self.embeddings(input_ids, pos, typ)
        for layer in self.layers:
            x = layer(x, B, S, mask_bias)
        first = x.view(B, S, -1)[:, 0].contiguous()
        pooled = torch.tanh(self.pooler(first).float())
        if pooled.is_cuda:
            pooled = pooled.to(torch.bfloat16)
        return self.classifier(pooled)


    # ------------------------------------------------------------------ Hugging-Face checkpoint compatibility
    _HF_LAYER = (("attention.output.dense", "attn_out"), ("attention.output.LayerNorm", "attn_ln"),
                 ("intermediate.dense", "ffn_in"), ("output.dense", "ffn_out"), ("output.LayerNorm", "ffn_ln"))

    @torch.no_grad()
    def load_hf_state_dict(self, hf_state: dict, strict: bool = True) -> None:
        """Load a stock ``transformers.BertForSequenceClassification`` ``state_dict``: the separate query / key /
        value projections are concatenated (rows ``[q; k; v]``) into the packed ``qkv`` GEMM, everything else is a
        rename.  Note: this model uses the tanh GELU (``hidden_act="gelu_pytorch_tanh"`` in HF terms)."""
        own = {}
        sd = dict(hf_state)
        for k in ("word_embeddings", "position_embeddings", "token_type_embeddings"):
            own["embeddings.{}.weight".format(k)] = sd.pop("bert.embeddings.{}.weight".format(k))
        for p in ("weight", "bias"):
            own["embeddings.LayerNorm." + p] = sd.pop("bert.embeddings.LayerNorm." + p)
        for i in range(self.config.num_hidden_layers):
            hf, me = "bert.encoder.layer.{}.".format(i), "layers.{}.".format(i)
            for p in ("weight", "bias"):
                own[me + "qkv." + p] = torch.cat([sd.pop(hf + "attention.self.{}.{}".format(n, p))
                                                  for n in ("query", "key", "value")], dim=0)
                for a, b in self._HF_LAYER:
                    own[me + b + "." + p] = sd.pop(hf + a + "." + p)
        for p in ("weight", "bias"):
            own["pooler." + p] = sd.pop("bert.pooler.dense." + p)
            own["classifier." + p] = sd.pop("classifier." + p)
        sd.pop("bert.embeddings.position_ids", None)
        sd.pop("bert.embeddings.token_type_ids", None)
        if strict and sd:
            raise KeyError("unexpected Hugging-Face keys: {}".format(sorted(sd)[:5]))
        self.load_state_dict(own, strict=strict)
        arena = getattr(self, "_arena", None)
        if arena is not None:
            arena.commit_global()

    def hf_state_dict(self) -> dict:
        """Inverse of :meth:`load_hf_state_dict`: a ``state_dict`` a stock Hugging-Face model loads with ``strict=True``."""
        own = {k: v.detach() for k, v in self.state_dict().items()}
        D = self.config.hidden_size
        out = {}
        for k in ("word_embeddings", "position_embeddings", "token_type_embeddings"):
            out["bert.embeddings.{}.weight".format(k)] = own["embeddings.{}.weight".format(k)]
        for p in ("weight", "bias"):
            out["bert.embeddings.LayerNorm." + p] = own["embeddings.LayerNorm." + p]
        for i in range(self.config.num_hidden_layers):
            hf, me = "bert.encoder.layer.{}.".format(i), "layers.{}.".format(i)
            for p in ("weight", "bias"):
                q = own[me + "qkv." + p]
                for j, n in enumerate(("query", "key", "value")):
                    out[hf + "attention.self.{}.{}".format(n, p)] = q[j * D:(j + 1) * D].clone()
                for a, b in self._HF_LAYER:
                    out[hf + a + "." + p] = own[me + b + "." + p]
        for p in ("weight", "bias"):
            out["bert.pooler.dense." + p] = own["pooler." + p]
            out["classifier." + p] = own["classifier." + p]
        return out


def bert_base(num_labels: int = 2, **kw) -> BertForSequenceClassification:
    return BertForSequenceClassification(BertConfig(num_labels=num_labels, **kw))


def bert_tiny(num_labels: int = 2) -> BertForSequenceClassification:
    """2-layer, 128-wide model for tests."""
    return BertForSequenceClassification(BertConfig(vocab_size=1024, hidden_size=128, num_hidden_layers=2,
                                                    num_attention_heads=2, intermediate_size=512,
                                                    max_position_embeddings=128, num_labels=num_labels), name="bert_tiny")
