"""BERT-base style encoder (Devlin et al.) on ``torch.nn`` + SDPA.

12 layers, hidden 768, 12 heads, FFN 3072, vocab 30522, learned positions,
post-LayerNorm -- ~110 M parameters with the MLM head tied to the embeddings.
Used by the P2P GeometricMedian target config (BASELINE.json config 4); random
init, synthetic token batches (no network for checkpoints/datasets).  The reference ships no
transformer (its only model is the SmallCNN of reference examples/ps/nodes.py:46-61); its P2P step
functions (reference engine/node/mixin.py:59-80) are model-agnostic, which is all this needs.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch
import torch.nn as nn

from ..ops.fused_layers import ArenaLinear as _Linear   # nn.Linear unless a device worker enables direct gradients
import torch.nn.functional as F


@dataclass
class BertConfig:
    """Sizes of a BERT encoder; the defaults are BERT-base (12 layers, hidden 768, 12 heads, FFN 3072, vocabulary
    30 522, 512 positions).  ``dropout`` defaults to 0 so that a captured training step is deterministic.
    """

    vocab_size: int = 30522
    hidden: int = 768
    layers: int = 12
    heads: int = 12
    ffn: int = 3072
    max_pos: int = 512
    type_vocab: int = 2
    dropout: float = 0.0
    eps: float = 1e-12


class _Block(nn.Module):
    def __init__(self, c: BertConfig):
        super().__init__()
        self.heads = c.heads
        self.qkv = _Linear(c.hidden, 3 * c.hidden)
        self.proj = _Linear(c.hidden, c.hidden)
        self.ln1 = nn.LayerNorm(c.hidden, eps=c.eps)
        self.fc1 = _Linear(c.hidden, c.ffn)
        self.fc2 = _Linear(c.ffn, c.hidden)
        self.ln2 = nn.LayerNorm(c.hidden, eps=c.eps)
        self.drop = c.dropout

    def forward(self, x: torch.Tensor, mask: torch.Tensor | None) -> torch.Tensor:
        B, S, H = x.shape
        q, k, v = self.qkv(x).view(B, S, 3, self.heads, H // self.heads).permute(2, 0, 3, 1, 4)
        a = F.scaled_dot_product_attention(q, k, v, attn_mask=mask, dropout_p=self.drop if self.training else 0.0)
        x = self.ln1(x + self.proj(a.transpose(1, 2).reshape(B, S, H)))
        return self.ln2(x + self.fc2(F.gelu(self.fc1(x))))


class BertEncoder(nn.Module):
    """BERT encoder stack: token + position + type embeddings, LayerNorm, then ``config.layers`` post-LN transformer
    blocks (fused QKV projection, ``scaled_dot_product_attention``, GELU FFN).

    ``forward(ids, attention_mask=None)`` takes ``(batch, seq)`` token ids and an optional ``(batch, seq)`` 0/1 mask and
    returns ``(batch, seq, hidden)`` states.  The linear layers are arena-aware: inside a device worker their weight
    gradients are written straight into the node's gradient arena.
    """

    def __init__(self, config: BertConfig | None = None):
        super().__init__()
        c = config or BertConfig()
        self.config = c
        self.tok = nn.Embedding(c.vocab_size, c.hidden)
        self.pos = nn.Embedding(c.max_pos, c.hidden)
        self.typ = nn.Embedding(c.type_vocab, c.hidden)
        self.ln = nn.LayerNorm(c.hidden, eps=c.eps)
        self.blocks = nn.ModuleList(_Block(c) for _ in range(c.layers))
        self.apply(self._init)

    @staticmethod
    def _init(m: nn.Module) -> None:
        if isinstance(m, (nn.Linear, nn.Embedding)):
            nn.init.normal_(m.weight, std=0.02)
            if isinstance(m, nn.Linear) and m.bias is not None:
                nn.init.zeros_(m.bias)

    def forward(self, ids: torch.Tensor, attention_mask: torch.Tensor | None = None,
                token_type: torch.Tensor | None = None) -> torch.Tensor:
        B, S = ids.shape
        pos = torch.arange(S, device=ids.device)
        x = self.tok(ids) + self.pos(pos)[None]
        x = x + (self.typ(token_type) if token_type is not None else self.typ.weight[0])
        x = self.ln(x)
        mask = None
        if attention_mask is not None:
            mask = attention_mask[:, None, None, :].to(torch.bool)
        for blk in self.blocks:
            x = blk(x, mask)
        return x


class BertForMaskedLM(nn.Module):
    """Encoder + transform + decoder tied to the token embedding (BERT-base MLM head)."""

    def __init__(self, config: BertConfig | None = None):
        super().__init__()
        self.bert = BertEncoder(config)
        c = self.bert.config
        self.transform = _Linear(c.hidden, c.hidden)
        self.ln = nn.LayerNorm(c.hidden, eps=c.eps)
        self.bias = nn.Parameter(torch.zeros(c.vocab_size))

    def forward(self, ids: torch.Tensor, attention_mask: torch.Tensor | None = None) -> torch.Tensor:
        h = self.ln(F.gelu(self.transform(self.bert(ids, attention_mask))))
        return F.linear(h, self.bert.tok.weight, self.bias)


def bert_base(**overrides) -> BertForMaskedLM:
    """BERT-base with a masked-language-model head (110 M parameters); keyword arguments override
    :class:`BertConfig` fields, e.g. ``bert_base(layers=2, hidden=128, heads=2, ffn=512)`` for tests."""
    return BertForMaskedLM(BertConfig(**overrides))
