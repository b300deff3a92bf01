"""Batched greedy caption decode over a per-image cached vision K/V.

Mirrors the greedy branch of `FullModel.generate_caption` (full_model.py:241-363, `do_sample=False`):
start from BOS, every step embeds the whole prefix, runs the bridge over it (the bridge
self-attention is non-causal, so earlier positions change when a token is appended and the prefix
must be recomputed -- only the image's K/V are exactly cacheable), feeds the result to the language
model, takes `argmax` of the last position's logits (NaN logits -> zeros, Inf logits -> clamped to
+-100, as :270-283) and appends it.

The reference loop is batch-1 and synchronises with the host up to three times per step
(`next_token.item()`, :317,355-366). Here B images decode together with no host synchronisation:
every row runs `max_new_tokens` steps and is cut at its first EOS afterwards, which is what B
independent reference runs produce (generation is deterministic and rows never interact).

`embed_fn` / `lm_fn` stand for the frozen language model, which is outside the hot path:
`embed_fn(ids [B, s]) -> [B, s, D]` (LanguageModel.get_embeddings, language_model.py:146-164) and
`lm_fn(hidden [B, s, D]) -> logits [B, s, V]` or `[B, V]` for the last position
(LanguageModel.forward_from_embeddings, :111-144).
"""
from __future__ import annotations

from typing import Callable, Optional

import torch

from .kv_cache import VisionKVCache

__all__ = ["greedy_decode"]


@torch.no_grad()
def greedy_decode(bridge, vision_features: torch.Tensor, embed_fn: Callable[[torch.Tensor], torch.Tensor],
                  lm_fn: Callable[[torch.Tensor], torch.Tensor], *, bos_token_id: int,
                  eos_token_id: Optional[int] = None, max_new_tokens: int = 50,
                  kv_cache: Optional[VisionKVCache] = None, use_cache: bool = True):
    """Returns (ids [B, 1 + max_new_tokens] int64 incl. BOS, lengths [B] int64): row b's caption is
    ids[b, 1:lengths[b]] (EOS excluded); positions from lengths[b] on are what the lock-step loop kept
    generating and are to be ignored."""
    was_training = bridge.training
    bridge.eval()
    try:
        B = vision_features.shape[0]
        dev = vision_features.device
        if use_cache and kv_cache is None:
            kv_cache = VisionKVCache(bridge, vision_features)
        ids = torch.empty((B, 1 + max_new_tokens), dtype=torch.long, device=dev)
        ids[:, 0] = bos_token_id
        for step in range(max_new_tokens):
            prefix = ids[:, :step + 1]
            hidden = bridge(vision_features, embed_fn(prefix), kv_cache=kv_cache if use_cache else None)
            logits = lm_fn(hidden)
            if logits.dim() == 3:
                logits = logits[:, -1, :]
            logits = logits.float()
            # numerical guards of the reference, as tensor ops (no host sync)
            bad = torch.isnan(logits).any()
            logits = torch.where(bad, torch.zeros_like(logits), logits)
            logits = torch.where(torch.isinf(logits).any(), logits.clamp(min=-100, max=100), logits)
            ids[:, step + 1] = torch.argmax(logits, dim=-1)
        if eos_token_id is None:
            lengths = torch.full((B,), 1 + max_new_tokens, dtype=torch.long, device=dev)
        else:
            is_eos = ids[:, 1:] == eos_token_id
            first = torch.where(is_eos.any(dim=1), is_eos.float().argmax(dim=1), torch.full((B,), max_new_tokens, device=dev))
            lengths = first + 1
        return ids, lengths
    finally:
        bridge.train(was_training)
