"""Host-side bookkeeping of the decode caches (no GPU): which cache layout / kernel a cross-attention over
L positions reads, and the validity rules of the per-position rows."""
import types

import pytest
import torch


def _cache(max_positions=8, batch=2, dim=64, filled=0):
    from vlm_bridge_b200 import VisionKVCache

    c = object.__new__(VisionKVCache)                      # the constructor projects K/V on the GPU
    c.batch, c.len_vision, c.max_positions, c.positions_filled = batch, 5, max_positions, filled
    c.x1 = torch.empty(batch, max_positions, dim) if max_positions else None
    c.kv, c.kv_packed, c.kv_tc = "kv", "packed", "tc"
    return c


def test_position_rows_rules():
    c = _cache()
    dev = torch.device("cpu")
    with pytest.raises(RuntimeError, match="only 0 positions"):
        c.position_rows(3, 2, dev, 64)                     # rows 0..1 were never written
    assert c.position_rows(3, 0, dev, 64) is c.x1 and c.positions_filled == 3
    assert c.position_rows(4, 3, dev, 64) is c.x1 and c.positions_filled == 4
    assert c.position_rows(2, 1, dev, 64) is c.x1 and c.positions_filled == 2    # a shorter prefix shrinks the valid range
    with pytest.raises(RuntimeError, match="only 2 positions"):
        c.position_rows(5, 4, dev, 64)
    for bad_k in (-1, 4, 7):
        with pytest.raises(RuntimeError):
            c.position_rows(4, bad_k, dev, 64)             # 0 <= k < length
    assert c.position_rows(9, 2, dev, 64) is None          # longer than the store: caller computes every row
    with pytest.raises(RuntimeError, match="width"):
        c.position_rows(2, 1, dev, 128)
    assert _cache(max_positions=0).position_rows(2, 0, dev, 64) is None          # store disabled


def test_kernel_choice_per_prefix_length():
    from vlm_bridge_b200 import BridgeLite
    from vlm_bridge_b200.bridge import FLAG_KV_PACKED, FLAG_KV_TC, TC_DECODE_MIN_LEN

    torch.manual_seed(0)
    m = BridgeLite(vision_dim=32, language_dim=128, num_heads_cross=2, num_heads_self=1, dropout=0.1).eval()   # d = 64
    c = _cache()
    assert m._kv_flag(c, 1, False) == ("packed", FLAG_KV_PACKED)
    assert m._kv_flag(c, TC_DECODE_MIN_LEN - 1, False) == ("packed", FLAG_KV_PACKED)
    assert m._kv_flag(c, TC_DECODE_MIN_LEN, False) == ("tc", FLAG_KV_TC)
    assert m._kv_flag(c, 64, False) == ("tc", FLAG_KV_TC)
    assert m._kv_flag(c, 65, False) == ("kv", 0)           # longer prefixes read the projection output
    assert m._kv_flag(c, 8, True) == ("kv", 0)             # a forward that keeps activations for backward
    c.kv_tc = None
    assert m._kv_flag(c, 40, False) == ("packed", FLAG_KV_PACKED)
    m.train()
    assert m._kv_flag(c, 8, False) == ("kv", 0)            # active dropout: the training kernel draws the masks
    m2 = BridgeLite(vision_dim=32, language_dim=96, num_heads_cross=2, num_heads_self=1).eval()                # d = 48
    assert m2._kv_flag(_cache(), 8, False) == ("kv", 0)    # head dims the decode kernels are not built for


def test_cache_constructor_and_refill_argument_rules():
    """Checks that fail before any kernel is touched (so they can run without a GPU)."""
    from vlm_bridge_b200 import VisionKVCache

    bridge = types.SimpleNamespace()
    with pytest.raises(RuntimeError, match=r"\[B, Nv, vision_dim\]"):
        VisionKVCache(bridge, torch.randn(5, 32))
    with pytest.raises(ValueError, match="precision"):
        VisionKVCache(bridge, torch.randn(1, 5, 32), precision="fp16")
    c = _cache(batch=2)
    c._bridge, c.precision = bridge, "bf16"
    with pytest.raises(RuntimeError, match="refill needs"):
        c.refill(torch.randn(3, 5, 32))                    # another batch size than the cache was built for
    with pytest.raises(RuntimeError, match="refill needs"):
        c.refill(torch.randn(2, 6, 32))                    # another number of vision tokens


def test_greedy_decode_argument_rules():
    from vlm_bridge_b200 import greedy_decode

    bridge = types.SimpleNamespace(training=False, eval=lambda: None, train=lambda mode=True: None)
    v = torch.randn(1, 5, 32)
    with pytest.raises(ValueError, match="precision"):
        greedy_decode(bridge, v, None, None, bos_token_id=2, precision="tf32")
    with pytest.raises(RuntimeError, match="use_cache must stay True"):
        greedy_decode(bridge, v, None, None, bos_token_id=2, precision="fp32", use_cache=False)
    c = _cache(batch=1)
    c.precision = "bf16"
    with pytest.raises(RuntimeError, match="holds bf16 K/V"):
        greedy_decode(bridge, v, None, None, bos_token_id=2, precision="fp32", kv_cache=c)
