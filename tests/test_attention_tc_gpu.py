"""The tcgen05 training attention (csrc/attention_train_tc.cu) against softmax(Q K^T / sqrt(d)) V in fp32 torch
(the arithmetic of F.scaled_dot_product_attention at bridge_module.py:132-139 / :230-237) and against the mma.sync
kernels it replaces, through the C ABI (b200b_attention_fwd / _bwd).

Tolerances: output and gradients max|d| / max|ref| <= 1e-2 (bf16 operands and probabilities), log-sum-exp within
1e-3; with dropout the two implementations draw the SAME Philox mask (same element indexing), so they are compared
with each other at the same tolerance, and the backward must regenerate the forward's mask.
"""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

# (B, H, HD, Lq, Lk): C2 cross, C2 self, C5 cross (1370 keys), two query blocks + ragged keys, short ragged, toy dims
SHAPES = [(2, 8, 288, 128, 257), (2, 18, 128, 128, 128), (1, 8, 288, 128, 1370), (1, 8, 288, 200, 300),
          (3, 8, 288, 65, 40), (2, 2, 64, 70, 100), (1, 1, 128, 129, 17), (2, 8, 288, 24, 33)]


def _ref(q, k, v, B, H, HD, Lq, Lk):
    qh = q.float().reshape(B, Lq, H, HD).transpose(1, 2)
    kh = k.float().reshape(B, Lk, H, HD).transpose(1, 2)
    vh = v.float().reshape(B, Lk, H, HD).transpose(1, 2)
    s = qh @ kh.transpose(-1, -2) / math.sqrt(HD)
    o = (torch.softmax(s, -1) @ vh).transpose(1, 2).reshape(B * Lq, H * HD)
    return o, torch.logsumexp(s, -1)


def _make(B, H, HD, Lq, Lk, seed):
    g = torch.Generator().manual_seed(seed)
    D = H * HD
    # Q inside a fused [T, 3D] projection output and K / V inside a [Tv, 4D] one: row pitches larger than the slices
    qkv = torch.randn(B * Lq, 3 * D, generator=g).bfloat16().cuda()
    kv = torch.randn(B * Lk, 4 * D, generator=g).bfloat16().cuda()
    return qkv[:, D:2 * D], kv[:, 2 * D:3 * D], kv[:, 3 * D:4 * D]


def _set_tc(mask):
    from vlm_bridge_b200 import _lib
    return _lib.lib().b200b_attention_set_tc(mask)


@pytest.mark.parametrize("B,H,HD,Lq,Lk", SHAPES)
def test_forward_matches_fp32_reference_and_legacy_kernel(B, H, HD, Lq, Lk):
    from vlm_bridge_b200 import ops

    q, k, v = _make(B, H, HD, Lq, Lk, seed=Lq * 7 + Lk)
    kw = dict(batch=B, heads=H, len_q=Lq, len_k=Lk, head_dim=HD)
    prev = _set_tc(3)
    try:
        # dropout_p > 0 would change the values; a tiny p keeps Lq <= 64 shapes off the decode kernel is not needed:
        # shapes with Lq <= 64 and no dropout take the decode kernel on both settings, which is fine to compare too
        o_tc, lse_tc = ops.attention_fwd(q, k, v, **kw)
        _set_tc(0)
        o_old, lse_old = ops.attention_fwd(q, k, v, **kw)
    finally:
        _set_tc(prev)
    ref, lse_ref = _ref(q, k, v, B, H, HD, Lq, Lk)
    scale = float(ref.abs().max())
    assert float((o_tc.float() - ref).abs().max()) <= 1e-2 * scale
    assert float((o_old.float() - ref).abs().max()) <= 1e-2 * scale
    assert float((lse_tc * math.log(2) - lse_ref).abs().max()) <= 1e-3
    assert float((lse_tc - lse_old).abs().max()) <= 1e-3


@pytest.mark.parametrize("B,H,HD,Lq,Lk", SHAPES[:5])
def test_forward_with_dropout_draws_the_same_mask_as_the_legacy_kernel(B, H, HD, Lq, Lk):
    from vlm_bridge_b200 import ops

    q, k, v = _make(B, H, HD, Lq, Lk, seed=Lq * 11 + Lk)
    kw = dict(batch=B, heads=H, len_q=Lq, len_k=Lk, head_dim=HD, dropout_p=0.1, seed=12345, dropout_stream=3)
    prev = _set_tc(3)
    try:
        o_tc, lse_tc = ops.attention_fwd(q, k, v, **kw)
        _set_tc(0)
        o_old, lse_old = ops.attention_fwd(q, k, v, **kw)
    finally:
        _set_tc(prev)
    ref, _ = _ref(q, k, v, B, H, HD, Lq, Lk)
    scale = float(ref.abs().max())
    assert float((o_tc.float() - o_old.float()).abs().max()) <= 1e-2 * scale      # same mask, same values
    assert float((o_tc.float() - ref).abs().max()) > 1e-2 * scale                 # and dropout really acted
    assert float((lse_tc - lse_old).abs().max()) <= 1e-3                          # the normaliser ignores the mask


def _bwd(q, k, v, o, lse, d_o, kw):
    from vlm_bridge_b200 import ops

    dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
    # gradients are written into strided views, as the bridge does (dQ / dK / dV live inside fused buffers)
    ops.attention_bwd(d_o, q, k, v, o, lse, dq, dk, dv, **kw)
    return dq.float(), dk.float(), dv.float()


# backward shapes: the tcgen05 path covers query blocks of up to 128 rows; (1, 8, 288, 200, 300) must fall back cleanly
# (6, 8, 288, 128, 1370): 11 key blocks x 48 (image, head) pairs exceed two waves of CTAs, so the key-major pass walks
# several key blocks per CTA (the C5 stress shape does the same at batch 16)
BWD_SHAPES = [(2, 8, 288, 128, 257), (2, 18, 128, 128, 128), (1, 8, 288, 128, 1370), (3, 8, 288, 65, 40),
              (2, 2, 64, 70, 100), (1, 1, 128, 100, 17), (2, 8, 288, 24, 33), (1, 8, 288, 200, 300),
              (6, 8, 288, 128, 1370)]


@pytest.mark.parametrize("B,H,HD,Lq,Lk", BWD_SHAPES)
def test_backward_matches_fp32_autograd_and_legacy_kernels(B, H, HD, Lq, Lk):
    from vlm_bridge_b200 import ops

    q, k, v = _make(B, H, HD, Lq, Lk, seed=Lq * 13 + Lk)
    g = torch.Generator().manual_seed(5)
    d_o = torch.randn(B * Lq, H * HD, generator=g).bfloat16().cuda()
    kw = dict(batch=B, heads=H, len_q=Lq, len_k=Lk, head_dim=HD)
    prev = _set_tc(3)
    try:
        o, lse = ops.attention_fwd(q, k, v, **kw)
        got = _bwd(q, k, v, o, lse, d_o, kw)
        _set_tc(0)
        o_old, lse_old = ops.attention_fwd(q, k, v, **kw)
        old = _bwd(q, k, v, o_old, lse_old, d_o, kw)
    finally:
        _set_tc(prev)
    qf, kf, vf = (t.float().clone().requires_grad_() for t in (q, k, v))
    ref, _ = _ref(qf, kf, vf, B, H, HD, Lq, Lk)
    ref.backward(d_o.float())
    for name, a, b_, r in zip(("dq", "dk", "dv"), got, old, (qf.grad, kf.grad, vf.grad)):
        scale = float(r.abs().max())
        assert float((a - r).abs().max()) <= 1.5e-2 * scale, name
        assert float((b_ - r).abs().max()) <= 1.5e-2 * scale, name + " (legacy)"


@pytest.mark.parametrize("B,H,HD,Lq,Lk", BWD_SHAPES[:5] + BWD_SHAPES[-1:])
def test_backward_with_dropout_regenerates_the_forward_mask(B, H, HD, Lq, Lk):
    """Same seed -> the tcgen05 and the mma.sync kernels draw the same mask in forward and backward, so all four
    combinations agree; a wrong mask in either backward pass would show as an O(1) relative error."""
    from vlm_bridge_b200 import ops

    q, k, v = _make(B, H, HD, Lq, Lk, seed=Lq * 17 + Lk)
    g = torch.Generator().manual_seed(6)
    d_o = torch.randn(B * Lq, H * HD, generator=g).bfloat16().cuda()
    kw = dict(batch=B, heads=H, len_q=Lq, len_k=Lk, head_dim=HD, dropout_p=0.1, seed=777, dropout_stream=1)
    prev = _set_tc(3)
    try:
        o, lse = ops.attention_fwd(q, k, v, **kw)
        got = _bwd(q, k, v, o, lse, d_o, kw)
        _set_tc(0)
        old = _bwd(q, k, v, o, lse, d_o, kw)          # legacy backward on the tcgen05 forward's output
    finally:
        _set_tc(prev)
    for name, a, b_ in zip(("dq", "dk", "dv"), got, old):
        scale = float(b_.abs().max())
        assert float((a - b_).abs().max()) <= 1.5e-2 * scale, name
