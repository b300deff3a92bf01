"""Caption decode at the size BASELINE.json names (config C4: batch 32, 64 new tokens from BOS, 257 vision
tokens, real widths 1024 / 2304, 2 blocks, heads 8 / 18), greedy, per-image cached vision K/V.

The reference decodes WITHOUT autocast (generate_caption runs the bridge in fp32, full_model.py:221-261), so
the yardstick is the fp32 CPU oracle (teacher-forced on the CUDA path's own prefixes: rows are independent and
the loop is deterministic, so "every argmax equals the oracle's argmax on the same prefix" is the same
statement as "the free-running captions are identical").

  * precision="fp32" (csrc/exact_fp32.cu, fp32 operands / products / sums): ids equal the oracle's on EVERY
    step of every row -- 32 x 64 = 2048 decisions, no exemption. The test also reports the oracle's smallest
    top-2 margin and checks it is far above fp32 noise, i.e. that the equality is not luck.
  * precision="bf16" (tensor-core kernels, the reference's autocast numerics): a bf16 logit error can flip an
    argmax whose margin is below that error. Reported: decided / total (steps whose oracle margin exceeds
    2e-2 * max|logit|, the bf16 bound) and agree; asserted: agree == decided on every row up to its first
    divergence from the fp32 ids, and every first divergence sits on an undecided step.

The frozen language model is outside the hot path; as in oracle.greedy_decode_bridge_only it is a fixed
embedding table and a fixed linear read-out.
"""
import json
import os

import pytest
import torch

from oracle import bridge_oracle as O

pytestmark = pytest.mark.gpu
B, NV, STEPS, V = 32, 257, 64, 1024
FP32_NOISE = 1e-4     # relative to max|logit|: two orders above the fp32 CUDA-vs-CPU logit difference (asserted)
BF16_BOUND = 2e-2     # relative to max|logit|: the stated bf16 tolerance of the bridge output


def _setup():
    from vlm_bridge_b200 import BridgeLite

    sd = O.init_state_dict(1)
    g = torch.Generator().manual_seed(20)
    vision = torch.randn(B, NV, 1024, generator=g)
    embed = torch.randn(V, 2304, generator=g)
    head = torch.randn(V, 2304, generator=g) / 48.0
    m = BridgeLite(dropout=0.0)
    m.load_state_dict(sd, strict=True)
    return sd, vision, embed, head, m.cuda().eval()


def _decode(m, vision, embed, head, precision, **kw):
    from vlm_bridge_b200 import greedy_decode

    e, h = embed.cuda(), head.cuda()
    ids, lengths = greedy_decode(m, vision.cuda(), lambda t: e[t], lambda y: y[:, -1, :] @ h.t(), bos_token_id=2,
                                 eos_token_id=1, max_new_tokens=STEPS, precision=precision, **kw)
    return ids.cpu(), lengths.cpu()


def _oracle_logits(sd, vision, embed, head, ids):
    """fp32 oracle logits of the last position for every prefix ids[:, :s+1], s = 0..STEPS-1 (the image K/V are
    computed once: identical arithmetic to re-projecting them every step as the reference does)."""
    kvs = O.vision_kv(sd, vision)
    out = []
    for s in range(STEPS):
        y = O.bridge_forward_cached(sd, kvs, embed[ids[:, :s + 1]])
        out.append(y[:, -1, :] @ head.t())
    return torch.stack(out, dim=1)          # [B, STEPS, V]


@pytest.mark.timeout(1500)
def test_c4_greedy_ids_fp32_exact_and_bf16_explained_by_margins():
    sd, vision, embed, head, m = _setup()
    ids32, len32 = _decode(m, vision, embed, head, "fp32")
    ids16, _ = _decode(m, vision, embed, head, "bf16")
    ids16_g, _ = _decode(m, vision, embed, head, "bf16", use_graphs=True)
    assert torch.equal(ids16, ids16_g)                       # graph replay == eager launches, at size
    assert ids32.shape == (B, STEPS + 1) and bool((ids32[:, 0] == 2).all())

    logits = _oracle_logits(sd, vision, embed, head, ids32)                  # teacher-forced on the fp32 ids
    want = logits.argmax(-1)                                                 # [B, STEPS]
    top2 = logits.topk(2, dim=-1).values
    margin = (top2[..., 0] - top2[..., 1]) / logits.abs().amax(-1)           # relative top-2 margin
    # --- fp32 path: every one of the B * STEPS decisions equals the oracle's ---
    agree32 = int((want == ids32[:, 1:]).sum())
    report = {"decisions": B * STEPS, "fp32_agree": agree32, "min_rel_margin": float(margin.min()),
              "decided_fp32_noise": int((margin > FP32_NOISE).sum())}
    assert agree32 == B * STEPS, report
    assert float(margin.min()) > FP32_NOISE, report          # the equality above has two orders of head-room
    # how far the fp32 CUDA logits really are from the oracle's (last step, all rows)
    from vlm_bridge_b200 import VisionKVCache

    with torch.no_grad():
        cache = VisionKVCache(m, vision.cuda(), precision="fp32")
        y = m(None, embed[ids32[:, :STEPS]].cuda(), kv_cache=cache)
    lg = (y[:, -1, :].cpu() @ head.t())
    report["fp32_logit_rel_err"] = float((lg - logits[:, -1]).abs().max() / logits[:, -1].abs().max())
    assert report["fp32_logit_rel_err"] < 1e-5, report
    # --- bf16 path: on-trajectory decisions agree wherever the margin exceeds the bf16 bound ---
    same_prefix = torch.ones(B, dtype=torch.bool)
    decided = agree = on_traj = 0
    first_div_margins = []
    for s in range(STEPS):
        clear = margin[:, s] > BF16_BOUND
        eq = ids16[:, s + 1] == want[:, s]
        on_traj += int(same_prefix.sum())
        decided += int((same_prefix & clear).sum())
        agree += int((same_prefix & clear & eq).sum())
        for b in torch.nonzero(same_prefix & ~eq).flatten().tolist():
            first_div_margins.append(float(margin[b, s]))
        same_prefix &= eq
    report.update(bf16_on_trajectory=on_traj, bf16_decided=decided, bf16_agree=agree,
                  bf16_rows_identical_to_fp32=int(same_prefix.sum()), bf16_first_divergence_margins=first_div_margins)
    out_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(out_dir):
        with open(os.path.join(out_dir, "c4_decode_parity.json"), "w") as f:
            json.dump(report, f, indent=1)
    print("C4 decode parity:", json.dumps(report))
    assert agree == decided, report
    assert all(mg <= BF16_BOUND for mg in first_div_margins), report
    assert decided >= 0.7 * on_traj, report       # (2e-2 * max|logit| is a loose bound: ~20 % of the steps fall under it)
    # lengths: first EOS (token 1) per row, else the full length
    for b in range(B):
        row = ids32[b, 1:].tolist()
        assert int(len32[b]) == ((row.index(1) + 1) if 1 in row else STEPS + 1)


def test_fp32_forward_matches_oracle_to_fp32_rounding():
    """The fp32 path against the oracle on a ragged shape (rows and keys that are not tile multiples), with and
    without position rows: max|d| / max|ref| <= 2e-5."""
    from vlm_bridge_b200 import BridgeLite, VisionKVCache

    sd = O.init_state_dict(3)
    g = torch.Generator().manual_seed(31)
    vision = torch.randn(3, 70, 1024, generator=g)
    text = torch.randn(3, 11, 2304, generator=g)
    m = BridgeLite(dropout=0.0)
    m.load_state_dict(sd, strict=True)
    m = m.cuda().eval()
    y_ref = O.bridge_forward(sd, vision, text)
    with torch.no_grad():
        cache = VisionKVCache(m, vision.cuda(), precision="fp32")
        y = m(None, text.cuda(), kv_cache=cache)
        for j in range(1, 12):
            y_inc = m(None, text[:, :j].cuda(), kv_cache=cache, cached_positions=j - 1)
    assert float((y.cpu() - y_ref).abs().max() / y_ref.abs().max()) <= 2e-5
    assert float((y_inc.cpu() - y_ref).abs().max() / y_ref.abs().max()) <= 2e-5
    with pytest.raises(RuntimeError):                    # inference only
        m(None, text.cuda().requires_grad_(), kv_cache=cache)


def test_cache_refill_in_place_keeps_graphs_valid():
    """VisionKVCache.refill(): new images, same buffers -> graphs captured over the cache are replayed for the
    new images and give the ids of a freshly built cache."""
    from vlm_bridge_b200 import BridgeLite, VisionKVCache, greedy_decode
    from vlm_bridge_b200.decode import DecodeStepGraphs

    sd = O.init_state_dict(1)
    g = torch.Generator().manual_seed(41)
    v1, v2 = torch.randn(4, 257, 1024, generator=g).cuda(), torch.randn(4, 257, 1024, generator=g).cuda()
    embed = torch.randn(512, 2304, generator=g).cuda()
    head = (torch.randn(512, 2304, generator=g) / 48.0).cuda()
    m = BridgeLite(dropout=0.0)
    m.load_state_dict(sd, strict=True)
    m = m.cuda().eval()
    kw = dict(bos_token_id=2, eos_token_id=1, max_new_tokens=8)
    e_fn, l_fn = (lambda t: embed[t]), (lambda y: y[:, -1, :] @ head.t())
    cache = VisionKVCache(m, v1)
    ptrs = (cache.kv.data_ptr(), cache.kv_packed.data_ptr(), cache.kv_tc.data_ptr(), cache.x1.data_ptr())
    graphs = DecodeStepGraphs(m, cache)
    a1, _ = greedy_decode(m, v1, e_fn, l_fn, kv_cache=cache, step_graphs=graphs, **kw)
    n_graphs = len(graphs._graphs)
    a2, _ = greedy_decode(m, v2, e_fn, l_fn, kv_cache=cache, step_graphs=graphs, refill_cache=True, **kw)
    assert len(graphs._graphs) == n_graphs                               # nothing was re-captured
    assert ptrs == (cache.kv.data_ptr(), cache.kv_packed.data_ptr(), cache.kv_tc.data_ptr(), cache.x1.data_ptr())
    b1, _ = greedy_decode(m, v1, e_fn, l_fn, **kw)
    b2, _ = greedy_decode(m, v2, e_fn, l_fn, **kw)
    assert torch.equal(a1, b1) and torch.equal(a2, b2)
    assert not torch.equal(a1, a2)
    with pytest.raises(RuntimeError):
        cache.refill(v1[:2])


def test_nan_and_inf_guards_are_per_row():
    """The reference guards one caption at a time (full_model.py:270-283): a NaN / Inf in one row's logits must
    not change another row's token."""
    from vlm_bridge_b200 import BridgeLite, greedy_decode

    sd = O.init_state_dict(1)
    g = torch.Generator().manual_seed(43)
    v = torch.randn(3, 33, 1024, generator=g).cuda()
    embed = torch.randn(64, 2304, generator=g).cuda()
    head = (torch.randn(64, 2304, generator=g) / 48.0).cuda()
    m = BridgeLite(dropout=0.0)
    m.load_state_dict(sd, strict=True)
    m = m.cuda().eval()

    def lm_clean(y):
        return y[:, -1, :] @ head.t()

    def lm_poisoned(y):
        lg = lm_clean(y).clone()
        lg[0, 5] = float("nan")          # row 0: NaN -> all zeros -> argmax 0
        lg[1, 7] = float("inf")          # row 1: clamped to [-100, 100] -> argmax 7
        return lg

    kw = dict(bos_token_id=2, max_new_tokens=1)
    clean, _ = greedy_decode(m, v, lambda t: embed[t], lm_clean, **kw)
    pois, _ = greedy_decode(m, v, lambda t: embed[t], lm_poisoned, **kw)
    assert int(pois[0, 1]) == 0 and int(pois[1, 1]) == 7
    assert int(pois[2, 1]) == int(clean[2, 1])
