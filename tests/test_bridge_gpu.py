"""Parity of the CUDA bridge (through the C ABI) against the CPU oracle -- the first gate.

Tolerances (stated, per BASELINE.json north_star "within a stated bf16 tolerance"):
  * outputs:   max|cuda - oracle_fp32| / max|oracle| <= 2e-2   (reference's own bf16-autocast
               deviation from fp32 on the same inputs is 2.9e-3, tests/golden/full_fingerprint.json)
  * gradients: per tensor, max|cuda - oracle| / max|oracle| <= 2e-2 (same definition as outputs, SURVEY.md
               section 8c) and Frobenius relative error <= 3e-2, with a floor for the key-bias gradients
               that are identically zero in exact arithmetic
  * loss:      |loss_cuda - loss_oracle| <= 1e-3
"""
import json
import os

import numpy as np
import pytest
import torch

from oracle import bridge_oracle as O

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _maxrel(a, b):
    return float((a.float().cpu() - b).abs().max() / b.abs().max())


def _frorel(a, b, floor):
    return float((a.float().cpu() - b).norm() / b.norm().clamp_min(floor))


def _grad_err(a, b, floor):
    """(max-rel, frobenius-rel) of gradient `a` against oracle `b`; floor is an absolute scale"""
    d = a.float().cpu() - b
    return (float(d.abs().max() / b.abs().max().clamp_min(floor / max(1.0, b.numel() ** 0.5))),
            float(d.norm() / b.norm().clamp_min(floor)))


def _floor(name, ref_norm_of):
    """Error floor for gradient `name`. The key-projection biases have an exactly-zero gradient in
    exact arithmetic (softmax is invariant to a per-row shift of the scores), so the reference holds
    fp32 rounding noise there and the bf16 path holds bf16 rounding noise of dK; compare those against
    the scale of the matching key-weight gradient instead of against ~0."""
    if name.endswith("w_k.bias"):
        return ref_norm_of(name[:-len("bias")] + "weight")
    return 1e-6


def _make(cfg, sd, dropout=0.0):
    from vlm_bridge_b200 import BridgeLite

    m = BridgeLite(dropout=dropout, **cfg)
    m.load_state_dict(sd, strict=True)
    return m.cuda()


def _check_fwd_bwd(cfg, sd, vision, text, d_out=None, tol=2e-2):
    kw = dict(num_blocks=cfg["num_blocks"], heads_cross=cfg["num_heads_cross"], heads_self=cfg["num_heads_self"])
    y_ref, loss_ref, dtext_ref, g_ref = O.bridge_loss_and_grads(sd, vision, text, d_out=d_out, **kw)
    m = _make(cfg, sd).eval()
    t = text.cuda().requires_grad_()
    y = m(vision.cuda(), t)
    assert y.dtype == torch.float32 and y.shape == text.shape
    loss = y.float().square().mean()
    if d_out is None:
        loss.backward()
    else:
        y.backward(d_out.cuda())
    assert _maxrel(y.detach(), y_ref) <= tol
    assert abs(float(loss) - loss_ref) <= 1e-3 * max(1.0, abs(loss_ref))
    assert _frorel(t.grad, dtext_ref, 1e-2 * float(dtext_ref.norm())) <= tol
    worst = {}
    for n, p in m.named_parameters():
        assert p.grad is not None and p.grad.dtype == torch.float32 and p.grad.shape == p.shape, n
        worst[n] = _grad_err(p.grad, g_ref[n], _floor(n, lambda k: float(g_ref[k].norm())))
    bad = {k: v for k, v in worst.items() if v[0] > tol or v[1] > 1.5 * tol}
    assert not bad, bad
    return m


def test_tiny_golden_fixture():
    z = np.load(os.path.join(HERE, "golden", "tiny_bridge.npz"))
    cfg = json.loads(bytes(z["cfg_json"]).decode())
    cfg.pop("dropout")
    sd = {k[len("param/"):]: torch.from_numpy(z[k]) for k in z.files if k.startswith("param/")}
    m = _make(cfg, sd).eval()
    t = torch.from_numpy(z["text"]).cuda().requires_grad_()
    y = m(torch.from_numpy(z["vision"]).cuda(), t)
    y.backward(torch.from_numpy(z["d_out"]).cuda())
    assert _maxrel(y.detach(), torch.from_numpy(z["y"])) <= 2e-2
    assert _frorel(t.grad, torch.from_numpy(z["d_text"]), 1e-3) <= 2e-2
    for n, p in m.named_parameters():
        floor = _floor(n, lambda k: float(np.linalg.norm(z["grad/" + k])))
        mx, fro = _grad_err(p.grad, torch.from_numpy(z["grad/" + n]), floor)
        assert mx <= 2e-2 and fro <= 3e-2, (n, mx, fro)


def test_c1_full_dims_forward_backward():
    cfg = dict(vision_dim=1024, language_dim=2304, num_blocks=2, num_heads_cross=8, num_heads_self=18)
    sd = O.init_state_dict(0)
    g = torch.Generator().manual_seed(1234)
    vision = torch.randn(2, 257, 1024, generator=g)
    text = torch.randn(2, 64, 2304, generator=g)
    m = _check_fwd_bwd(cfg, sd, vision, text)
    # golden fingerprint of the reference itself (not only the oracle)
    with open(os.path.join(HERE, "golden", "full_fingerprint.json")) as f:
        fp = json.load(f)
    with torch.no_grad():
        y = m(vision.cuda(), text.cuda())
    assert abs(float(y.std()) - fp["y_std"]) < 5e-3
    assert torch.allclose(y[0, 0, :8].cpu(), torch.tensor(fp["y_samples"]["[0,0,:8]"]), atol=3e-2)


def test_ragged_shapes_and_nontrivial_affine():
    """odd batch / lengths (row tails in every kernel) and random biases + LayerNorm affine"""
    cfg = dict(vision_dim=1024, language_dim=2304, num_blocks=2, num_heads_cross=8, num_heads_self=18)
    sd = O.init_state_dict(3)
    g = torch.Generator().manual_seed(77)
    for k in sd:
        if k.endswith("bias"):
            sd[k] = torch.randn(sd[k].shape, generator=g) * 0.05
        elif "ln_" in k:
            sd[k] = 1.0 + torch.randn(sd[k].shape, generator=g) * 0.1
    vision = torch.randn(3, 50, 1024, generator=g)
    text = torch.randn(3, 7, 2304, generator=g) * 2.0
    d_out = torch.randn(3, 7, 2304, generator=g)
    _check_fwd_bwd(cfg, sd, vision, text, d_out=d_out)


def test_init_state_dict_and_checkpoint_contract(tmp_path):
    from vlm_bridge_b200 import BridgeLite

    torch.manual_seed(0)
    m = BridgeLite()
    sd_ref = O.init_state_dict(0)
    sd = m.state_dict()
    assert list(sd.keys()) == list(sd_ref.keys()) and len(sd) == 52
    for k in sd:
        assert torch.equal(sd[k], sd_ref[k]), k          # same RNG stream as the reference constructor
    assert [n for n, _ in m.named_parameters()] == O.param_names(2)
    info = m.get_model_info()
    assert info["total_parameters"] == 158160384 and info["architecture"] == "Bridge-Lite"
    m = m.cuda()
    with torch.no_grad():
        m(torch.randn(1, 257, 1024).cuda(), torch.randn(1, 5, 2304).cuda())
    # format A (full_model.py:450-461) and format B (training_orchestrator.py:114-136) round trips
    torch.save({"bridge_module_state_dict": m.state_dict()}, tmp_path / "a.pth")
    torch.save({"model_state_dict": {"bridge_module." + k: v for k, v in m.state_dict().items()}}, tmp_path / "b.pth")
    m2 = BridgeLite().cuda()
    m2.load_state_dict(torch.load(tmp_path / "a.pth")["bridge_module_state_dict"], strict=True)
    b = torch.load(tmp_path / "b.pth")["model_state_dict"]
    m2.load_state_dict({k[len("bridge_module."):]: v for k, v in b.items()}, strict=True)
    for (n1, p1), (n2, p2) in zip(m.named_parameters(), m2.named_parameters()):
        assert n1 == n2 and torch.equal(p1, p2) and p1.dtype == torch.float32


def test_native_library_is_what_runs():
    from vlm_bridge_b200 import BridgeLite, _lib

    m = BridgeLite(vision_dim=32, language_dim=64, num_heads_cross=1, num_heads_self=1).cuda().eval()
    before = _lib.launch_count()
    with torch.no_grad():
        m(torch.randn(1, 4, 32).cuda(), torch.randn(1, 3, 64).cuda())
    assert _lib.launch_count() - before >= 20
    with pytest.raises(RuntimeError):
        BridgeLite(vision_dim=32, language_dim=64, num_heads_cross=1, num_heads_self=1)(torch.randn(1, 4, 32),
                                                                                     torch.randn(1, 3, 64))


def test_c5_vision_length_and_single_token():
    """high-resolution vision length (BASELINE config 5: 1370 patch tokens) and a one-token prefix
    (rows < one MMA tile everywhere) through forward + backward"""
    cfg = dict(vision_dim=1024, language_dim=2304, num_blocks=2, num_heads_cross=8, num_heads_self=18)
    sd = O.init_state_dict(5)
    g = torch.Generator().manual_seed(55)
    _check_fwd_bwd(cfg, sd, torch.randn(1, 1370, 1024, generator=g), torch.randn(1, 128, 2304, generator=g))
    _check_fwd_bwd(cfg, sd, torch.randn(2, 257, 1024, generator=g), torch.randn(2, 3, 2304, generator=g))
    # one token: self-attention over a single key has identically zero dQ / dK, so only the forward and the
    # input gradient are compared (the weight gradients of w_q / w_k are rounding noise around 0 on both sides)
    vision, text = torch.randn(2, 257, 1024, generator=g), torch.randn(2, 1, 2304, generator=g)
    y_ref, loss_ref, dtext_ref, g_ref = O.bridge_loss_and_grads(sd, vision, text)
    m = _make(cfg, sd).eval()
    t = text.cuda().requires_grad_()
    y = m(vision.cuda(), t)
    y.float().square().mean().backward()
    assert _maxrel(y.detach(), y_ref) <= 2e-2
    assert _frorel(t.grad, dtext_ref, 1e-2 * float(dtext_ref.norm())) <= 2e-2
    assert all(torch.isfinite(p.grad).all() for p in m.parameters())
    w = "bridge_blocks.1.ffn.3.weight"
    assert _grad_err(dict(m.named_parameters())[w].grad, g_ref[w], 1e-6)[1] <= 3e-2


def test_training_mode_dropout_is_unbiased_and_reproducible_in_backward():
    """dropout (p = 0.1, train mode): the output averaged over many masks approaches the eval output, and
    the backward regenerates the forward's masks (a gradient computed twice from one forward's state,
    via retain_graph, is identical; two forwards draw different masks)."""
    from vlm_bridge_b200 import BridgeLite

    cfg = dict(vision_dim=64, language_dim=128, num_blocks=2, num_heads_cross=2, num_heads_self=1)
    torch.manual_seed(3)
    m = BridgeLite(dropout=0.1, **cfg).cuda()
    g = torch.Generator().manual_seed(9)
    vision = torch.randn(4, 33, 64, generator=g).cuda()
    text = torch.randn(4, 24, 128, generator=g).cuda()
    m.eval()
    with torch.no_grad():
        y_eval = m(vision, text)
    m.train()
    with torch.no_grad():
        ys = torch.stack([m(vision, text) for _ in range(200)])
    assert not torch.equal(ys[0], ys[1])
    bias = float((ys.mean(0) - y_eval).abs().max() / y_eval.abs().max())
    spread = float(ys.std(0).max() / y_eval.abs().max())
    assert bias < 0.05 and spread > 1e-3, (bias, spread)
    t = text.clone().requires_grad_()
    y = m(vision, t)
    (g1,) = torch.autograd.grad(y.square().mean(), t, retain_graph=True)
    (g2,) = torch.autograd.grad(y.square().mean(), t)
    assert torch.equal(g1, g2)


def test_dropout_statistics_at_real_widths():
    """The d = 288 / d = 128 Philox paths at the real widths (1024 / 2304, heads 8 / 18, F = 9216): with p = 0.1
    in train mode the mean over masks approaches the eval output, masks differ between calls, and the backward
    regenerates the forward's masks (gradient from one forward computed twice is identical, and the gradient
    averaged over masks approaches the eval gradient)."""
    from vlm_bridge_b200 import BridgeLite

    sd = O.init_state_dict(4)
    cfg = dict(vision_dim=1024, language_dim=2304, num_blocks=2, num_heads_cross=8, num_heads_self=18)
    m = BridgeLite(dropout=0.1, **cfg)
    m.load_state_dict(sd, strict=True)
    m = m.cuda()
    g = torch.Generator().manual_seed(91)
    vision = torch.randn(2, 257, 1024, generator=g).cuda()
    text = torch.randn(2, 40, 2304, generator=g).cuda()
    m.eval()
    t0 = text.clone().requires_grad_()
    y_eval = m(vision, t0)
    (g_eval,) = torch.autograd.grad(y_eval.float().square().mean(), t0)
    y_eval = y_eval.detach()
    m.train()
    n = 64
    acc_y, acc_g, acc_y2 = torch.zeros_like(y_eval), torch.zeros_like(g_eval), torch.zeros_like(y_eval)
    first = None
    for i in range(n):
        t = text.clone().requires_grad_()
        y = m(vision, t)
        (gt,) = torch.autograd.grad(y.float().square().mean(), t, retain_graph=(i == 0))
        if i == 0:
            (gt2,) = torch.autograd.grad(y.float().square().mean(), t)
            assert torch.equal(gt, gt2)                # masks regenerated identically in the backward
            first = y.detach().clone()
        elif i == 1:
            assert not torch.equal(first, y.detach())  # a new forward draws new masks
        acc_y += y.detach(); acc_y2 += y.detach() ** 2; acc_g += gt
    mean_y, mean_g = acc_y / n, acc_g / n
    std_y = (acc_y2 / n - mean_y ** 2).clamp_min(0).sqrt()
    # the mean of n masked outputs deviates from the eval output by ~ std / sqrt(n) plus the second-order
    # effect of dropout inside the nonlinearities; 6 standard errors + 3 % of the output scale bounds both
    bias = (mean_y - y_eval).abs()
    assert float(std_y.mean()) > 1e-2 * float(y_eval.abs().mean())           # dropout really perturbs
    assert bool((bias <= 6 * std_y / n ** 0.5 + 0.03 * y_eval.abs().max()).all())
    assert float((mean_g - g_eval).norm() / g_eval.norm()) < 0.2


def test_debug_forward_prints_reference_statistics(capsys):
    """forward(..., debug=True) (bridge_module.py:427-454; called by every validation sample generation,
    core_training_loop.py:317): same return value as debug=False, the reference's per-block lines, and the
    NaN warning when a block output is not finite."""
    cfg = dict(vision_dim=1024, language_dim=2304, num_blocks=2, num_heads_cross=8, num_heads_self=18)
    sd = O.init_state_dict(6)
    m = _make(cfg, sd).eval()
    g = torch.Generator().manual_seed(66)
    vision, text = torch.randn(2, 33, 1024, generator=g).cuda(), torch.randn(2, 9, 2304, generator=g).cuda()
    with torch.no_grad():
        y = m(vision, text)
        capsys.readouterr()
        y_dbg = m(vision, text, debug=True)
    out = capsys.readouterr().out
    assert torch.equal(y, y_dbg)
    assert "Bridge Input - Vision: torch.Size([2, 33, 1024]), Text: torch.Size([2, 9, 2304])" in out
    assert "Text stats: mean=" in out and "Vision stats: mean=" in out
    assert "Block 1:" in out and "Block 2:" in out and "→" in out and "±" in out
    assert "NaN" not in out and "Inf" not in out
    # statistics printed for block 2 are those of the returned tensor
    assert f"{float(y.mean()):.4f}±{float(y.std()):.4f}" in out
    # autograd still works through the debug path (the reference's debug forward is an ordinary forward)
    t = text.clone().requires_grad_()
    yg = m(vision, t, debug=True)
    yg.float().square().mean().backward()
    assert t.grad is not None and torch.isfinite(t.grad).all()
    capsys.readouterr()
    bad = text.clone()
    bad[0, 0, 0] = float("nan")
    with torch.no_grad():
        m(vision, bad, debug=True)
    assert "NaN detected in Block 1 output" in capsys.readouterr().out


def test_vision_features_requiring_grad_are_rejected():
    cfg = dict(vision_dim=64, language_dim=128, num_blocks=2, num_heads_cross=2, num_heads_self=1)
    from vlm_bridge_b200 import BridgeLite

    m = BridgeLite(dropout=0.0, **cfg).cuda()
    with pytest.raises(RuntimeError, match="vision_features must not require grad"):
        m(torch.randn(1, 5, 64).cuda().requires_grad_(), torch.randn(1, 3, 128).cuda())


def test_invalidate_weight_cache_after_data_writes():
    """Writes through `.data` do not bump the version counter; `invalidate_weight_cache()` is the documented
    remedy."""
    from vlm_bridge_b200 import BridgeLite

    cfg = dict(vision_dim=64, language_dim=128, num_blocks=2, num_heads_cross=2, num_heads_self=1)
    torch.manual_seed(1)
    m = BridgeLite(dropout=0.0, **cfg).cuda().eval()
    v, t = torch.randn(1, 5, 64).cuda(), torch.randn(1, 3, 128).cuda()
    with torch.no_grad():
        y0 = m(v, t)
        m.bridge_blocks[1].ffn[3].weight.data.mul_(0.0)
        m.invalidate_weight_cache()
        y1 = m(v, t)
    assert not torch.equal(y0, y1)


def test_layer_classes_are_working_modules():
    """The reference exports BridgeBlock / MultiHeadCrossAttention / MultiHeadSelfAttention as modules of their own
    (model_architecture/__init__.py:22-27). Here their forwards run the library's kernels (inference only) and a
    stack of two BridgeBlock forwards reproduces BridgeLite.forward and the oracle."""
    from vlm_bridge_b200 import BridgeLite

    cfg = dict(vision_dim=1024, language_dim=2304, num_blocks=2, num_heads_cross=8, num_heads_self=18)
    sd = O.init_state_dict(8)
    m = _make(cfg, sd).eval()
    g = torch.Generator().manual_seed(88)
    vision, text = torch.randn(2, 257, 1024, generator=g), torch.randn(2, 70, 2304, generator=g)
    y_ref, blocks_ref = O.bridge_forward(sd, vision, text, return_blocks=True)
    with torch.no_grad():
        y = m(vision.cuda(), text.cuda())
        x = text.cuda()
        for blk in m.bridge_blocks:
            x = blk(x, vision.cuda())
        b0 = m.bridge_blocks[0]
        xn = torch.nn.functional.layer_norm(text, (2304,), sd["bridge_blocks.0.ln_cross.weight"],
                                            sd["bridge_blocks.0.ln_cross.bias"])
        ca = b0.cross_attention(xn.cuda(), vision.cuda(), vision.cuda())
        sa = b0.self_attention(xn.cuda())
    assert _maxrel(x, y_ref) <= 2e-2 and _maxrel(x, y.cpu()) <= 1e-2
    pre = "bridge_blocks.0.cross_attention."
    k = O.linear(vision, sd[pre + "w_k.weight"], sd[pre + "w_k.bias"], False)
    v = O.linear(vision, sd[pre + "w_v.weight"], sd[pre + "w_v.bias"], False)
    q = O.linear(xn, sd[pre + "w_q.weight"], sd[pre + "w_q.bias"], False)
    ca_ref = O.linear(O.attention(q, k, v, 8, False), sd[pre + "w_o.weight"], sd[pre + "w_o.bias"], False)
    assert _maxrel(ca, ca_ref) <= 2e-2
    pre = "bridge_blocks.0.self_attention."
    qs, ks, vs = (O.linear(xn, sd[pre + f"w_{n}.weight"], sd[pre + f"w_{n}.bias"], False) for n in "qkv")
    sa_ref = O.linear(O.attention(qs, ks, vs, 18, False), sd[pre + "w_o.weight"], sd[pre + "w_o.bias"], False)
    assert _maxrel(sa, sa_ref) <= 2e-2
    with pytest.raises(RuntimeError):
        b0.self_attention(xn.cuda(), mask=torch.ones(2, 70, 70).cuda())
    with pytest.raises(RuntimeError):
        b0(text.cuda().requires_grad_(), vision.cuda())
    with pytest.raises(RuntimeError):
        b0(text, vision)
