"""Fused clip + AdamW step against torch: clip_grad_norm_ + torch.optim.AdamW on
a copy of the same parameters and gradients. fp32 elementwise arithmetic with the same formulas:
parameters must agree to 2e-6 relative (fma contraction only); the global gradient norm is a
158 M-term fp32 sum taken in a different order than torch's (per-tensor norms, then the norm of
those), so it -- and through the clip coefficient the moments -- agree to 1e-5 / 5e-5 relative."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu
SMALL = dict(vision_dim=64, language_dim=128, num_blocks=2, num_heads_cross=2, num_heads_self=1)


def _rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-12))


def _run(cfg, B, L, Nv, steps, max_norm, use_scaler):
    from vlm_bridge_b200 import BridgeAdamW, BridgeLite

    torch.manual_seed(0)
    m = BridgeLite(dropout=0.0, **cfg).cuda().train()
    g = torch.Generator().manual_seed(5)
    vision = torch.randn(B, Nv, cfg["vision_dim"], generator=g).cuda()
    text = torch.randn(B, L, cfg["language_dim"], generator=g).cuda()
    with torch.no_grad():
        m(vision, text)                                   # flatten
    ref_params = [p.detach().clone().requires_grad_() for p in m.parameters()]
    ref_opt = torch.optim.AdamW(ref_params, lr=3e-4, weight_decay=0.01, betas=(0.9, 0.999), eps=1e-8)
    opt = BridgeAdamW(m, lr=3e-4, weight_decay=0.01, max_grad_norm=max_norm)
    scaler = torch.amp.GradScaler("cuda", init_scale=1024.0) if use_scaler else None
    for s in range(steps):
        opt.zero_grad(set_to_none=True)
        loss = m(vision, text).float().square().mean() * (1.0 + s)
        if scaler is not None:
            scaler.scale(loss).backward()
        else:
            loss.backward()
        scale = float(scaler.get_scale()) if scaler is not None else 1.0
        for rp, p in zip(ref_params, m.parameters()):
            rp.grad = p.grad.detach().clone() / scale
        want_norm = torch.nn.utils.clip_grad_norm_(ref_params, max_norm) if max_norm else None
        ref_opt.step()
        if scaler is not None:
            scaler.step(opt)
            scaler.update()
        else:
            opt.step()
        if want_norm is not None:
            assert abs(float(opt.last_grad_norm) - float(want_norm)) <= 1e-5 * float(want_norm)
        for (n, p), rp in zip(m.named_parameters(), ref_params):
            assert _rel(p.detach(), rp.detach()) <= 2e-6, (s, n)
        # the bf16 operand copy is current: the next forward must not need (or do) a re-cast
        assert m._w16_key == tuple(p._version for p in m.parameters())
        w16 = m._w16[:m._layout.n_weights].float()
        assert torch.equal(w16, m._flat[:m._layout.n_weights].bfloat16().float())
    return m, opt, ref_opt, ref_params


def test_fused_adamw_matches_torch_small_clip():
    _run(SMALL, B=2, L=5, Nv=7, steps=4, max_norm=0.3, use_scaler=False)


def test_fused_adamw_with_gradscaler_and_no_clip():
    _run(SMALL, B=2, L=5, Nv=7, steps=3, max_norm=None, use_scaler=True)


def test_fused_adamw_full_dims_and_state_dict_round_trip():
    cfg = dict(vision_dim=1024, language_dim=2304, num_blocks=2, num_heads_cross=8, num_heads_self=18)
    m, opt, ref_opt, ref_params = _run(cfg, B=1, L=8, Nv=17, steps=2, max_norm=0.3, use_scaler=False)
    sd, ref_sd = opt.state_dict(), ref_opt.state_dict()
    assert sorted(sd["state"].keys()) == sorted(ref_sd["state"].keys()) == list(range(52))
    for i in range(52):
        assert set(sd["state"][i].keys()) == {"step", "exp_avg", "exp_avg_sq"}
        assert float(sd["state"][i]["step"]) == float(ref_sd["state"][i]["step"]) == 2.0
        assert _rel(sd["state"][i]["exp_avg"], ref_sd["state"][i]["exp_avg"]) <= 2.5e-5
        assert _rel(sd["state"][i]["exp_avg_sq"], ref_sd["state"][i]["exp_avg_sq"]) <= 5e-5
    # our state loads into torch's AdamW and torch's into ours (checkpoint contract)
    ref2 = torch.optim.AdamW([p.detach().clone().requires_grad_() for p in ref_params], lr=3e-4)
    ref2.load_state_dict(copy.deepcopy(sd))
    from vlm_bridge_b200 import BridgeAdamW

    opt2 = BridgeAdamW(m, lr=3e-4, weight_decay=0.01, max_grad_norm=0.3)
    opt2.load_state_dict(copy.deepcopy(ref_sd))
    opt2._ensure_state()
    assert opt2._steps == 2
    assert _rel(opt2._m, opt._m) <= 2.5e-5 and _rel(opt2._v, opt._v) <= 5e-5


def test_nonfinite_gradients_skip_the_step():
    from vlm_bridge_b200 import BridgeAdamW, BridgeLite

    torch.manual_seed(0)
    m = BridgeLite(dropout=0.0, **SMALL).cuda().train()
    opt = BridgeAdamW(m, lr=1e-2, max_grad_norm=1.0)
    m(torch.randn(1, 3, 64).cuda(), torch.randn(1, 2, 128).cuda()).sum().backward()
    before = m._flat.clone()
    next(m.parameters()).grad[0, 0] = float("inf")
    opt.step()
    assert torch.equal(m._flat, before)
    # a skipped step does not advance the step count (torch's fused AdamW under GradScaler behaves the same;
    # the next step's bias corrections depend on it) and is counted as skipped
    assert opt.skipped_steps == 1
    assert float(opt.state_dict()["state"][0]["step"]) == 0.0
    ref_params = [p.detach().clone().requires_grad_() for p in m.parameters()]
    ref_opt = torch.optim.AdamW(ref_params, lr=1e-2, weight_decay=1e-2)
    opt.zero_grad(set_to_none=True)
    m(torch.randn(1, 3, 64).cuda(), torch.randn(1, 2, 128).cuda()).sum().backward()
    for rp, p in zip(ref_params, m.parameters()):
        rp.grad = p.grad.detach().clone()
    torch.nn.utils.clip_grad_norm_(ref_params, 1.0)
    ref_opt.step()
    opt.step()
    assert opt.skipped_steps == 1 and float(opt.state_dict()["state"][0]["step"]) == 1.0
    for p, rp in zip(m.parameters(), ref_params):       # first applied step uses the step-1 bias corrections
        assert _rel(p.detach(), rp.detach()) <= 2e-6


def test_loss_trajectory_100_steps_matches_reference_numerics():
    """BASELINE.json north_star: "loss within 1e-3 over 100 steps". The CUDA training loop (bridge fwd+bwd
    kernels + fused clip/AdamW) against the CPU oracle training with torch.optim.AdamW +
    clip_grad_norm_ (core_training_loop.py:84-104; clip 0.3 = config/training-default.yaml:9) on the same
    batches from the same initial weights, 100 steps. Compared with the oracle in the reference's
    training numerics (bf16-rounded GEMM operands, as torch.autocast gives) at
    |d loss| <= 1e-3 * max(1, loss) per step, and with the pure-fp32 oracle at twice that (the fp32 and
    bf16-emulating oracles themselves differ by up to 5e-4 on this run)."""
    from oracle import bridge_oracle as O
    from vlm_bridge_b200 import BridgeAdamW, BridgeLite

    cfg = dict(vision_dim=64, language_dim=128, num_blocks=2, num_heads_cross=2, num_heads_self=1)
    steps, lr, wd, clip = 100, 1e-4, 0.01, 0.3
    g = torch.Generator().manual_seed(5)
    batches = [(torch.randn(4, 17, 64, generator=g), torch.randn(4, 12, 128, generator=g)) for _ in range(4)]
    sd0 = O.init_state_dict(0, vision_dim=64, language_dim=128, num_blocks=2)

    def oracle_run(emulate_bf16):
        params = {k: v.clone().requires_grad_() for k, v in sd0.items()}
        opt = torch.optim.AdamW(list(params.values()), lr=lr, weight_decay=wd)
        out = []
        for s in range(steps):
            v, t = batches[s % 4]
            opt.zero_grad()
            loss = O.bridge_forward(params, v, t, num_blocks=2, heads_cross=2, heads_self=1,
                                    emulate_bf16=emulate_bf16).square().mean()
            loss.backward()
            torch.nn.utils.clip_grad_norm_(list(params.values()), clip)
            opt.step()
            out.append(float(loss.detach()))
        return torch.tensor(out)

    m = BridgeLite(dropout=0.0, **cfg)
    m.load_state_dict(sd0, strict=True)
    m = m.cuda().train()
    opt = BridgeAdamW(m, lr=lr, weight_decay=wd, max_grad_norm=clip)
    dev_batches = [(v.cuda(), t.cuda()) for v, t in batches]
    losses = []
    for s in range(steps):
        v, t = dev_batches[s % 4]
        opt.zero_grad(set_to_none=True)
        loss = m(v, t).float().square().mean()
        loss.backward()
        opt.step()
        losses.append(loss.detach())
    got = torch.stack(losses).cpu()
    ref16, ref32 = oracle_run(True), oracle_run(False)
    assert float(ref32[-1]) < 0.5 * float(ref32[0])        # the run really trains
    dev16 = ((got - ref16).abs() / ref16.clamp_min(1.0)).max()
    dev32 = ((got - ref32).abs() / ref32.clamp_min(1.0)).max()
    assert float(dev16) <= 1e-3, (float(dev16), float(dev32))
    assert float(dev32) <= 2e-3, (float(dev16), float(dev32))
