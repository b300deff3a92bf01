"""Parity at BASELINE.json's full sizes.

C2 (batch 8, L=128, Nv=257) and C5 (batch 16, L=128, Nv=1370) forward + backward directly against the
CPU oracle (it finishes them in seconds), the embedding-scale variants of SURVEY.md 8d (x48 / x0.02:
the two places the Gemma sqrt(2304) scale can live), and size-independent properties at C2: the batch
is a set of independent samples (permutation equivariance, gradients of a batch = sum over its halves)
and the backward is linear in the upstream gradient.

Tolerances: as tests/test_bridge_gpu.py (2e-2 max-rel outputs / gradients, 3e-2 Frobenius). Properties:
permutation 1e-6 of the output maximum (same arithmetic per row, rows land in other tiles); linearity
rtol 1e-6 (scaling by 2 commutes with every rounding); halves vs whole 5e-3 Frobenius (bf16 rounding of
the weight-gradient operands is per element, only the fp32 summation order differs)."""
import pytest
import torch

from oracle import bridge_oracle as O

pytestmark = pytest.mark.gpu
CFG = dict(vision_dim=1024, language_dim=2304, num_blocks=2, num_heads_cross=8, num_heads_self=18)


def _model(sd, train=False):
    from vlm_bridge_b200 import BridgeLite

    m = BridgeLite(dropout=0.0, **CFG)
    m.load_state_dict(sd, strict=True)
    m = m.cuda()
    return m.train() if train else m.eval()


def _against_oracle(sd, vision, text, tol=2e-2, conditioning=False):
    """conditioning=True: gradient tensors that are ill-conditioned in bf16 -- those on which the oracle run
    with the reference's autocast rounding (bf16 GEMM operands) itself deviates from the fp32 oracle by more
    than `tol` -- are held to an absolute bound instead, 1e-3 of the global gradient norm. They arise with
    x0.02 embeddings: after block 0 every position carries nearly the same vector, the self-attention of
    block 1 is uniform to ~1e-3, and its dQ / dK are differences of nearly equal numbers (share of the total
    gradient norm: 4.5e-4). Any flash-style backward, the reference's SDPA included, takes
    delta = rowsum(dO * O) from the bf16-rounded O, which bounds the accuracy of such a gradient by
    ~2^-9 |delta| rather than by its own size."""
    y_ref, loss_ref, dtext_ref, g_ref = O.bridge_loss_and_grads(sd, vision, text)
    loose = set()
    if conditioning:
        g16 = O.bridge_loss_and_grads(sd, vision, text, emulate_bf16=True)[3]
        loose = {n for n in g_ref if float((g16[n] - g_ref[n]).norm() / g_ref[n].norm().clamp_min(1e-30)) > tol}
    g_total = float(torch.sqrt(sum(v.double().norm() ** 2 for v in g_ref.values())))
    m = _model(sd)
    t = text.cuda().requires_grad_()
    y = m(vision.cuda(), t)
    loss = y.float().square().mean()
    loss.backward()
    assert float((y.detach().cpu() - y_ref).abs().max() / y_ref.abs().max()) <= tol
    assert abs(float(loss) - loss_ref) <= 1e-3 * max(1.0, abs(loss_ref))
    assert float((t.grad.cpu() - dtext_ref).norm() / dtext_ref.norm()) <= tol
    bad = {}
    for n, p in m.named_parameters():
        ref = g_ref[n]
        floor = float(g_ref[n[:-len("bias")] + "weight"].norm()) if n.endswith("w_k.bias") else 1e-6
        d = p.grad.cpu() - ref
        if n in loose:
            if float(d.norm()) > 1e-3 * g_total:
                bad[n] = ("abs", float(d.norm()) / g_total)
            continue
        mx = float(d.abs().max() / ref.abs().max().clamp_min(floor / max(1.0, ref.numel() ** 0.5)))
        fro = float(d.norm() / ref.norm().clamp_min(floor))
        if mx > tol or fro > 1.5 * tol:
            bad[n] = (mx, fro)
    assert not bad, bad


def test_c2_full_size_against_oracle():
    g = torch.Generator().manual_seed(1234)
    _against_oracle(O.init_state_dict(0), torch.randn(8, 257, 1024, generator=g), torch.randn(8, 128, 2304, generator=g))


def test_c5_full_size_against_oracle():
    g = torch.Generator().manual_seed(1235)
    _against_oracle(O.init_state_dict(0), torch.randn(16, 1370, 1024, generator=g), torch.randn(16, 128, 2304, generator=g))


@pytest.mark.parametrize("scale", [48.0, 0.02])
def test_embedding_scale_variants(scale):
    g = torch.Generator().manual_seed(int(scale * 100))
    _against_oracle(O.init_state_dict(0), torch.randn(2, 257, 1024, generator=g),
                    torch.randn(2, 64, 2304, generator=g) * scale, conditioning=True)


def _grads(m, vision, text, d_out):
    for p in m.parameters():
        p.grad = None
    t = text.clone().requires_grad_()
    y = m(vision, t)
    y.backward(d_out)
    return y.detach(), t.grad, torch.cat([p.grad.reshape(-1) for p in m.parameters()])


def test_c2_size_independent_properties():
    g = torch.Generator().manual_seed(2024)
    B, L, Nv = 8, 128, 257
    vision = torch.randn(B, Nv, 1024, generator=g).cuda()
    text = torch.randn(B, L, 2304, generator=g).cuda()
    d_out = torch.randn(B, L, 2304, generator=g).cuda()
    m = _model(O.init_state_dict(4))
    y, dt, gp = _grads(m, vision, text, d_out)
    # 1. samples are independent: permuting the batch permutes outputs and input gradients, and leaves the
    #    parameter gradients (a sum over samples) unchanged up to summation order
    perm = torch.randperm(B, generator=g).cuda()
    y_p, dt_p, gp_p = _grads(m, vision[perm], text[perm], d_out[perm])
    assert float((y_p - y[perm]).abs().max() / y.abs().max()) <= 1e-6
    assert float((dt_p - dt[perm]).abs().max() / dt.abs().max()) <= 1e-6
    assert float((gp_p - gp).norm() / gp.norm()) <= 5e-3
    # 2. the backward is linear in the upstream gradient
    _, dt2, gp2 = _grads(m, vision, text, 2.0 * d_out)
    assert torch.allclose(dt2, 2.0 * dt, rtol=1e-6, atol=0.0)
    assert torch.allclose(gp2, 2.0 * gp, rtol=1e-6, atol=1e-30)
    # 3. gradients of the batch = sum of the gradients of its halves (what the data-parallel exchange relies on)
    h = B // 2
    _, dta, gpa = _grads(m, vision[:h], text[:h], d_out[:h])
    _, dtb, gpb = _grads(m, vision[h:], text[h:], d_out[h:])
    assert float((torch.cat([dta, dtb]) - dt).abs().max() / dt.abs().max()) <= 1e-6
    assert float((gpa + gpb - gp).norm() / gp.norm()) <= 5e-3
    # 4. the K/V cache is idempotent: building it twice and reading it twice gives the same bits
    from vlm_bridge_b200 import VisionKVCache

    with torch.no_grad():
        c1, c2 = VisionKVCache(m, vision), VisionKVCache(m, vision)
        assert torch.equal(c1.kv, c2.kv) and torch.equal(c1.kv_packed, c2.kv_packed)
        assert torch.equal(m(None, text[:, :16], kv_cache=c1), m(None, text[:, :16], kv_cache=c2))
