"""Fused cross-entropy against torch.nn.functional.cross_entropy in fp32 --
the arithmetic the reference's `nn.CrossEntropyLoss(ignore_index=-100)` runs under autocast
(core_training_loop.py:51-55,68-69).

Tolerances: loss |d| <= 1e-5 * max(1, |ref|) (fp32 sums in another order, ex2.approx); gradients of fp32
logits rtol 1e-4 + atol 1e-7 * (upstream / count); gradients of bf16 logits are the fp32 reference
rounded to bf16, compared at one bf16 ulp (rtol 8e-3)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _ref(logits, labels, upstream):
    x = logits.detach().float().requires_grad_()
    loss = F.cross_entropy(x, labels, ignore_index=-100)
    (loss * upstream).backward()
    return loss.detach(), x.grad


@pytest.mark.parametrize("rows,vocab,dtype", [(64, 256000, torch.float32), (48, 256000, torch.bfloat16),
                                              (37, 1003, torch.float32), (37, 1003, torch.bfloat16),
                                              (5, 8, torch.float32), (3, 40000, torch.float32)])
def test_matches_torch_cross_entropy(rows, vocab, dtype):
    from vlm_bridge_b200 import FusedCrossEntropyLoss

    g = torch.Generator().manual_seed(rows * 7 + vocab)
    logits = (torch.randn(rows, vocab, generator=g) * 3.0).to(dtype).cuda()
    labels = torch.randint(0, vocab, (rows,), generator=g)
    labels[::5] = -100                                     # ignored rows
    labels = labels.cuda()
    upstream = 1024.0                                      # e.g. a GradScaler scale
    loss_ref, grad_ref = _ref(logits, labels, upstream)
    x = logits.clone().requires_grad_()
    loss = FusedCrossEntropyLoss(ignore_index=-100)(x, labels)
    (loss * upstream).backward()
    assert loss.dtype == torch.float32 and loss.dim() == 0
    assert abs(float(loss) - float(loss_ref)) <= 1e-5 * max(1.0, abs(float(loss_ref)))
    assert x.grad.dtype == dtype and x.grad.shape == logits.shape
    count = int((labels != -100).sum())
    if dtype == torch.float32:
        assert torch.allclose(x.grad, grad_ref, rtol=1e-4, atol=1e-7 * upstream / count)
    else:
        assert torch.allclose(x.grad.float(), grad_ref.bfloat16().float(), rtol=8e-3, atol=1e-6 * upstream / count)
    assert bool((x.grad[::5] == 0).all())                  # ignored rows get an exactly-zero gradient


def test_label_shift_inside_the_kernel_and_row_pitch():
    """forward_shifted(logits [B, L, V], input_ids [B, L]) == the reference's host-side shift + loss; the
    logits may be a column slice of a wider matrix (row pitch > vocab)."""
    from vlm_bridge_b200 import FusedCrossEntropyLoss

    g = torch.Generator().manual_seed(11)
    B, L, V = 3, 9, 4096
    wide = torch.randn(B, L, V + 64, generator=g).cuda()
    ids = torch.randint(3, V, (B, L), generator=g).cuda()
    labels = ids.clone()
    labels[:, :-1] = ids[:, 1:]
    labels[:, -1] = -100                                   # core_training_loop.py:52-54
    logits = wide[..., :V]
    loss_ref, grad_ref = _ref(logits.reshape(B * L, V), labels.reshape(-1), 1.0)
    x = wide.clone().requires_grad_()
    loss = FusedCrossEntropyLoss().forward_shifted(x[..., :V], ids)
    loss.backward()
    assert abs(float(loss) - float(loss_ref)) <= 1e-5 * max(1.0, abs(float(loss_ref)))
    assert torch.allclose(x.grad[..., :V].reshape(B * L, V), grad_ref, rtol=1e-4, atol=1e-7 / (B * (L - 1)))
    assert bool((x.grad[..., V:] == 0).all())


def test_edge_cases():
    from vlm_bridge_b200 import FusedCrossEntropyLoss, fused_cross_entropy

    logits = torch.randn(4, 64).cuda()
    all_ignored = torch.full((4,), -100, dtype=torch.int64).cuda()
    assert torch.isnan(fused_cross_entropy(logits, all_ignored))          # mean over zero rows, as torch
    bad = torch.tensor([1, 2, 64, 3]).cuda()
    assert torch.isnan(fused_cross_entropy(logits, bad))                  # label outside [0, vocab)
    big = logits.clone()
    big[0, 3] = 3e4                                                       # no overflow: max is subtracted
    lab = torch.tensor([3, 1, 2, 0]).cuda()
    assert abs(float(fused_cross_entropy(big, lab)) - float(F.cross_entropy(big, lab))) <= 1e-5
    loss, rows = fused_cross_entropy(logits, lab, return_row_losses=True)
    assert torch.allclose(rows, F.cross_entropy(logits, lab, reduction="none"), rtol=1e-5, atol=1e-5)
    with pytest.raises(RuntimeError):
        fused_cross_entropy(logits.cpu(), lab.cpu())                      # no CPU fallback
    with pytest.raises(RuntimeError):
        fused_cross_entropy(logits.half(), lab)
    with pytest.raises(ValueError):
        FusedCrossEntropyLoss(reduction="sum")
