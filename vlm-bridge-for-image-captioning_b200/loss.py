"""Fused cross-entropy over the language model's vocabulary logits.

Drop-in for the loss of the reference training step (core_training_loop.py:51-55,68-69):

    labels = input_ids.clone(); labels[:, :-1] = input_ids[:, 1:]; labels[:, -1] = -100
    loss = nn.CrossEntropyLoss(ignore_index=-100)(logits.view(-1, V), labels.view(-1))

`FusedCrossEntropyLoss` has the call signature of that `nn.CrossEntropyLoss` (mean reduction over the
non-ignored rows); `forward_shifted(logits [B, L, V], input_ids [B, L])` also does the label shift
inside the kernel. The logits may be fp32 or bf16 (arithmetic is fp32, as autocast runs
cross_entropy); the gradient comes back in the logits' dtype. Forward reads the logits once, backward
reads them once and writes the gradient once; only the per-row log-sum-exp is saved (PyTorch keeps a
full fp32 log-softmax, 1.05 GB at B=8, L=128, V=256000). CUDA only, no fallback.
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn as nn

from . import _lib

__all__ = ["FusedCrossEntropyLoss", "fused_cross_entropy"]

_DTYPES = {torch.float32: 0, torch.bfloat16: 1}


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class _FusedCE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits: torch.Tensor, labels: torch.Tensor, ignore_index: int, shift_len: int):
        rows, vocab = logits.shape
        dev = logits.device
        lse = torch.empty(rows, device=dev, dtype=torch.float32)
        loss_rows = torch.empty(rows, device=dev, dtype=torch.float32)
        out2 = torch.empty(2, device=dev, dtype=torch.float32)
        _lib.check(_lib.lib().b200b_cross_entropy_fwd(
            logits.data_ptr(), _DTYPES[logits.dtype], logits.stride(0), labels.data_ptr(), rows, vocab, ignore_index,
            shift_len, lse.data_ptr(), loss_rows.data_ptr(), out2.data_ptr(), _stream()), "cross_entropy_fwd")
        ctx.save_for_backward(logits, labels, lse, out2)
        ctx.ignore_index, ctx.shift_len = ignore_index, shift_len
        ctx.mark_non_differentiable(loss_rows)
        return out2[0], loss_rows

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_loss: torch.Tensor, _grad_rows):
        logits, labels, lse, out2 = ctx.saved_tensors
        rows, vocab = logits.shape
        g = grad_loss.detach().to(device=logits.device, dtype=torch.float32).reshape(1).contiguous()
        dlogits = torch.empty((rows, vocab), device=logits.device, dtype=logits.dtype)
        _lib.check(_lib.lib().b200b_cross_entropy_bwd(
            logits.data_ptr(), _DTYPES[logits.dtype], logits.stride(0), labels.data_ptr(), rows, vocab,
            ctx.ignore_index, ctx.shift_len, lse.data_ptr(), out2.data_ptr(), g.data_ptr(), dlogits.data_ptr(),
            dlogits.stride(0), _stream()), "cross_entropy_bwd")
        return dlogits, None, None, None


def _check(logits: torch.Tensor, labels: torch.Tensor) -> None:
    if logits.device.type != "cuda" or labels.device != logits.device:
        raise RuntimeError("fused_cross_entropy runs on CUDA only (logits and labels on the same device); "
                           "there is no CPU fallback")
    if logits.dtype not in _DTYPES:
        raise RuntimeError(f"fused_cross_entropy: logits must be float32 or bfloat16, got {logits.dtype}")
    if labels.dtype != torch.int64:
        raise RuntimeError(f"fused_cross_entropy: labels must be int64, got {labels.dtype}")


def fused_cross_entropy(logits: torch.Tensor, labels: torch.Tensor, ignore_index: int = -100,
                        return_row_losses: bool = False):
    """Mean cross-entropy of `logits` [N, V] against `labels` [N] over the rows whose label is not
    `ignore_index` == F.cross_entropy(logits.float(), labels, ignore_index=ignore_index)."""
    _check(logits, labels)
    if logits.dim() != 2 or labels.dim() != 1 or labels.shape[0] != logits.shape[0]:
        raise RuntimeError("fused_cross_entropy: expected logits [N, V] and labels [N]")
    if logits.stride(1) != 1:
        logits = logits.contiguous()
    loss, rows = _FusedCE.apply(logits, labels.contiguous(), int(ignore_index), 0)
    return (loss, rows) if return_row_losses else loss


class FusedCrossEntropyLoss(nn.Module):
    """`nn.CrossEntropyLoss(ignore_index=...)` as the reference uses it (core_training_loop.py:68-69):
    no class weights, no label smoothing, mean reduction."""

    def __init__(self, ignore_index: int = -100, reduction: str = "mean"):
        super().__init__()
        if reduction != "mean":
            raise ValueError("FusedCrossEntropyLoss implements reduction='mean' (the reference's) only")
        self.ignore_index = int(ignore_index)
        self.reduction = reduction

    def forward(self, input: torch.Tensor, target: torch.Tensor) -> torch.Tensor:  # noqa: A002
        return fused_cross_entropy(input, target, self.ignore_index)

    def forward_shifted(self, logits: torch.Tensor, input_ids: torch.Tensor) -> torch.Tensor:
        """Loss of next-token prediction straight from `input_ids` [B, L]: position j is scored against
        input_ids[:, j + 1] and the last position is ignored -- the shift the reference builds on the
        host (core_training_loop.py:51-54) happens inside the kernel. logits [B, L, V] or [B*L, V]."""
        _check(logits, input_ids)
        if input_ids.dim() != 2:
            raise RuntimeError("forward_shifted: input_ids must be [B, L]")
        B, L = input_ids.shape
        flat = logits.reshape(B * L, logits.shape[-1])
        if flat.stride(1) != 1:
            flat = flat.contiguous()
        loss, _ = _FusedCE.apply(flat, input_ids.contiguous(), self.ignore_index, L)
        return loss
