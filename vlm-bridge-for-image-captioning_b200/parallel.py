"""Data-parallel training of the bridge: one process per GPU, batch sharded across ranks, only the
bridge gradients exchanged (the frozen encoders carry no gradients).

The reference has no distributed code at all, so this is new code beside its training loop, not a
mirror of any reference file. The exchange step is a bucketed NCCL all-reduce (average) over the
flat gradient arena that `BridgeLite` writes during backward: buckets are the contiguous slabs a
block's backward has just finished (last block first), launched asynchronously on NCCL's stream as
soon as the producing kernels are enqueued, so the reduction of block i overlaps the backward
kernels of block i-1. `finish()` joins the communication stream before autograd hands the
gradients to the optimizer path (GradScaler.unscale_, clip_grad_norm_, AdamW).
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist

__all__ = ["GradBucketReducer", "enable_data_parallel", "disable_data_parallel", "broadcast_parameters"]


class GradBucketReducer:
    """Callable bucket hook: reducer(arena, start, end) starts averaging arena[start:end] over ranks."""

    def __init__(self, process_group: Optional[dist.ProcessGroup] = None, max_bucket_elems: int = 1 << 26):
        self.group = process_group
        self.world_size = dist.get_world_size(process_group)
        self.max_bucket_elems = int(max_bucket_elems)
        self._pending: list[tuple[dist.Work, torch.Tensor]] = []
        self.bytes_reduced = 0
        backend = dist.get_backend(process_group)
        self._native_avg = backend == "nccl"

    def __call__(self, arena: torch.Tensor, start: int, end: int) -> None:
        if self.world_size == 1 or end <= start:
            return
        pos = start
        while pos < end:
            stop = min(end, pos + self.max_bucket_elems)
            chunk = arena[pos:stop]
            op = dist.ReduceOp.AVG if self._native_avg else dist.ReduceOp.SUM
            work = dist.all_reduce(chunk, op=op, group=self.group, async_op=True)
            self._pending.append((work, chunk))
            self.bytes_reduced += chunk.numel() * chunk.element_size()
            pos = stop

    def finish(self) -> None:
        """Make the current stream wait for every outstanding bucket (and finish the average)."""
        for work, chunk in self._pending:
            work.wait()
            if not self._native_avg:
                chunk.div_(self.world_size)
        self._pending.clear()


def broadcast_parameters(module: torch.nn.Module, src: int = 0, process_group=None) -> None:
    """Every rank starts from rank `src`'s parameters (one broadcast of the flat buffer when the
    module has been flattened, else one per tensor)."""
    flat = getattr(module, "_flat", None)
    if flat is not None:
        dist.broadcast(flat, src=src, group=process_group)
        module._w16_key = None
        return
    for p in module.parameters():
        dist.broadcast(p.data, src=src, group=process_group)


def enable_data_parallel(module, process_group=None, max_bucket_elems: int = 1 << 26) -> GradBucketReducer:
    """Attach a bucketed all-reduce to `module` (a B200 BridgeLite). Returns the reducer."""
    if not dist.is_initialized():
        raise RuntimeError("torch.distributed is not initialised")
    reducer = GradBucketReducer(process_group, max_bucket_elems)
    module._bucket_hook = reducer
    return reducer


def disable_data_parallel(module) -> None:
    module._bucket_hook = None
