"""Whole-step CUDA graph for the bridge: forward + loss + backward captured once, replayed per step.

The bridge step is ~100 short kernels; enqueueing them from Python costs about as long as they run
(and more when eight ranks share the host), so a training loop that is otherwise static can replay
one captured graph instead. Everything the module does per step is capture-safe: the kernels take
no host-side per-step values (the dropout seed lives in device memory and is advanced by a captured
kernel; the data-parallel collectives number themselves from a device counter), scratch memory comes
from the graph's private pool, and the bf16 weight copies are re-cast inside the graph, so an
optimizer may update the fp32 parameters (in place) between replays.

The reference trains eagerly (core_training_loop.py:60-104); this is an additive helper, not a
mirror of a reference file.
"""
from __future__ import annotations

from typing import Callable

import torch

__all__ = ["GraphedBridgeStep"]


class GraphedBridgeStep:
    """step = GraphedBridgeStep(bridge, loss_fn, vision, text); loss = step(vision, text)

    After every call the parameters' `.grad` hold this step's gradients (the same tensors each time;
    they are overwritten, not accumulated) and the returned loss tensor is the graph's static output.
    Shapes are fixed by the example inputs.
    """

    def __init__(self, bridge: torch.nn.Module, loss_fn: Callable[[torch.Tensor], torch.Tensor],
                 vision: torch.Tensor, text: torch.Tensor, warmup: int = 3):
        self.bridge, self.loss_fn = bridge, loss_fn
        self.vision = vision.detach().clone()
        self.text = text.detach().clone()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):                       # allocator / lazy-init warm-up, as torch requires
            for _ in range(max(1, warmup)):
                self._eager()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        for p in bridge.parameters():
            p.grad = None                                   # .grad must be allocated from the graph's pool
        from . import _lib

        before = _lib.launch_count()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss = self._eager()
        self.kernels_per_replay = _lib.launch_count() - before   # this library's launches in one replay

    def _eager(self) -> torch.Tensor:
        out = self.bridge(self.vision, self.text)
        loss = self.loss_fn(out)
        loss.backward()
        return loss

    def replay(self) -> torch.Tensor:
        """Run the step on whatever the static input buffers hold."""
        self.graph.replay()
        return self.loss

    def __call__(self, vision: torch.Tensor, text: torch.Tensor) -> torch.Tensor:
        self.vision.copy_(vision, non_blocking=True)
        self.text.copy_(text, non_blocking=True)
        return self.replay()
