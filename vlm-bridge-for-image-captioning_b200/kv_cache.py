"""Per-image cached vision K/V for autoregressive caption decode.

In the reference decode loop (full_model.py:241-363) every step calls the bridge on the whole token
prefix and re-projects the unchanged image through w_k / w_v of every block
(bridge_module.py:99-100) -- 155 GFLOP per step at batch 32 that depend only on the image. Because the
bridge self-attention is non-causal and unmasked (:138,236) the text side cannot be cached exactly,
but the vision K/V can: they are computed once per image here and every decode step reads them.

The cache is two plain torch tensors owned by this object: `kv` (bf16, [B*Nv, num_blocks*2*D]; block
i's K at columns [2iD, 2iD+D), V in the next D columns -- the projection output, used for steps of more
than 64 positions), `kv_packed` (the same values as [B][num_blocks][heads][2][Nv][d+8], the layout
the mma.sync decode kernel streams with one bulk copy per 16-key tile; steps of up to 32 positions) and
`kv_tc` (32-key tiles as the swizzled operand images of the tcgen05 decode kernel; steps of 33..64). The CUDA library only ever sees their
pointers for the duration of a call.

Position rows (`x1`, fp32 [B, max_positions, D]): block 0's cross-attention sub-layer
(bridge_module.py:316-323) maps text row (b, j) to x1[b, j] = x + W_o * SDPA(W_q * LN(x), K_b, V_b),
which involves no other text position, so a decode loop that appends tokens computes it once per
position (`BridgeLite.forward(..., kv_cache=cache, cached_positions=k)`) instead of once per position
per step. Everything after it (block 0's non-causal self-attention onwards) mixes positions and is
recomputed on the whole prefix, as in the reference.
"""
from __future__ import annotations

import torch

__all__ = ["VisionKVCache"]


class VisionKVCache:
    def __init__(self, bridge, vision_features: torch.Tensor, max_positions: int = 64):
        if vision_features.dim() != 3:
            raise RuntimeError("vision_features must be [B, Nv, vision_dim]")
        self.batch, self.len_vision = int(vision_features.shape[0]), int(vision_features.shape[1])
        with torch.no_grad():
            self.vision_bf16, self.kv = bridge.project_vision_kv(vision_features)
            # decode layout (per image / block / head: padded K rows then V rows), read by the K/V-streaming
            # cross-attention kernel whenever a step has <= 64 text positions
            self.kv_packed = bridge.pack_vision_kv(self.kv, self.batch, self.len_vision)
            # tcgen05 decode layout, read by steps of 33..64 positions (where the mma.sync kernel is bound by the
            # legacy tensor pipe); only built for the head dims that kernel is instantiated for
            d = bridge.language_dim // bridge.num_heads_cross
            self.kv_tc = bridge.pack_vision_kv_tc(self.kv, self.batch, self.len_vision) if d in (64, 128, 288) else None
        self._versions = tuple(p._version for p in bridge.parameters())
        self._bridge = bridge
        # per-position rows of block 0's cross-attention sub-layer output (see the module docstring)
        self.max_positions = int(max_positions)
        self.x1 = (torch.empty((self.batch, self.max_positions, bridge.language_dim), device=self.kv.device,
                               dtype=torch.float32) if self.max_positions > 0 else None)
        self.positions_filled = 0

    def position_rows(self, length: int, cached_positions: int, device, dim: int):
        """The row store for a forward over `length` text positions of which the first
        `cached_positions` are reused, or None when the prefix is longer than the store (the caller then
        computes every row). Raises if rows are asked for that were never written."""
        if self.x1 is None or length > self.max_positions:
            return None
        k = int(cached_positions)
        if not 0 <= k < length:
            raise RuntimeError(f"cached_positions must be in [0, {length}) for a prefix of {length} positions, got {k}")
        if k > self.positions_filled:
            raise RuntimeError(f"cached_positions={k} but only {self.positions_filled} positions are cached")
        if self.x1.device != device or self.x1.shape[-1] != dim:
            raise RuntimeError("position rows live on another device / have another width than the text")
        self.positions_filled = length
        return self.x1

    @property
    def nbytes(self) -> int:
        """bytes of the decode-layout cache (what a decode step reads)"""
        return self.kv_packed.numel() * self.kv_packed.element_size()

    def is_current(self) -> bool:
        """False once the bridge weights were updated after the cache was filled."""
        return self._versions == tuple(p._version for p in self._bridge.parameters())
