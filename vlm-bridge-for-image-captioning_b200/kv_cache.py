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
"""
from __future__ import annotations

import torch

__all__ = ["VisionKVCache"]


class VisionKVCache:
    def __init__(self, bridge, vision_features: torch.Tensor):
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

    @property
    def nbytes(self) -> int:
        """bytes of the decode-layout cache (what a decode step reads)"""
        return self.kv_packed.numel() * self.kv_packed.element_size()

    def is_current(self) -> bool:
        """False once the bridge weights were updated after the cache was filled."""
        return self._versions == tuple(p._version for p in self._bridge.parameters())
