"""Decode cross-attention over the packed K/V cache at the C4 shape for an `ncu --set full` capture:
1 query position per image (every block-0 launch of a decode step with position rows) and 16."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

from vlm_bridge_b200 import ops

B, H, HD, NV, NB = 32, 8, 288, 257, 2
D = H * HD
kv = torch.randn(B * NV, NB * 2 * D, device="cuda").bfloat16()
kvp = ops.kv_cache_pack(kv, batch=B, len_k=NV, heads=H, head_dim=HD, num_blocks=NB)
for s in (1, 16):
    q = torch.randn(B * s, D, device="cuda").bfloat16()
    for i in (0, 1):
        ops.attention_decode_packed(q, kvp, block_index=i, num_blocks=NB, batch=B, heads=H, len_q=s, len_k=NV, head_dim=HD)
torch.cuda.synchronize()
