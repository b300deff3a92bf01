"""A few tcgen05 decode cross-attention launches (s = 64, C4 shape) for an `ncu --set full` capture."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

from vlm_bridge_b200 import ops

B, H, HD, NV, NB = 32, 8, 288, 257, 2
D = H * HD
kv = torch.randn(B * NV, NB * 2 * D, device="cuda").bfloat16()
kvt = ops.kv_cache_pack_tc(kv, batch=B, len_k=NV, heads=H, head_dim=HD, num_blocks=NB)
q = torch.randn(B * 64, D, device="cuda").bfloat16()
for i in (0, 1, 0, 1):
    ops.attention_decode_tc(q, kvt, block_index=i, num_blocks=NB, batch=B, heads=H, len_q=64, len_k=NV, head_dim=HD)
torch.cuda.synchronize()
