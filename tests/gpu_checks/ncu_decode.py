"""Two decode cross-attention launches (s = 16 and s = 64, C4 shape) for an `ncu --set full` capture."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

from vlm_bridge_b200 import ops

B, H, HD, NV = 32, 8, 288, 257
D = H * HD
kv = torch.randn(B * NV, 4 * D, device="cuda").bfloat16()
for s in (16, 64):
    q = torch.randn(B * s, D, device="cuda").bfloat16()
    for i in (0, 1, 0, 1):
        ops.attention_fwd(q, kv[:, 2 * D * i:2 * D * i + D], kv[:, 2 * D * i + D:2 * D * (i + 1)], batch=B, heads=H,
                          len_q=s, len_k=NV, head_dim=HD)
torch.cuda.synchronize()
