"""ncu target: one forward + backward of the training attention at the C2 shapes (cross d=288 over 257 keys, self
d=128 over 128 keys), dropout 0.1, tcgen05 kernels; and the same through the mma.sync kernels.
ncu --metrics gpu__time_duration.sum -k regex:attn_ python tests/gpu_checks/ncu_attn_tc.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

from vlm_bridge_b200 import _lib, ops

lib = _lib.lib()
for B, H, HD, Lq, Lk in ((8, 8, 288, 128, 257), (8, 18, 128, 128, 128)):
    D = H * HD
    g = torch.Generator().manual_seed(1)
    q = torch.randn(B * Lq, D, generator=g).bfloat16().cuda()
    kv = torch.randn(B * Lk, 2 * D, generator=g).bfloat16().cuda()
    k, v = kv[:, :D], kv[:, D:]
    d_o = torch.randn(B * Lq, D, generator=g).bfloat16().cuda()
    dq, dkv = torch.empty_like(q), torch.empty_like(kv)
    kw = dict(batch=B, heads=H, len_q=Lq, len_k=Lk, head_dim=HD, dropout_p=0.1, seed=99, dropout_stream=2)
    for mask in (3, 0):
        lib.b200b_attention_set_tc(mask)
        for _ in range(2):
            o, lse = ops.attention_fwd(q, k, v, **kw)
            ops.attention_bwd(d_o, q, k, v, o, lse, dq, dkv[:, :D], dkv[:, D:], **kw)
        torch.cuda.synchronize()
