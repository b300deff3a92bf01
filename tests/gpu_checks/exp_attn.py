"""Timing (and ncu target) of the attention kernels at the bridge's shapes (run under gpurun)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

from vlm_bridge_b200 import ops

SHAPES = {"cross": (8, 8, 128, 257, 288), "self": (8, 18, 128, 128, 128), "cross_c5": (16, 8, 128, 1370, 288),
          "decode32": (32, 8, 32, 257, 288), "decode1": (32, 8, 1, 257, 288)}
which = os.environ.get("ATTN", "cross,self").split(",")
iters = int(os.environ.get("ITERS", "20"))
p = float(os.environ.get("P", "0.1"))
for name in which:
    B, H, Lq, Lk, d = SHAPES[name]
    D = H * d
    q = torch.randn(B * Lq, D, device="cuda").bfloat16()
    kv = torch.randn(B * Lk, 2 * D, device="cuda").bfloat16()
    k, v = kv[:, :D], kv[:, D:]
    kw = dict(batch=B, heads=H, len_q=Lq, len_k=Lk, head_dim=d, dropout_p=p, seed=5, dropout_stream=1)
    o, lse = ops.attention_fwd(q, k, v, **kw)
    d_o = torch.randn(B * Lq, D, device="cuda").bfloat16()
    dq = torch.empty_like(q)
    dkv = torch.empty_like(kv)
    ws = torch.empty(ops._lib.lib().b200b_attention_bwd_workspace_bytes(B, H, Lq, Lk), device="cuda", dtype=torch.uint8)

    def fwd():
        ops.attention_fwd(q, k, v, out=o, **kw)

    def bwd():
        ops.attention_bwd(d_o, q, k, v, o, lse, dq, dkv[:, :D], dkv[:, D:], workspace=ws, **kw)

    res = {"shape": name, "B": B, "H": H, "Lq": Lq, "Lk": Lk, "d": d, "p": p}
    for label, fn, mult in (("fwd", fwd, 4.0), ("bwd", bwd, 10.0)):
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / iters * 1e3
        res[label + "_us"] = round(us, 1)
        res[label + "_tflops"] = round(mult * B * H * Lq * Lk * d / us / 1e6, 1)
    print(json.dumps(res), flush=True)
