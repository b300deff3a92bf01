"""tcgen05 decode cross-attention against a torch fp32 reference (and the mma.sync decode kernel), then the
graph-timed rate per prefix length at the C4 shape."""
import json
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

from vlm_bridge_b200 import ops

torch.manual_seed(0)
dev = "cuda"
ok_all = True
for B, H, HD, Lq, Lk, NB in ((1, 1, 64, 16, 64, 1), (2, 2, 128, 5, 100, 2), (2, 8, 288, 1, 257, 2), (3, 8, 288, 17, 257, 2),
                             (2, 8, 288, 64, 257, 2), (1, 8, 288, 40, 16, 1), (1, 8, 288, 33, 130, 2)):
    D = H * HD
    kv = torch.randn(B * Lk, NB * 2 * D, device=dev).bfloat16()
    q = torch.randn(B * Lq, D, device=dev).bfloat16()
    kvt = ops.kv_cache_pack_tc(kv, batch=B, len_k=Lk, heads=H, head_dim=HD, num_blocks=NB)
    blk = NB - 1
    o, lse = ops.attention_decode_tc(q, kvt, block_index=blk, num_blocks=NB, batch=B, heads=H, len_q=Lq, len_k=Lk, head_dim=HD)
    torch.cuda.synchronize()
    k = kv[:, 2 * D * blk:2 * D * blk + D].float().reshape(B, Lk, H, HD).transpose(1, 2)
    v = kv[:, 2 * D * blk + D:2 * D * (blk + 1)].float().reshape(B, Lk, H, HD).transpose(1, 2)
    qq = q.float().reshape(B, Lq, H, HD).transpose(1, 2)
    s = qq @ k.transpose(-1, -2) / math.sqrt(HD)
    ref = (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(B * Lq, D)
    e_o = float((o.float() - ref).abs().max() / ref.abs().max())
    e_lse = float((lse * math.log(2) - torch.logsumexp(s, -1)).abs().max())
    ok = e_o < 1e-2 and e_lse < 1e-3
    ok_all &= ok
    print(json.dumps({"case": [B, H, HD, Lq, Lk, NB], "e_o": e_o, "e_lse": e_lse, "ok": ok}), flush=True)
print("all ok" if ok_all else "FAILED", flush=True)
if (not ok_all and not os.environ.get("B200B_DECODE_DEBUG")) or "--no-time" in sys.argv:
    sys.exit(0 if ok_all else 1)
B, H, HD, NV, NB = 32, 8, 288, 257, 2
D = H * HD
kv = torch.randn(B * NV, NB * 2 * D, device=dev).bfloat16()
kvt = ops.kv_cache_pack_tc(kv, batch=B, len_k=NV, heads=H, head_dim=HD, num_blocks=NB)
for s in (1, 16, 17, 32, 33, 48, 64):
    q = torch.randn(B * s, D, device=dev).bfloat16()
    o = torch.empty(B * s, D, device=dev, dtype=torch.bfloat16)

    def go(i):
        ops.attention_decode_tc(q, kvt, block_index=i, num_blocks=NB, batch=B, heads=H, len_q=s, len_k=NV, head_dim=HD,
                                out=o, want_lse=False)

    for _ in range(3):
        go(0); go(1)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(20):
            go(0); go(1)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 40 * 1e3
    nbytes = B * 2 * NV * D * 2 + 2 * B * s * D * 2
    print(json.dumps({"s": s, "us_per_launch": round(us, 2), "GBs": round(nbytes / us * 1e-3, 1),
                      "frac_of_6464": round(nbytes / us * 1e-3 / 6463.7, 3)}), flush=True)
