"""Microbenchmark of the decode cross-attention (config C4: batch 32, 8 heads of 288, 257 cached
vision keys, prefix length s = 1..64): time per launch and achieved HBM GB/s of the algorithmic bytes
(K and V read once, Q read, O written), with the cache of BOTH blocks touched between launches so the
126 MB L2 cannot hold a block's 76 MB of K/V from one step to the next."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

from vlm_bridge_b200 import ops

B, H, HD, NV, NB = 32, 8, 288, 257, 2
D = H * HD
dev = "cuda"
torch.manual_seed(0)
kv = torch.randn(B * NV, NB * 2 * D, device=dev).bfloat16()
PACKED = "--rows" not in sys.argv
kvp = ops.kv_cache_pack(kv, batch=B, len_k=NV, heads=H, head_dim=HD, num_blocks=NB)
out = []
for s in (1, 8, 16, 17, 32, 33, 48, 49, 64):
    q = torch.randn(B * s, D, device=dev).bfloat16()
    o = torch.empty(B * s, D, device=dev, dtype=torch.bfloat16)

    def go(i):
        if PACKED:
            ops.attention_decode_packed(q, kvp, block_index=i, num_blocks=NB, batch=B, heads=H, len_q=s, len_k=NV,
                                        head_dim=HD, out=o, want_lse=False)
            return
        k = kv[:, 2 * D * i:2 * D * i + D]
        v = kv[:, 2 * D * i + D:2 * D * (i + 1)]
        ops.attention_fwd(q, k, v, batch=B, heads=H, len_q=s, len_k=NV, head_dim=HD, out=o)

    for _ in range(3):
        go(0); go(1)
    torch.cuda.synchronize()
    if PACKED and not os.environ.get("B200B_DECODE_DEBUG"):   # the two layouts run the same arithmetic in the same order: results must be identical
        o_p = o.clone()
        ops.attention_fwd(q, kv[:, 2 * D:3 * D], kv[:, 3 * D:4 * D], batch=B, heads=H, len_q=s, len_k=NV, head_dim=HD, out=o)
        assert torch.equal(o, o_p), "packed decode differs from the row-major decode"
    reps = 20
    graph = torch.cuda.CUDAGraph()      # replayed: the launches are back to back on the device, not host paced
    with torch.cuda.graph(graph):
        for _ in range(reps):
            go(0); go(1)                # alternate blocks: 152 MB working set > L2
    graph.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    graph.replay()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / (2 * reps) * 1e3
    nbytes = B * 2 * NV * D * 2 + 2 * B * s * D * 2
    out.append({"s": s, "us_per_launch": round(us, 2), "GBs": round(nbytes / us * 1e-3, 1),
                "frac_of_6464": round(nbytes / us * 1e-3 / 6463.7, 3)})
    print(json.dumps(out[-1]), flush=True)
