"""Throughput sweep of the GEMM tile configurations on the bridge's shapes (run under gpurun)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from vlm_bridge_b200 import ops

def bench(M, N, K, bn, cg, am=0, bm=0, epi=0, iters=20):
    a = (torch.randn((K, M) if am else (M, K), device="cuda") * 0.5).bfloat16()
    b = (torch.randn((K, N) if bm else (N, K), device="cuda") * 0.5).bfloat16()
    o = torch.empty(M, N, device="cuda", dtype=torch.float32 if epi == 4 else torch.bfloat16)
    kw = dict(a_major=am, b_major=bm, epilogue=epi, block_n=bn, cta_group=cg, out=o)
    for _ in range(3):
        ops.gemm(a, b, **kw)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        ops.gemm(a, b, **kw)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    return ms, 2.0 * M * N * K / ms / 1e9

mode = os.environ.get("B200B_GEMM_DEBUG", "0")
# (name, M, N, K, a_major, b_major, epi)
SHAPES = [("ffn_up", 1024, 9216, 2304, 0, 0, 0), ("ffn_down", 1024, 2304, 9216, 0, 0, 0),
          ("proj", 1024, 2304, 2304, 0, 0, 0), ("qkv", 1024, 6912, 2304, 0, 0, 0),
          ("kv_proj", 2056, 9216, 1024, 0, 0, 0),
          ("dgrad_ffn_down", 1024, 9216, 2304, 0, 1, 0), ("dgrad_ffn_up", 1024, 2304, 9216, 0, 1, 0),
          ("dgrad_proj", 1024, 2304, 2304, 0, 1, 0), ("dgrad_qkv", 1024, 2304, 6912, 0, 1, 0),
          ("wgrad_w1", 9216, 2304, 1024, 1, 1, 4), ("wgrad_w2", 2304, 9216, 1024, 1, 1, 4),
          ("wgrad_proj", 2304, 2304, 1024, 1, 1, 4), ("wgrad_qkv", 6912, 2304, 1024, 1, 1, 4),
          ("wgrad_kv", 9216, 1024, 2056, 1, 1, 4), ("big", 4096, 4096, 4096, 0, 0, 0)]
for name, M, N, K, am, bm, epi in SHAPES:
    row = {"mode": mode, "shape": name, "M": M, "N": N, "K": K}
    for bn, cg in [(128, 1), (256, 1), (128, 2), (256, 2), (0, 0)]:
        ms, tf = bench(M, N, K, bn, cg, am, bm, epi)
        row[f"cg{cg}_bn{bn}"] = round(tf)
    a = torch.randn(M, K, device="cuda").bfloat16(); b = torch.randn(N, K, device="cuda").bfloat16()
    for _ in range(3): a @ b.t()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): a @ b.t()
    e1.record(); torch.cuda.synchronize()
    row["cublas"] = round(2.0 * M * N * K / (e0.elapsed_time(e1) / 20) / 1e9)
    print(json.dumps(row), flush=True)
