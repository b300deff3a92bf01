"""Cost of each fused epilogue on the bridge's FFN shapes (run under gpurun; prints JSON lines)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

from vlm_bridge_b200 import ops


def bench(M, N, K, epi, p, bm=0, iters=20, **extra):
    a = (torch.randn(M, K, device="cuda") * 0.5).bfloat16()
    b = (torch.randn((K, N) if bm else (N, K), device="cuda") * 0.05).bfloat16()
    f32 = epi in (2, 4)
    o = torch.empty(M, N, device="cuda", dtype=torch.float32 if f32 else torch.bfloat16)
    kw = dict(b_major=bm, epilogue=epi, out=o, dropout_p=p, seed=1234, dropout_stream=3, **extra)
    if epi in (0, 1, 2):
        kw["bias"] = torch.randn(N, device="cuda")
    if epi in (1, 3):
        kw["aux"] = torch.randn(M, N, device="cuda").bfloat16()
    if epi == 2:
        kw["resid"] = torch.randn(M, N, device="cuda")
    for _ in range(3):
        ops.gemm(a, b, **kw)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        ops.gemm(a, b, **kw)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    return round(ms * 1e3, 1), round(2.0 * M * N * K / ms / 1e9)


NAMES = {0: "bias", 1: "gelu", 2: "resid", 3: "dgelu", 4: "f32"}
SHAPES = [(1024, 9216, 2304, 0), (1024, 9216, 2304, 1), (1024, 2304, 9216, 0), (1024, 2304, 2304, 0)]
if os.environ.get('EXP_SHAPES'):
    SHAPES = SHAPES[:int(os.environ['EXP_SHAPES'])]
for (M, N, K, bm) in SHAPES:
    for epi in ((0, 3) if bm else (0, 1, 3, 2)):
        for p in (0.0, 0.1):
            if epi == 0 and p > 0:
                continue
            us, tf = bench(M, N, K, epi, p, bm)
            print(json.dumps({"M": M, "N": N, "K": K, "b_major": bm, "epi": NAMES[epi], "p": p, "us": us, "tflops": tf}),
                  flush=True)
