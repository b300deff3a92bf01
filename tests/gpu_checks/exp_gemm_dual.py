"""Grouped dgrad + wgrad launch (b200b_gemm_dual) against the two separate launches, per layer shape of the C2
backward; 20 calls captured in one CUDA graph, replayed 5 times. Run: python tests/gpu_checks/exp_gemm_dual.py"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

from vlm_bridge_b200 import _lib, ops

lib = _lib.lib()
T, D, F = 1024, 2304, 9216


def timed(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (5 * n) * 1e3


for name, rows, n_out, n_in in (("proj 2304x2304", T, D, D), ("qkv 6912x2304", T, 3 * D, D), ("ffn.0 9216x2304 (falls back)", T, F, D),
                                ("proj at C5 rows", 2048, D, D)):
    dy = (torch.randn(rows, n_out, device="cuda") * 0.5).bfloat16()
    w = (torch.randn(n_out, n_in, device="cuda") * 0.02).bfloat16()
    x = (torch.randn(rows, n_in, device="cuda") * 0.5).bfloat16()
    row = {"layer": name, "flop": 4.0 * rows * n_out * n_in}
    for on, tag in ((1, "grouped_us"), (0, "two_launches_us")):
        lib.b200b_gemm_set_dual(on)
        row[tag] = round(timed(lambda: ops.gemm_grad_pair(dy, w, x)), 2)
    lib.b200b_gemm_set_dual(1)
    row["grouped_tflops"] = round(row["flop"] / row["grouped_us"] / 1e6)
    row["two_launches_tflops"] = round(row["flop"] / row["two_launches_us"] / 1e6)
    print(json.dumps(row), flush=True)
