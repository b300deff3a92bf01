"""Round-2 experiment: can the per-step bf16 re-cast of the weights (148 us, 8 % of the C2 step, HBM-bound
at 99 %) hide behind the forward? The cast of block 1's weights (74 M fp32 -> bf16: 297 MB read, 149 MB
written) is issued on a side stream while a block-0-shaped chain of forward GEMMs runs on the compute stream.

  gemm_alone_us   the GEMM chain alone          cast_alone_us   the cast alone
  serial_us       cast, then the chain, one stream (today: the whole cast precedes the forward)
  overlap_us      chain on the compute stream + cast on a side stream, joined at the end
A win needs overlap_us well below serial_us, i.e. the chain must not slow down by what the cast costs alone
(the gradient exchange, which also streams through L2/HBM next to the GEMMs, did slow them).

Run: python tests/gpu_checks/exp_cast_overlap.py > gpurun_out/exp_cast_overlap.jsonl   (not a pytest file)
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

from vlm_bridge_b200 import ops

T, D, F = 1024, 2304, 9216
dev = torch.device("cuda")
x = torch.randn(T, D, device=dev).bfloat16()
hid = torch.randn(T, F, device=dev).bfloat16()
w_dd = [torch.randn(D, D, device=dev).bfloat16() for _ in range(3)]
w_qkv = torch.randn(3 * D, D, device=dev).bfloat16()
w_up = torch.randn(F, D, device=dev).bfloat16()
w_dn = torch.randn(D, F, device=dev).bfloat16()
o_d = torch.empty(T, D, device=dev, dtype=torch.bfloat16)
o_3d = torch.empty(T, 3 * D, device=dev, dtype=torch.bfloat16)
o_f = torch.empty(T, F, device=dev, dtype=torch.bfloat16)
n_block = 2 * D * D + 4 * D * D + 2 * D * F            # 2-D weights of one block (without K/V): 74.3 M
w32 = torch.randn(n_block, device=dev)
w16 = torch.empty(n_block, device=dev, dtype=torch.bfloat16)
side = torch.cuda.Stream()


def chain():                                           # the GEMMs of one block's forward
    ops.gemm(x, w_dd[0], out=o_d)
    ops.gemm(x, w_dd[1], out=o_d)
    ops.gemm(x, w_qkv, out=o_3d)
    ops.gemm(x, w_dd[2], out=o_d)
    ops.gemm(x, w_up, out=o_f)
    ops.gemm(hid, w_dn, out=o_d)


def cast():
    ops.cast_bf16(w32, out=w16)


def serial():
    cast()
    chain()


def overlap():
    cur = torch.cuda.current_stream()
    side.wait_stream(cur)
    with torch.cuda.stream(side):
        cast()
    chain()
    cur.wait_stream(side)


def timed(fn, iters=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return round(e0.elapsed_time(e1) / iters * 1e3, 2)


row = {"gemm_alone_us": timed(chain), "cast_alone_us": timed(cast), "serial_us": timed(serial), "overlap_us": timed(overlap),
       "cast_bytes": n_block * 6}
row["hidden_fraction_of_cast"] = round((row["serial_us"] - row["overlap_us"]) / row["cast_alone_us"], 3)
print(json.dumps(row), flush=True)
