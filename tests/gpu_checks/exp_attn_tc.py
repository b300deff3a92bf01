"""Round-2 measurement: the tcgen05 training attention (csrc/attention_train_tc.cu) against the mma.sync kernels
(csrc/attention.cu) at the training shapes, forward and backward, with and without dropout. Kernel names come from
the library's own launch profile (so the table also proves which implementation ran); times are CUDA events
around 20 back-to-back launches. Run: python tests/gpu_checks/exp_attn_tc.py > gpurun_out/exp_attn_tc.jsonl"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

from vlm_bridge_b200 import _lib, ops

# (name, B, H, HD, Lq, Lk)
SHAPES = [("C2 cross", 8, 8, 288, 128, 257), ("C2 self", 8, 18, 128, 128, 128), ("C5 cross", 16, 8, 288, 128, 1370),
          ("C5 self", 16, 18, 128, 128, 128)]
lib = _lib.lib()


def timed(fn, n=20):
    """microseconds per call: n calls captured in ONE CUDA graph (the Python / ctypes enqueue cost of ~15 us per
    call would otherwise hide kernels this short), replayed 5 times"""
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


for name, B, H, HD, Lq, Lk in SHAPES:
    D = H * HD
    g = torch.Generator().manual_seed(1)
    q = torch.randn(B * Lq, D, generator=g).bfloat16().cuda()
    kv = torch.randn(B * Lk, 2 * D, generator=g).bfloat16().cuda()
    k, v = kv[:, :D], kv[:, D:]
    d_o = torch.randn(B * Lq, D, generator=g).bfloat16().cuda()
    dq, dkv = torch.empty_like(q), torch.empty_like(kv)
    dk, dv = dkv[:, :D], dkv[:, D:]
    for p in (0.0, 0.1):
        kw = dict(batch=B, heads=H, len_q=Lq, len_k=Lk, head_dim=HD, dropout_p=p, seed=99, dropout_stream=2)
        row = {"shape": name, "B": B, "H": H, "HD": HD, "Lq": Lq, "Lk": Lk, "dropout_p": p,
               "fwd_gflop": 4.0 * B * H * Lq * Lk * HD / 1e9}
        for mask, tag in ((3, "tc"), (0, "legacy")):
            lib.b200b_attention_set_tc(mask)
            o, lse = ops.attention_fwd(q, k, v, **kw)
            ws = torch.empty(lib.b200b_attention_bwd_workspace_bytes(B, H, Lq, Lk), device="cuda", dtype=torch.uint8)
            _lib.profile_begin(torch.cuda.current_stream().cuda_stream)
            ops.attention_fwd(q, k, v, **kw)
            ops.attention_bwd(d_o, q, k, v, o, lse, dq, dk, dv, workspace=ws, **kw)
            row[tag + "_kernels"] = [n_ for n_, _ in _lib.profile_end()]
            row[tag + "_fwd_us"] = round(timed(lambda: ops.attention_fwd(q, k, v, **kw)), 2)
            row[tag + "_bwd_us"] = round(timed(lambda: ops.attention_bwd(d_o, q, k, v, o, lse, dq, dk, dv, workspace=ws, **kw)), 2)
        lib.b200b_attention_set_tc(3)
        row["fwd_speedup"] = round(row["legacy_fwd_us"] / row["tc_fwd_us"], 2)
        row["bwd_speedup"] = round(row["legacy_bwd_us"] / row["tc_bwd_us"], 2)
        row["tc_fwd_tflops"] = round(row["fwd_gflop"] / row["tc_fwd_us"] / 1e3, 1)
        print(json.dumps(row), flush=True)
