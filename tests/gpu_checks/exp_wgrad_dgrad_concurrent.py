"""Round-2 experiment: the weight-gradient and data-gradient GEMMs of one
Linear are independent. Does running them CONCURRENTLY -- two streams, each persistent GEMM limited to half
of the SMs with b200b_set_sm_limit -- beat running them back to back on all SMs? What is measured per layer
shape of the C2 backward (T = 1024 rows):

  seq_us               wgrad then dgrad on one stream, all 148 SMs each (what block_backward does today)
  conc_half_limit_us   wgrad on stream A, dgrad on stream B, SM limit 74 (37 CTA pairs each), fork / join by events
  conc_full_limit_us   the same with no limit (the second kernel's CTAs fill in as the first's retire; the
                       limit is process-wide, so uneven splits are not possible)
  seq_half_limit_us    back to back with the limit of 74 (what each kernel costs alone on half of the SMs)

Run: python tests/gpu_checks/exp_wgrad_dgrad_concurrent.py > gpurun_out/exp_wgrad_dgrad.jsonl
Not a pytest file; nothing in the product calls it.
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

from vlm_bridge_b200 import _lib, ops

T, D, F = 1024, 2304, 9216
EPI_BF16, EPI_F32 = ops.EPI_BF16_BIAS, ops.EPI_F32
# (name, out_features N_w, in_features K_w): dY [T, N_w], X [T, K_w], W [N_w, K_w]
LAYERS = [("proj 2304x2304", D, D), ("qkv 6912x2304", 3 * D, D), ("ffn.0 9216x2304", F, D), ("ffn.3 2304x9216", D, F)]


def make(nw, kw):
    dy = (torch.randn(T, nw, device="cuda") * 0.5).bfloat16()
    x = (torch.randn(T, kw, device="cuda") * 0.5).bfloat16()
    w = (torch.randn(nw, kw, device="cuda") * 0.02).bfloat16()
    dw = torch.empty(nw, kw, device="cuda", dtype=torch.float32)
    dx = torch.empty(T, kw, device="cuda", dtype=torch.bfloat16)
    return dy, x, w, dw, dx


def wgrad(dy, x, dw):      # dW[nw, kw] = dY^T X : A = dY stored [K=T, M=nw], B = X stored [K=T, N=kw]
    ops.gemm(dy, x, a_major=1, b_major=1, epilogue=EPI_F32, out=dw)


def dgrad(dy, w, dx):      # dX[T, kw] = dY W : A = dY [M=T, K=nw], B = W stored [K=nw, N=kw]
    ops.gemm(dy, w, a_major=0, b_major=1, epilogue=EPI_BF16, out=dx)


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
    return e0.elapsed_time(e1) / iters * 1e3


lib = _lib.lib()
sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
for name, nw, kw in LAYERS:
    dy, x, w, dw, dx = make(nw, kw)
    ref_dw = dy.float().t() @ x.float()
    ref_dx = dy.float() @ w.float()

    def seq():
        wgrad(dy, x, dw)
        dgrad(dy, w, dx)

    def conc():
        cur = torch.cuda.current_stream()
        fork = torch.cuda.Event()
        fork.record(cur)
        sa.wait_event(fork)
        sb.wait_event(fork)
        with torch.cuda.stream(sa):
            wgrad(dy, x, dw)
        with torch.cuda.stream(sb):
            dgrad(dy, w, dx)
        cur.wait_stream(sa)
        cur.wait_stream(sb)

    row = {"layer": name, "flop": 4.0 * T * nw * kw}
    lib.b200b_set_sm_limit(0)
    row["seq_us"] = round(timed(seq), 2)
    row["conc_full_limit_us"] = round(timed(conc), 2)
    lib.b200b_set_sm_limit(74)
    row["conc_half_limit_us"] = round(timed(conc), 2)
    row["seq_half_limit_us"] = round(timed(seq), 2)
    lib.b200b_set_sm_limit(0)
    torch.cuda.synchronize()
    row["err_dw"] = float((dw - ref_dw).abs().max() / ref_dw.abs().max())
    row["err_dx"] = float((dx.float() - ref_dx).abs().max() / ref_dx.abs().max())
    row["seq_tflops"] = round(row["flop"] / row["seq_us"] / 1e6)
    row["best_conc_tflops"] = round(row["flop"] / min(row["conc_full_limit_us"], row["conc_half_limit_us"]) / 1e6)
    print(json.dumps(row), flush=True)
