"""Timing of one optimizer step over the 158.16 M bridge parameters: the reference's sequence
(GradScaler-free: per-parameter norm loop with .item(), clip_grad_norm_, torch.optim.AdamW) against
BridgeAdamW (b200b_grad_sqnorm + b200b_adamw_fused), and the HBM roofline of the fused pass."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

from vlm_bridge_b200 import BridgeAdamW, BridgeLite, _lib

torch.manual_seed(0)
m = BridgeLite(dropout=0.1).cuda().train()
v = torch.randn(8, 257, 1024).cuda()
t = torch.randn(8, 128, 2304).cuda()
m(v, t).float().square().mean().backward()
params = list(m.parameters())
n = sum(p.numel() for p in params)


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, (time.perf_counter() - t0) * 1e3 / reps


ref_params = [p.detach().clone().requires_grad_() for p in params]
for rp, p in zip(ref_params, params):
    rp.grad = p.grad.detach().clone()
ref_opt = torch.optim.AdamW(ref_params, lr=1e-5, weight_decay=0.01)


def ref_step():      # core_training_loop.py:84-104 without the GradScaler
    total = 0.0
    for p in ref_params:
        total += p.grad.data.norm(2).item() ** 2
    torch.nn.utils.clip_grad_norm_(ref_params, 0.3)
    ref_opt.step()


def ref_step_nolog():
    torch.nn.utils.clip_grad_norm_(ref_params, 0.3)
    ref_opt.step()


opt = BridgeAdamW(m, lr=1e-5, weight_decay=0.01, max_grad_norm=0.3)
fused_ms, fused_wall = timed(opt.step)
ref_ms, ref_wall = timed(ref_step)
ref2_ms, ref2_wall = timed(ref_step_nolog)
# per-kernel
_lib.profile_begin(torch.cuda.current_stream().cuda_stream)
for _ in range(5):
    opt.step()
ent = _lib.profile_end()
adam = sorted(ms for k, ms in ent if k == "adamw_fused")[len(ent) // 4]
norm = sorted(ms for k, ms in ent if k == "grad_sqnorm")[len(ent) // 4]
nw = m._layout.n_weights
bytes_adam = n * 28 + nw * 2
print(json.dumps({"params": n, "fused_step_ms": round(fused_ms, 4), "fused_wall_ms": round(fused_wall, 4),
                  "reference_sequence_ms": round(ref_ms, 4), "reference_sequence_wall_ms": round(ref_wall, 4),
                  "reference_without_norm_loop_ms": round(ref2_ms, 4), "adamw_fused_kernel_ms": round(adam, 4),
                  "adamw_fused_GBs": round(bytes_adam / adam * 1e-6, 1), "adamw_fused_frac_of_6464": round(bytes_adam / adam * 1e-6 / 6463.7, 3),
                  "grad_sqnorm_kernel_ms": round(norm, 4), "grad_sqnorm_GBs": round(n * 4 / norm * 1e-6, 1)}))
