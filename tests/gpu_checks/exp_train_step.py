"""Per-step wall/device time of fwd + bwd + BridgeAdamW.step at C2 (diagnostics for bench.py's train_step leg)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

from vlm_bridge_b200 import BridgeAdamW, BridgeLite

torch.manual_seed(0)
m = BridgeLite(dropout=0.1).cuda().train()
v = torch.randn(8, 257, 1024).cuda()
t = torch.randn(8, 128, 2304).cuda()
params = list(m.parameters())
opt = BridgeAdamW(m, lr=1e-5, weight_decay=0.01, max_grad_norm=0.3)
times = []
for i in range(30):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for p in params:
        p.grad = None
    t1 = time.perf_counter()
    loss = m(v, t).float().square().mean()
    t2 = time.perf_counter()
    loss.backward()
    t3 = time.perf_counter()
    opt.step()
    t4 = time.perf_counter()
    torch.cuda.synchronize()
    t5 = time.perf_counter()
    times.append([round((b - a) * 1e3, 2) for a, b in ((t0, t1), (t1, t2), (t2, t3), (t3, t4), (t4, t5), (t0, t5))])
for i in (0, 1, 2, 3, 4, 10, 20, 29):
    print(i, "zero/fwd/bwd/opt/sync/total ms:", times[i], flush=True)
