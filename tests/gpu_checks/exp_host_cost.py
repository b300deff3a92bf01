"""Where does the HOST time of an eagerly launched step go? cProfile over 30 eager fwd+bwd steps at C2 (the GPU is
never waited for inside the loop), plus wall-clock per phase. Run: python tests/gpu_checks/exp_host_cost.py"""
import cProfile
import io
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

from vlm_bridge_b200 import BridgeLite

torch.manual_seed(0)
m = BridgeLite(dropout=0.1).cuda().train()
g = torch.Generator().manual_seed(1)
v = torch.randn(8, 257, 1024, generator=g).cuda()
t = torch.randn(8, 128, 2304, generator=g).cuda()
params = list(m.parameters())


def step():
    for p in params:
        p.grad = None
    y = m(v, t)
    loss = y.float().square().mean()
    loss.backward()


for _ in range(5):
    step()
torch.cuda.synchronize()
n = 30
t0 = time.perf_counter()
for _ in range(n):
    step()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"host enqueue {1e3 * (t1 - t0) / n:.3f} ms/step; incl. drain {1e3 * (t2 - t0) / n:.3f} ms/step", flush=True)
# forward only / backward only
tf = tb = 0.0
for _ in range(n):
    for p in params:
        p.grad = None
    a = time.perf_counter()
    y = m(v, t)
    loss = y.float().square().mean()
    b = time.perf_counter()
    loss.backward()
    c = time.perf_counter()
    tf += b - a
    tb += c - b
torch.cuda.synchronize()
print(f"forward {1e3 * tf / n:.3f} ms, backward {1e3 * tb / n:.3f} ms (host)", flush=True)
pr = cProfile.Profile()
pr.enable()
for _ in range(n):
    step()
pr.disable()
torch.cuda.synchronize()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(18)
print(s.getvalue())
