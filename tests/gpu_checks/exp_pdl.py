"""One C2 bridge step (eager) then a timed loop: used with B200B_PDL=<mask> to test programmatic
dependent launch per kernel family."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

from vlm_bridge_b200 import BridgeLite, GraphedBridgeStep

torch.manual_seed(0)
m = BridgeLite(dropout=0.1).cuda().train()
v = torch.randn(8, 257, 1024).cuda()
t = torch.randn(8, 128, 2304).cuda()
params = list(m.parameters())


def step():
    m._w16_key = None
    for p in params:
        p.grad = None
    loss = m(v, t).float().square().mean()
    loss.backward()
    return loss


for i in range(3):
    step()
    torch.cuda.synchronize()
    print("eager step", i, "ok", flush=True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(30):
    step()
e1.record()
torch.cuda.synchronize()
print("eager ms/step", round(e0.elapsed_time(e1) / 30, 4), flush=True)
if "--graph" in sys.argv:
    g = GraphedBridgeStep(m, lambda y: y.float().square().mean(), v, t)
    for _ in range(5):
        g.replay()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(50):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    print("graph ms/step", round(e0.elapsed_time(e1) / 50, 4), flush=True)
