"""Does building DecodeStepGraphs slow later eager steps? Times an eager C2 fwd+bwd step before the graphs
exist, while they exist, and after they are deleted."""
import gc
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

from vlm_bridge_b200 import BridgeLite, DecodeStepGraphs, VisionKVCache

torch.manual_seed(0)
m = BridgeLite(dropout=0.1).cuda().train()
v = torch.randn(8, 257, 1024).cuda()
t = torch.randn(8, 128, 2304).cuda()
params = list(m.parameters())


def step():
    m._w16_key = None
    for p in params:
        p.grad = None
    m(v, t).float().square().mean().backward()


def timed(tag):
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        step()
    e1.record()
    torch.cuda.synchronize()
    print(tag, round(e0.elapsed_time(e1) / 20, 3), "ms/step; reserved GB", round(torch.cuda.memory_reserved() / 2**30, 2), flush=True)


timed("before graphs")
m.eval()
vis = torch.randn(32, 257, 1024).cuda()
txt = torch.randn(32, 64, 2304).cuda()
with torch.no_grad():
    cache = VisionKVCache(m, vis)
graphs = DecodeStepGraphs(m, cache)
with torch.no_grad():
    for s in range(1, 65):
        graphs(txt[:, :s])
torch.cuda.synchronize()
m.train()
timed("with 64 decode graphs alive")
del graphs
gc.collect()
torch.cuda.empty_cache()
timed("after deleting the graphs")
