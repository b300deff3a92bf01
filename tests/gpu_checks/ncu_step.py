"""Two eager C2 bridge steps for ncu captures of the step's kernels."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

from vlm_bridge_b200 import BridgeLite

torch.manual_seed(0)
m = BridgeLite(dropout=0.1).cuda().train()
v = torch.randn(8, 257, 1024).cuda()
t = torch.randn(8, 128, 2304).cuda()
for _ in range(2):
    m._w16_key = None
    for p in m.parameters():
        p.grad = None
    m(v, t).float().square().mean().backward()
torch.cuda.synchronize()
