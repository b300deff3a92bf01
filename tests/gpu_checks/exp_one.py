"""Run one GEMM configuration a few times (target for an ncu capture)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from vlm_bridge_b200 import ops

M, N, K = 1024, 9216, 2304
epi = int(os.environ.get("EPI", "1"))
p = float(os.environ.get("P", "0.1"))
a = (torch.randn(M, K, device="cuda") * 0.5).bfloat16()
b = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
o = torch.empty(M, N, device="cuda", dtype=torch.float32 if epi in (2, 4) else torch.bfloat16)
kw = dict(epilogue=epi, out=o, dropout_p=p, seed=1234, dropout_stream=3)
if epi in (0, 1, 2):
    kw["bias"] = torch.randn(N, device="cuda")
if epi in (1, 3):
    kw["aux"] = torch.randn(M, N, device="cuda").bfloat16()
if epi == 2:
    kw["resid"] = torch.randn(M, N, device="cuda")
for _ in range(4):
    ops.gemm(a, b, **kw)
torch.cuda.synchronize()
print("ok")
