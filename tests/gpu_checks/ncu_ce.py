"""Fused cross-entropy at the C2 loss shape ([1024, 256000] fp32 logits) for an `ncu --set full` capture:
two forward + backward rounds (the second is the one to read)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

from vlm_bridge_b200 import FusedCrossEntropyLoss

rows, L, V = 1024, 128, 256000
logits = torch.randn(rows, V, device="cuda").requires_grad_()
ids = torch.randint(3, V, (rows // L, L), device="cuda")
for _ in range(2):
    logits.grad = None
    FusedCrossEntropyLoss().forward_shifted(logits, ids).backward()
torch.cuda.synchronize()
