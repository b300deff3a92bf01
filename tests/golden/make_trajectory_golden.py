"""Generate tests/golden/trajectory_c1.json: the loss curve of 100 training steps of the UNMODIFIED
reference `BridgeLite` at the real widths (vision 1024, language 2304, 2 blocks, heads 8 / 18; config C1:
batch 2, 257 vision tokens, 64 text positions), from `torch.manual_seed(0)` weights, on four cycled
synthetic batches, with the reference's own update rule (core_training_loop.py:84-104: clip_grad_norm_
then AdamW; clip 0.3 = config/training-default.yaml:9, weight decay 0.01 = training_setup.py:248-254).

Two curves: `fp32` (plain CPU fp32) and `autocast_bf16` (the reference's training numerics,
`torch.autocast("cpu", dtype=torch.bfloat16)` around the forward as core_training_loop.py:60-66 does on
CUDA). Dropout 0 (torch's dropout RNG stream cannot be reproduced from outside). lr 1e-4 instead of the
reference's 1e-5 so that the loss falls by a factor > 5 within the 100 steps and a wrong gradient or update
would show. loss = mean(y^2).

Run in the build container only (needs /root/reference; ~10 minutes of CPU):
    python tests/golden/make_trajectory_golden.py
tests/test_trajectory_gpu.py regenerates the same batches from the seeds recorded in the file.
"""
from __future__ import annotations

import importlib.util
import json
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/src/vlm_bridge/model_architecture/bridge_module.py"
CFG = dict(weight_seed=0, batch_seed=4242, n_batches=4, batch=2, len_vision=257, len_text=64, steps=100, lr=1e-4,
           weight_decay=0.01, clip=0.3)


def batches():
    g = torch.Generator().manual_seed(CFG["batch_seed"])
    return [(torch.randn(CFG["batch"], CFG["len_vision"], 1024, generator=g),
             torch.randn(CFG["batch"], CFG["len_text"], 2304, generator=g)) for _ in range(CFG["n_batches"])]


def run(ref, autocast: bool) -> list[float]:
    torch.manual_seed(CFG["weight_seed"])
    m = ref.BridgeLite(dropout=0.0).train()
    opt = torch.optim.AdamW(m.parameters(), lr=CFG["lr"], weight_decay=CFG["weight_decay"])
    data = batches()
    out = []
    for s in range(CFG["steps"]):
        v, t = data[s % len(data)]
        opt.zero_grad()
        if autocast:
            with torch.autocast("cpu", dtype=torch.bfloat16):
                y = m(v, t)
        else:
            y = m(v, t)
        loss = y.float().square().mean()
        loss.backward()
        torch.nn.utils.clip_grad_norm_(m.parameters(), CFG["clip"])
        opt.step()
        out.append(float(loss.detach()))
        print(("bf16" if autocast else "fp32"), s, out[-1], flush=True)
    return out


def main():
    sys.dont_write_bytecode = True
    spec = importlib.util.spec_from_file_location("ref_bridge_module", REF)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    res = dict(CFG)
    res["torch"] = torch.__version__
    res["fp32"] = run(ref, False)
    res["autocast_bf16"] = run(ref, True)
    with open(os.path.join(HERE, "trajectory_c1.json"), "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
