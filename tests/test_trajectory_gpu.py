"""BASELINE.json north_star: "loss within 1e-3 over 100 steps", at the REAL widths (1024 / 2304, 2 blocks,
heads 8 / 18; config C1: batch 2, 257 vision tokens, 64 text positions).

The yardstick is tests/golden/trajectory_c1.json: 100 training steps of the UNMODIFIED reference module
(imported in the build container by tests/golden/make_trajectory_golden.py, ~10 CPU-minutes) with the
reference's update rule (clip_grad_norm_ 0.3 then AdamW, core_training_loop.py:84-104), once under
`torch.autocast(bfloat16)` -- the reference's training numerics -- and once in plain fp32.

Here: the CUDA loop (bridge fwd+bwd kernels + the fused clip/AdamW of `BridgeAdamW`) from the same weights on the
same batches. Asserted per step: |loss - reference autocast loss| <= 1e-3 * max(1, loss), and <= 2e-3 against
the fp32 curve. The 2-rank variant shards every batch over two GPUs (one sample each) with the gradient
exchange on, and must reproduce the same curve (the mean of the two rank losses is the batch loss).
"""
import json
import os
import socket
import sys

import pytest
import torch

from oracle import bridge_oracle as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden", "trajectory_c1.json")


def _golden():
    with open(GOLDEN) as f:
        return json.load(f)


def _batches(cfg):
    g = torch.Generator().manual_seed(cfg["batch_seed"])
    return [(torch.randn(cfg["batch"], cfg["len_vision"], 1024, generator=g),
             torch.randn(cfg["batch"], cfg["len_text"], 2304, generator=g)) for _ in range(cfg["n_batches"])]


def _check(got, cfg, tag):
    ref16, ref32 = torch.tensor(cfg["autocast_bf16"]), torch.tensor(cfg["fp32"])
    assert float(ref32[-1]) < 0.2 * float(ref32[0])                       # the run really trains
    dev16 = float(((got - ref16).abs() / ref16.clamp_min(1.0)).max())
    dev32 = float(((got - ref32).abs() / ref32.clamp_min(1.0)).max())
    print(f"{tag}: loss {float(got[0]):.5f} -> {float(got[-1]):.5f}; max |d| vs reference autocast {dev16:.2e}, "
          f"vs reference fp32 {dev32:.2e}")
    out_dir = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out_dir):
        with open(os.path.join(out_dir, f"trajectory_{tag}.json"), "w") as f:
            json.dump({"loss": got.tolist(), "max_dev_vs_autocast_bf16": dev16, "max_dev_vs_fp32": dev32}, f)
    assert dev16 <= 1e-3, (dev16, dev32)
    assert dev32 <= 2e-3, (dev16, dev32)


@pytest.mark.timeout(600)
def test_loss_trajectory_100_steps_real_widths_matches_reference():
    from vlm_bridge_b200 import BridgeAdamW, BridgeLite

    cfg = _golden()
    sd0 = O.init_state_dict(cfg["weight_seed"])
    m = BridgeLite(dropout=0.0)
    m.load_state_dict(sd0, strict=True)
    m = m.cuda().train()
    opt = BridgeAdamW(m, lr=cfg["lr"], weight_decay=cfg["weight_decay"], max_grad_norm=cfg["clip"])
    data = [(v.cuda(), t.cuda()) for v, t in _batches(cfg)]
    losses = []
    for s in range(cfg["steps"]):
        v, t = data[s % len(data)]
        opt.zero_grad(set_to_none=True)
        loss = m(v, t).float().square().mean()
        loss.backward()
        opt.step()
        losses.append(loss.detach())
    _check(torch.stack(losses).cpu(), cfg, "1gpu")


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _dp_worker(rank, world, port, out):
    import datetime

    import torch.distributed as dist

    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev, timeout=datetime.timedelta(seconds=120))
    from vlm_bridge_b200 import BridgeAdamW, BridgeLite
    from vlm_bridge_b200.parallel import broadcast_parameters, enable_data_parallel

    cfg = _golden()
    m = BridgeLite(dropout=0.0)
    m.load_state_dict(O.init_state_dict(cfg["weight_seed"]), strict=True)
    m = m.to(dev).train()
    per = cfg["batch"] // world
    data = [(v[rank * per:(rank + 1) * per].to(dev), t[rank * per:(rank + 1) * per].to(dev)) for v, t in _batches(cfg)]
    with torch.no_grad():
        m(*data[0])
    broadcast_parameters(m)
    enable_data_parallel(m)
    opt = BridgeAdamW(m, lr=cfg["lr"], weight_decay=cfg["weight_decay"], max_grad_norm=cfg["clip"])
    losses = []
    for s in range(cfg["steps"]):
        v, t = data[s % len(data)]
        opt.zero_grad(set_to_none=True)
        loss = m(v, t).float().square().mean()
        loss.backward()
        opt.step()
        l = loss.detach().clone()
        dist.all_reduce(l, op=dist.ReduceOp.AVG)
        losses.append(l)
    out[rank] = torch.stack(losses).cpu().tolist()
    # replicas stay in lock step: identical parameters after 100 exchanged steps
    flat = m._flat.clone()
    dist.broadcast(flat, src=0)
    out[f"same{rank}"] = bool(torch.equal(flat, m._flat))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(900)
def test_loss_trajectory_100_steps_two_rank_data_parallel():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp

    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_dp_worker, args=(2, _free_port(), out), nprocs=2, join=True)
        res = dict(out)
    assert res[0] == res[1] and res["same0"] and res["same1"]
    _check(torch.tensor(res[0]), _golden(), "2gpu_dp")
