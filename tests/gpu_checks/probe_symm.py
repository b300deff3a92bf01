"""Probe (2+ ranks under torchrun): is torch symmetric memory / NVLS multicast usable on this box?"""
import datetime
import json
import os
import sys

import torch
import torch.distributed as dist

rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=60))
res = {"rank": rank, "world": world}
try:
    import torch.distributed._symmetric_memory as symm

    t = symm.empty(1 << 20, dtype=torch.float32, device=dev)
    h = symm.rendezvous(t, dist.group.WORLD.group_name)
    res["buffer_ptrs"] = [hex(p) for p in h.buffer_ptrs]
    res["signal_pad_ptrs"] = [hex(p) for p in h.signal_pad_ptrs]
    res["multicast_ptr"] = hex(h.multicast_ptr) if getattr(h, "multicast_ptr", 0) else 0
    res["signal_pad_size"] = getattr(h, "signal_pad_size", None)
    res["attrs"] = [a for a in dir(h) if not a.startswith("_")]
    t.fill_(rank + 1)
    h.barrier()
    peer = h.get_buffer((rank + 1) % world, (1 << 20,), torch.float32)
    res["peer_value"] = float(peer[0].item())
    h.barrier()
    if res["multicast_ptr"]:
        torch.ops.symm_mem.multimem_all_reduce_(t, "sum", dist.group.WORLD.group_name)
        torch.cuda.synchronize()
        res["multimem_allreduce_value"] = float(t[0].item())
except Exception as e:  # noqa: BLE001
    res["error"] = repr(e)[:400]
# NCCL all-reduce bandwidth at the bucket sizes the reducer uses
for dtype in (torch.bfloat16, torch.float32):
    for mb in (32, 128, 316 if dtype == torch.bfloat16 else 632):
        n = mb * (1 << 20) // (2 if dtype == torch.bfloat16 else 4)
        x = torch.ones(n, device=dev, dtype=dtype)
        for _ in range(3):
            dist.all_reduce(x, op=dist.ReduceOp.AVG)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            dist.all_reduce(x, op=dist.ReduceOp.AVG)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        res[f"nccl_{str(dtype)[6:]}_{mb}MB_ms"] = round(ms, 3)
        res[f"nccl_{str(dtype)[6:]}_{mb}MB_busbw_GBs"] = round(mb * 1.048576e-3 / (ms * 1e-3) * 2 * (world - 1) / world, 1)
print("PROBE " + json.dumps(res), flush=True)
dist.barrier()
dist.destroy_process_group()
