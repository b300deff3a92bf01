"""Microbenchmark of the NVLS all-reduce kernel alone (torchrun, >= 2 ranks): time per bucket size and
grid shape for the bf16 -> fp32-multicast exchange, next to NCCL on a bf16 buffer of the same size."""
import datetime
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import torch.distributed as dist

rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=60))
from vlm_bridge_b200.parallel import GradBucketReducer

n = 128 << 20  # bf16 elements (256 MiB)
red = GradBucketReducer(backend="nvls", grad_dtype=torch.bfloat16)
arena32, arena16 = red.arenas(n, n + 4096, dev)
arena16.fill_(1.0)
red._post = torch.cuda.current_stream()
nv = red._nvls
torch.cuda.synchronize(); dist.barrier()
out = []
for mb in (8, 32, 128, 256):
    elems = mb << 19
    for blocks, threads, excl in ((4, 1024, True), (2, 1024, True), (8, 1024, True), (4, 512, True), (32, 512, False),
                                  (148, 128, False)):
        red.nvls_blocks, red.nvls_threads, red.exclusive_sms = blocks, threads, excl

        def go():
            red._launch_nvls(nv["off16"], 2 * elems, True, nv["mc"])

        for _ in range(3):
            go()
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            go()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        ok = bool((arena32[:elems] == 1.0).all().item())
        out.append({"MB": mb, "blocks": blocks, "threads": threads, "exclusive": excl, "us": round(ms * 1e3, 1),
                    "algbw_GBs": round(mb * 1.048576e-3 / (ms * 1e-3), 1), "correct": ok})
    x = torch.ones(elems, device=dev, dtype=torch.bfloat16)
    for _ in range(3):
        dist.all_reduce(x, op=dist.ReduceOp.AVG)
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        dist.all_reduce(x, op=dist.ReduceOp.AVG)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    out.append({"MB": mb, "nccl": True, "us": round(ms * 1e3, 1), "algbw_GBs": round(mb * 1.048576e-3 / (ms * 1e-3), 1)})
if rank == 0:
    for o in out:
        print(json.dumps(o), flush=True)
dist.barrier()
dist.destroy_process_group()
