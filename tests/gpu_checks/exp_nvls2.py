"""Round-2 microbenchmark of the NVLS all-reduce kernel alone (torchrun, >= 2 ranks): bf16 bucket averaged in
place, time per (grid shape, units in flight per thread, exclusive SMs). The exchange is bound by bytes in flight
(one multimem.ld_reduce round trip ~5 us): GB/s should scale with blocks x threads x unroll until the links fill.
Run: torchrun --nproc-per-node N tests/gpu_checks/exp_nvls2.py > gpurun_out/exp_nvls2.jsonl"""
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

n = 64 << 20  # bf16 elements (128 MiB)
red = GradBucketReducer(backend="nvls", grad_dtype=torch.bfloat16, timeout_s=30)
arena32, arena16 = red.arenas(n, n + 4096, dev)
red._post = torch.cuda.current_stream()
nv = red._nvls
torch.cuda.synchronize(); dist.barrier()
out = []
SHAPES = [(32, 512, 4, False), (32, 512, 8, False), (32, 512, 16, False), (32, 1024, 8, False), (16, 1024, 16, False),
          (64, 512, 8, False), (148, 256, 8, False), (148, 512, 8, False), (148, 1024, 8, False), (148, 1024, 16, False),
          (4, 1024, 16, True), (8, 1024, 16, True), (16, 1024, 16, True), (4, 1024, 8, True)]
for mb in (40, 128):
    elems = mb << 19
    for blocks, threads, unroll, excl in SHAPES:
        red.nvls_blocks, red.nvls_threads, red.nvls_unroll, red.exclusive_sms = blocks, threads, unroll, excl
        arena16[:elems].fill_(float(rank + 1))

        def go():
            red._launch_nvls(nv["off16"], 2 * elems, True)

        go()
        torch.cuda.synchronize(); dist.barrier()
        ok = bool((arena16[:elems].float() == (world + 1) / 2.0).all().item())     # mean of 1..world
        for _ in range(2):
            go()
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            go()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        out.append({"MB": mb, "blocks": blocks, "threads": threads, "unroll": unroll, "exclusive": excl,
                    "us": round(ms * 1e3, 1), "algbw_GBs": round(mb * 1.048576e-3 / (ms * 1e-3), 1), "correct": ok,
                    "MB_in_flight": round(blocks * threads * unroll * 16 / 2 ** 20, 2)})
        if rank == 0:
            print(json.dumps(out[-1]), flush=True)
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
    if rank == 0:
        print(json.dumps({"MB": mb, "nccl": True, "us": round(ms * 1e3, 1), "algbw_GBs": round(mb * 1.048576e-3 / (ms * 1e-3), 1)}), flush=True)
red.check_errors()
dist.barrier()
dist.destroy_process_group()
