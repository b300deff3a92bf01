"""Diagnostic (torchrun, 2+ ranks): how much does a concurrent gradient exchange slow the compute
kernels? A fixed compute loop (the FFN forward + weight-gradient GEMM pair, or an HBM-bound cast) is
timed on the compute stream while variants of the exchange run on a high-priority side stream."""
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
from vlm_bridge_b200 import _lib, ops
from vlm_bridge_b200.parallel import GradBucketReducer

n = 64 << 20  # bf16 elements (128 MiB)
red = GradBucketReducer(backend="nvls", grad_dtype=torch.bfloat16)
arena32, arena16 = red.arenas(n, n + 4096, dev)
nv = red._nvls
plain16 = torch.ones(n, device=dev, dtype=torch.bfloat16)
arena16.fill_(1.0)
side = torch.cuda.Stream(priority=-1)
red._post = side
T, D, F = 1024, 2304, 9216
x = torch.randn(T, D, device=dev).bfloat16()
w1 = torch.randn(F, D, device=dev).bfloat16()
wq = torch.randn(D, D, device=dev).bfloat16()
dy = torch.randn(T, D, device=dev).bfloat16()
h = torch.randn(T, F, device=dev).bfloat16()
o1 = torch.empty(T, F, device=dev, dtype=torch.bfloat16)
o2 = torch.empty(T, D, device=dev, dtype=torch.bfloat16)
gw = torch.empty(D, F, device=dev, dtype=torch.float32)
big = torch.randn(32 << 20, device=dev)
big16 = torch.empty(32 << 20, device=dev, dtype=torch.bfloat16)
SMS = torch.cuda.get_device_properties(dev).multi_processor_count


def compute_gemm():
    ops.gemm(x, w1, out=o1)                                             # fwd 1024x9216x2304
    ops.gemm(dy, h, a_major=1, b_major=1, epilogue=ops.EPI_F32, out=gw)  # wgrad 2304x9216x1024
    ops.gemm(x, wq, out=o2)                                             # 1024x2304x2304 (72 pair tiles)


def compute_cast():
    ops.cast_bf16(big, out=big16)


def comm_none(reps):
    pass


def mk_nvls(blocks, threads, excl, nbytes):
    def f(reps):
        red.nvls_blocks, red.nvls_threads, red.exclusive_sms = blocks, threads, excl
        for _ in range(reps):
            red._launch_nvls(nv["off16"], nbytes, True, nv["mc"])
    return f


def comm_nccl(reps):
    with torch.cuda.stream(side):
        for _ in range(reps):
            dist.all_reduce(plain16, op=dist.ReduceOp.AVG)


# (name, comm fn, reps, compute SM limit)
variants = [("none_148", comm_none, 0, 0), ("none_144", comm_none, 0, SMS - 4),
            ("nvls_4x1024_excl_limit144", mk_nvls(4, 1024, True, 2 * n), 10, SMS - 4),
            ("nvls_2x1024_excl_limit146", mk_nvls(2, 1024, True, 2 * n), 8, SMS - 2),
            ("nvls_4x1024_excl_nolimit", mk_nvls(4, 1024, True, 2 * n), 10, 0),
            ("nvls_4x1024_shared_limit144", mk_nvls(4, 1024, False, 2 * n), 10, SMS - 4),
            ("nvls_148x128_shared", mk_nvls(148, 128, False, 2 * n), 10, 0),
            ("nccl_128MB", comm_nccl, 12, 0)]
out = []
for cname, cfn, creps in (("gemm_trio", compute_gemm, 60), ("cast_128MB", compute_cast, 60)):
    for vname, vfn, vreps, limit in variants:
        _lib.lib().b200b_set_sm_limit(limit)
        for _ in range(3):
            cfn()
        torch.cuda.synchronize(); dist.barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record(side)
        vfn(vreps)
        s1.record(side)
        e0.record()
        for _ in range(creps):
            cfn()
        e1.record()
        torch.cuda.synchronize()
        out.append({"compute": cname, "comm": vname, "compute_us_per_iter": round(e0.elapsed_time(e1) / creps * 1e3, 1),
                    "comm_total_ms": round(s0.elapsed_time(s1), 2), "compute_total_ms": round(e0.elapsed_time(e1), 2),
                    "comm_GBs": round(vreps * 2 * n / (s0.elapsed_time(s1) * 1e-3) / 1e9, 1) if vreps else None})
        dist.barrier()
_lib.lib().b200b_set_sm_limit(0)
if rank == 0:
    for o in out:
        print(json.dumps(o), flush=True)
dist.barrier()
dist.destroy_process_group()
