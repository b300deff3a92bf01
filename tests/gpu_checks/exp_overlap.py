"""Diagnostic (torchrun, 2+ ranks): how much does a concurrent gradient exchange slow the compute
kernels, and which part of it does? A fixed compute loop (GEMMs, or an HBM-bound cast) is timed on the
compute stream while variants of communication work run on a side stream."""
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
arena16 = red.weight_arena(n, 1 << 16, dev)
arena32 = torch.empty(n, device=dev, dtype=torch.float32)
plain16 = torch.ones(n, device=dev, dtype=torch.bfloat16)
arena16.fill_(1.0)
side = torch.cuda.Stream()
red._post = side
red._arena32 = arena32
red._n_weights = n
T, D, F = 1024, 2304, 9216
x = torch.randn(T, D, device=dev).bfloat16()
w1 = torch.randn(F, D, device=dev).bfloat16()
dy = torch.randn(T, D, device=dev).bfloat16()
h = torch.randn(T, F, device=dev).bfloat16()
o1 = torch.empty(T, F, device=dev, dtype=torch.bfloat16)
gw = torch.empty(D, F, device=dev, dtype=torch.float32)
big = torch.randn(32 << 20, device=dev)
big16 = torch.empty(32 << 20, device=dev, dtype=torch.bfloat16)


def compute_gemm():
    ops.gemm(x, w1, out=o1)                                             # fwd 1024x9216x2304
    ops.gemm(dy, h, a_major=1, b_major=1, epilogue=ops.EPI_F32, out=gw)  # wgrad 2304x9216x1024


def compute_cast():
    ops.cast_bf16(big, out=big16)


def comm_none(reps):
    pass


def mk_nvls(blocks, threads, nbytes):
    def f(reps):
        red.nvls_blocks, red.nvls_threads = blocks, threads
        for _ in range(reps):
            red._launch_nvls(0, nbytes, True, 0)
    return f


def comm_convert(reps):
    for _ in range(reps):
        _lib.check(_lib.lib().b200b_bf16_to_f32(plain16.data_ptr(), arena32.data_ptr(), n, 1.0, side.cuda_stream), "cv")


def comm_nccl(reps):
    with torch.cuda.stream(side):
        for _ in range(reps):
            dist.all_reduce(plain16, op=dist.ReduceOp.AVG)


variants = [("none", comm_none, 0), ("nvls_32x512_128MB", mk_nvls(32, 512, 2 * n), 12),
            ("nvls_148x128_128MB", mk_nvls(148, 128, 2 * n), 12), ("nvls_8x512_128MB", mk_nvls(8, 512, 2 * n), 6),
            ("nvls_barrier_only_32x512", mk_nvls(32, 512, 16 * 1024), 400), ("convert_only", comm_convert, 60),
            ("nccl_128MB", comm_nccl, 12)]
out = []
for cname, cfn, creps in (("gemm_pair", compute_gemm, 60), ("cast_128MB", compute_cast, 60)):
    for vname, vfn, vreps in variants:
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
                    "comm_total_ms": round(s0.elapsed_time(s1), 2), "compute_total_ms": round(e0.elapsed_time(e1), 2)})
        dist.barrier()
if rank == 0:
    for o in out:
        print(json.dumps(o), flush=True)
dist.barrier()
dist.destroy_process_group()
