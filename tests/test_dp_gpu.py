"""Data-parallel gradient exchange on real GPUs (needs >= 2 devices; skipped otherwise).

Each rank runs fwd+bwd on its own batch shard with the reducer attached; the averaged gradients it
ends with must equal the average of the per-shard gradients computed without any exchange
(fp32 exchange: to reduction-order tolerance; bf16 exchange: to bf16 rounding of the buckets)."""
import os
import socket
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    import datetime

    import torch.distributed as dist

    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev,
                            timeout=datetime.timedelta(seconds=90))
    from vlm_bridge_b200 import BridgeLite
    from vlm_bridge_b200.parallel import broadcast_parameters, disable_data_parallel, enable_data_parallel

    torch.manual_seed(0)
    m = BridgeLite(dropout=0.0).to(dev).train()

    def shard(r):
        g = torch.Generator().manual_seed(1234 + r)
        return torch.randn(2, 33, 1024, generator=g).to(dev), torch.randn(2, 24, 2304, generator=g).to(dev)

    def grads(v, t):
        for p in m.parameters():
            p.grad = None
        m(v, t).float().square().mean().backward()
        if m._grad16 is not None:              # materialize_fp32=False: the 2-D weights carry no .grad yet
            assert all(p.grad is None for p in m.parameters() if p.dim() == 2)
            m.materialize_grads()
        return torch.cat([p.grad.reshape(-1) for p in m.parameters()]).clone()

    with torch.no_grad():
        m(*shard(rank))                        # flattens the parameters
    broadcast_parameters(m)
    want = torch.stack([grads(*shard(r)) for r in range(world)]).mean(0)      # no exchange
    res = {}
    for name, backend, dtype, tol, kw in (
            ("f32", "nccl", torch.float32, 1e-5, {}), ("bf16", "nccl", torch.bfloat16, 1.5e-2, {}),
            ("nvls_f32", "nvls", torch.float32, 1e-5, {}), ("nvls_bf16", "nvls", torch.bfloat16, 1.5e-2, {}),
            # fp32 multicast straight into .grad, and the exchange on 4 SMs of its own (CTA pairs)
            ("nvls_bf16_mc32", "nvls", torch.bfloat16, 1.5e-2, dict(fp32_multicast=True)),
            ("nvls_bf16_excl", "nvls", torch.bfloat16, 1.5e-2,
             dict(fp32_multicast=True, exclusive_sms=True, nvls_blocks=4, nvls_threads=1024)),
            # deeper per-thread pipelining; averaged gradients left in the bf16 arena and materialised on demand
            ("nvls_bf16_u16", "nvls", torch.bfloat16, 1.5e-2, dict(nvls_unroll=16, nvls_blocks=8, nvls_threads=1024)),
            ("nvls_bf16_lazy", "nvls", torch.bfloat16, 1.5e-2, dict(nvls_unroll=8, materialize_fp32=False)),
            ("nccl_bf16_lazy", "nccl", torch.bfloat16, 1.5e-2, dict(materialize_fp32=False))):
        red = enable_data_parallel(m, bucket_bytes=8 << 20, grad_dtype=dtype, backend=backend, **kw)
        got = grads(*shard(rank))
        torch.cuda.synchronize()
        err = float((got - want).norm() / want.norm())
        res[name] = (err, tol, red.buckets_per_step, red.bytes_per_step)
        got2 = grads(*shard(rank))             # a second step reuses the symmetric buffers and flags
        torch.cuda.synchronize()
        res[name + "_again"] = (float((got2 - want).norm() / want.norm()), tol, red.buckets_per_step, red.bytes_per_step)
        disable_data_parallel(m)
    out[rank] = res
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_dp_gradients_match_average_of_shards():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp

    world = 2
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
        res = dict(out)
    assert set(res) == {0, 1}
    for rank, r in res.items():
        for name, (err, tol, buckets, nbytes) in r.items():
            assert err <= tol, f"rank {rank} {name}: rel err {err} > {tol}"
            assert buckets >= 5
        assert r["bf16"][3] < 0.51 * r["f32"][3] + 4 * 200000
        assert r["nvls_bf16"][3] < 0.51 * r["nvls_f32"][3] + 4 * 200000


def _skew_worker(rank, world, port, out):
    import datetime
    import time

    import torch.distributed as dist

    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev,
                            timeout=datetime.timedelta(seconds=90))
    from vlm_bridge_b200 import BridgeLite
    from vlm_bridge_b200.parallel import broadcast_parameters, enable_data_parallel

    torch.manual_seed(0)
    m = BridgeLite(dropout=0.0).to(dev).train()
    g = torch.Generator().manual_seed(1234 + rank)
    v, t = torch.randn(1, 33, 1024, generator=g).to(dev), torch.randn(1, 24, 2304, generator=g).to(dev)
    with torch.no_grad():
        m(v, t)
    broadcast_parameters(m)

    def step():
        for p in m.parameters():
            p.grad = None
        m(v, t).float().square().mean().backward()
        torch.cuda.synchronize()

    res = {}
    # (1) a rank that is 3 s late is simply waited for when the limit is generous (the default is the process
    #     group's timeout): same gradients on both ranks afterwards, no error
    red = enable_data_parallel(m, backend="nvls", timeout_s=60)
    step()                                   # sets the symmetric buffers up (contains a host-side barrier)
    dist.barrier()
    if rank == 1:
        time.sleep(3.0)
    step()
    gsum = torch.cat([p.grad.reshape(-1) for p in m.parameters()]).double().sum()
    both = [torch.zeros_like(gsum) for _ in range(world)]
    dist.all_gather(both, gsum)
    res["late_rank_waited_for"] = bool(both[0] == both[1])
    red.check_errors()
    # (2) with a 1 s limit the waiting rank gives the collective up and REPORTS it: RuntimeError at the end of the
    #     backward (or at the next one), CUDA context intact
    red.timeout_s = 1
    red._nvls["comm"].timeout_s = 1
    dist.barrier()
    if rank == 1:
        time.sleep(4.0)
    err = None
    try:
        step()
        step()
    except RuntimeError as e:
        err = str(e)
    res["error"] = err
    torch.cuda.synchronize()                                  # the context survived
    res["context_alive"] = bool(torch.ones(4, device=dev).sum().item() == 4.0)
    out[rank] = res
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_rank_skew_is_waited_for_and_a_timeout_is_reported_not_trapped():
    """ADVICE (round 1): ranks skew by seconds around checkpoint writes / validation; the nvls transport must wait
    like NCCL does and, past its limit, raise instead of destroying the CUDA context."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp

    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_skew_worker, args=(2, _free_port(), out), nprocs=2, join=True)
        res = dict(out)
    assert res[0]["late_rank_waited_for"] and res[1]["late_rank_waited_for"]
    assert res[0]["error"] is not None and "waited more than 1 s for rank 1" in res[0]["error"], res
    assert res[0]["context_alive"] and res[1]["context_alive"]
