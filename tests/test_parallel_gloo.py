"""Host-side logic of the data-parallel gradient exchange on CPU: world_size 2, gloo backend.
The same GradBucketReducer object drives NCCL on the GPU box; here it reduces CPU arenas."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from vlm_bridge_b200.bridge import _Layout
    from vlm_bridge_b200.parallel import GradBucketReducer, broadcast_parameters

    lay = _Layout(2, 64, 32, 256)
    g = torch.Generator().manual_seed(100 + rank)
    arena = torch.randn(lay.total, generator=g)
    mine = arena.clone()
    red = GradBucketReducer(max_bucket_elems=5000)      # forces several chunks per slab
    for s, e in lay.buckets():                          # the order BridgeLite._run_backward fires them
        red(arena, s, e)
    red.finish()
    others = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(others, mine)
    want = torch.stack(others).mean(0)
    covered = torch.zeros(lay.total, dtype=torch.bool)
    for s, e in lay.buckets():
        assert not covered[s:e].any()                   # buckets are disjoint
        covered[s:e] = True
    ok = bool(covered.all()) and torch.allclose(arena, want, atol=1e-6) and red.bytes_reduced == 4 * lay.total
    # broadcast_parameters: every rank ends with rank 0's tensors
    lin = torch.nn.Linear(8, 8)
    broadcast_parameters(lin)
    ws = [torch.empty_like(lin.weight.data) for _ in range(world)]
    dist.all_gather(ws, lin.weight.data)
    ok = ok and all(torch.equal(w, ws[0]) for w in ws)
    out[rank] = ok
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_bucket_reducer_averages_every_slab_world2():
    world = 2
    port = _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
        assert dict(out) == {0: True, 1: True}


def test_layout_buckets_cover_the_arena_in_backward_order():
    sys.path.insert(0, ROOT)
    from vlm_bridge_b200.bridge import _Layout

    lay = _Layout(2, 2304, 1024, 9216)
    assert lay.total == 158160384
    b = lay.buckets()
    assert sum(e - s for s, e in b) == lay.total
    assert b[0] == (lay.block_w_start[1], lay.block_w_end[1])       # last block's weights ship first
    assert b[-2] == (lay.kv_w_start, lay.block_w_start[0])          # K/V weights (finished last) ship last
