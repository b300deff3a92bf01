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
    ok = True
    for dtype in (torch.float32, torch.bfloat16):
        g = torch.Generator().manual_seed(100 + rank)
        mine = torch.randn(lay.total, generator=g)
        if dtype == torch.bfloat16:                      # weight gradients are produced in bf16 in this mode
            mine[:lay.n_weights] = mine[:lay.n_weights].bfloat16().float()
        arena32 = mine.clone()
        arena16 = mine[:lay.n_weights].bfloat16() if dtype == torch.bfloat16 else None
        if arena16 is not None:
            arena32[:lay.n_weights] = float("nan")       # must be filled by the reducer's conversion
        red = GradBucketReducer(bucket_bytes=8000, grad_dtype=dtype)   # forces several buckets per slab
        red.begin(arena32, arena16, lay.n_weights)
        # the order BridgeLite._run_backward announces ranges: per block (last first) the six weight
        # matrices from the end of the slab to its start, then the block's vector slab; K/V last
        for i in reversed(range(lay.nb)):
            pre = f"bridge_blocks.{i}."
            names = ["ffn.3.weight", "ffn.0.weight", "self_attention.w_o.weight", "self_attention.w_q.weight",
                     "cross_attention.w_o.weight", "cross_attention.w_q.weight"]
            sizes = {"self_attention.w_q.weight": 3 * 64 * 64}
            for n in names:
                o = lay.offsets[pre + n]
                e = o + sizes.get(n, {"ffn.3.weight": 64 * 256, "ffn.0.weight": 256 * 64}.get(n, 64 * 64))
                red.weights_ready(o, e)
            red.flush()
            red.vectors_ready(lay.block_v_start[i], lay.block_v_end[i])
        red.weights_ready(lay.kv_w_start, lay.block_w_start[0])
        red.flush()
        red.vectors_ready(lay.kv_b_start, lay.block_v_start[0])
        red.finish()
        others = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(others, mine)
        want = torch.stack(others).mean(0)
        tol = 1e-6 if dtype == torch.float32 else 2e-2   # bf16 sum of two bf16 values, then /2
        ok = ok and torch.allclose(arena32, want, atol=tol, rtol=tol) and not torch.isnan(arena32).any()
        esize = 4 if dtype == torch.float32 else 2
        ok = ok and red.bytes_per_step == esize * lay.n_weights + 4 * (lay.total - lay.n_weights)
        ok = ok and red.buckets_per_step > 4
    covered = torch.zeros(lay.total, dtype=torch.bool)
    for s, e in lay.buckets():
        assert not covered[s:e].any()                   # slabs are disjoint
        covered[s:e] = True
    ok = ok and bool(covered.all())
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


def _lazy_worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from vlm_bridge_b200.bridge import _Layout
    from vlm_bridge_b200.parallel import GradBucketReducer

    lay = _Layout(2, 64, 32, 256)
    g = torch.Generator().manual_seed(200 + rank)
    mine = torch.randn(lay.total, generator=g)
    mine[:lay.n_weights] = mine[:lay.n_weights].bfloat16().float()
    arena32 = mine.clone()
    arena32[:lay.n_weights] = float("nan")            # must stay untouched: the weights are NOT materialised as fp32
    arena16 = mine[:lay.n_weights].bfloat16()
    red = GradBucketReducer(bucket_bytes=8000, grad_dtype=torch.bfloat16, materialize_fp32=False)
    res = {"backend": red.backend, "lazy": not red.materialize_fp32}
    red.begin(arena32, arena16, lay.n_weights)
    for s, e in lay.buckets():
        if s < lay.n_weights:
            red.weights_ready(s, e)
            red.flush()
        else:
            red.vectors_ready(s, e)
    red.finish()
    red.check_errors()                                 # no nvls transport here: nothing to report, must not raise
    others = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(others, mine)
    want = torch.stack(others).mean(0)
    res["weights16_averaged"] = bool(torch.allclose(arena16.float(), want[:lay.n_weights], atol=2e-2, rtol=2e-2))
    res["weights32_untouched"] = bool(torch.isnan(arena32[:lay.n_weights]).all())
    res["vectors32_averaged"] = bool(torch.allclose(arena32[lay.n_weights:], want[lay.n_weights:], atol=1e-6, rtol=1e-6))
    res["bytes"] = red.bytes_per_step == 2 * lay.n_weights + 4 * (lay.total - lay.n_weights)
    # an fp32 exchange always materialises (there is nothing lazy about it)
    res["f32_forces_materialize"] = GradBucketReducer(grad_dtype=torch.float32, materialize_fp32=False).materialize_fp32
    # the diagnostics switch launches nothing and leaves every arena as it was
    red2 = GradBucketReducer(bucket_bytes=8000, grad_dtype=torch.bfloat16)
    red2.diag_skip_exchange = True
    a32, a16 = mine.clone(), mine[:lay.n_weights].bfloat16()
    red2.begin(a32, a16, lay.n_weights)
    red2.weights_ready(0, lay.n_weights)
    red2.finish()
    res["skip_exchange_is_a_no_op"] = bool(torch.equal(a32, mine) and torch.equal(a16, mine[:lay.n_weights].bfloat16()))
    out[rank] = res
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_lazy_bf16_arena_mode_and_argument_rules_world2():
    """materialize_fp32=False: the averaged weight gradients stay in the bf16 arena (what BridgeAdamW reads), the fp32
    arena keeps only the bias / LayerNorm gradients; unroll / backend arguments are validated."""
    sys.path.insert(0, ROOT)
    world = 2
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_lazy_worker, args=(world, _free_port(), out), nprocs=world, join=True)
        res = dict(out)
    for r in (0, 1):
        assert res[r] == {"backend": "nccl", "lazy": True, "weights16_averaged": True, "weights32_untouched": True,
                          "vectors32_averaged": True, "bytes": True, "f32_forces_materialize": True,
                          "skip_exchange_is_a_no_op": True}, res[r]


def test_reducer_argument_validation():
    sys.path.insert(0, ROOT)
    from vlm_bridge_b200.parallel import GradBucketReducer

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(_free_port())
    dist.init_process_group("gloo", rank=0, world_size=1)
    try:
        with pytest.raises(ValueError):
            GradBucketReducer(nvls_unroll=5)
        with pytest.raises(ValueError):
            GradBucketReducer(backend="ring")
        with pytest.raises(ValueError):
            GradBucketReducer(grad_dtype=torch.float16)
        with pytest.raises(RuntimeError):
            GradBucketReducer(backend="nvls")            # one rank: no exchange partner
        r = GradBucketReducer()                           # auto on gloo / one rank: the torch.distributed transport
        assert r.backend == "nccl" and r.timeout_s >= 1 and (r.nvls_blocks, r.nvls_threads, r.nvls_unroll) == (16, 1024, 8)
    finally:
        dist.destroy_process_group()
