"""Data-parallel training of the bridge: one process per GPU, batch sharded across ranks, only the
bridge gradients exchanged (the frozen encoders carry no gradients).

The reference has no distributed code at all, so this is new code beside its training loop, not a
mirror of any reference file. The exchange step is a bucketed all-reduce (average) over the flat
gradient arena that `BridgeLite` writes during backward:

* weight-gradient matrices are announced one by one, in the order backward finishes them (the C
  entry point calls back after enqueueing each weight-gradient GEMM); adjacent ranges are merged
  into buckets of at least `bucket_bytes` and every bucket's all-reduce is launched at once on
  NCCL's stream, so it runs under the remaining backward kernels;
* `grad_dtype=torch.bfloat16` (default): the weight-gradient GEMMs write bf16 -- the rounding the
  reference's autocast applies to these gradients anyway (the gradient of a bf16-cast weight is a
  bf16 tensor) -- the buckets travel as bf16 (316 MB instead of 633 MB per step) and a side stream
  converts each averaged bucket into the fp32 `.grad` arena as soon as its all-reduce is done.
  `grad_dtype=torch.float32` exchanges the fp32 arena in place (no conversion, twice the bytes);
* bias / LayerNorm gradients (0.1 % of the bytes) stay fp32 and go out as one bucket at the end;
* `finish()` joins the side stream before autograd hands the gradients to the optimizer path
  (GradScaler.unscale_, clip_grad_norm_, AdamW).

Two transports:
* `backend="nvls"` (default when the process group's GPUs offer NVSwitch multicast): the library's
  own all-reduce kernel (csrc/allreduce_nvls.cu) on symmetric gradient buffers -- in-switch
  reduction (multimem.ld_reduce) and multicast store (multimem.st) of the bf16 buckets, ~1x the
  bucket per link direction instead of a ring's 2(N-1)/N, followed by a bf16 -> fp32 pass into the
  `.grad` arena. Its CTAs use no shared memory and few registers, so they share SMs with the backward
  kernels. Measured alternatives kept as options: an fp32 multicast straight into `.grad`
  (`fp32_multicast`) and SMs reserved for the exchange (`exclusive_sms`); see `enable_data_parallel`.
  torch.distributed's symmetric memory is used only to allocate and map the buffers.
* `backend="nccl"`: torch.distributed.all_reduce per bucket (also what the CPU/gloo tests drive).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch
import torch.distributed as dist

__all__ = ["GradBucketReducer", "enable_data_parallel", "disable_data_parallel", "broadcast_parameters"]


class GradBucketReducer:
    """Driven by `BridgeLite._run_backward`: begin() -> weights_ready()* / flush() / vectors_ready()* -> finish()."""

    def __init__(self, process_group: Optional[dist.ProcessGroup] = None, bucket_bytes: int = 32 << 20,
                 grad_dtype: torch.dtype = torch.bfloat16, backend: str = "auto", nvls_blocks: int = 16,
                 nvls_threads: int = 1024, exclusive_sms: bool = False, fp32_multicast: bool = False,
                 nvls_unroll: int = 8, materialize_fp32: bool = True, timeout_s: Optional[int] = None):
        if grad_dtype not in (torch.bfloat16, torch.float32):
            raise ValueError("grad_dtype must be torch.bfloat16 or torch.float32")
        if backend not in ("auto", "nvls", "nccl"):
            raise ValueError("backend must be 'auto', 'nvls' or 'nccl'")
        self.group = process_group
        self.world_size = dist.get_world_size(process_group)
        self.bucket_bytes = int(bucket_bytes)
        self.grad_dtype = grad_dtype
        self.wgrad_bf16 = grad_dtype == torch.bfloat16 and self.world_size > 1
        self._nccl = dist.get_backend(process_group) == "nccl"
        if backend == "auto":
            # measured on B200 (profiles/r02_dp_sweep6_2gpu.log, _sweep7_4gpu.log, _sweep8_8gpu.log): between two GPUs
            # NCCL's point-to-point all-reduce disturbs the backward least (1.98 ms per step against 2.11 for the NVLS
            # kernel); from four GPUs on the in-switch reduction wins (2.00 against 2.17 at N = 4, 2.01 at N = 8)
            backend = "nvls" if (self._nccl and self.world_size > 2 and _nvls_available()) else "nccl"
        if backend == "nvls" and not (2 <= self.world_size <= 8):
            raise RuntimeError("the nvls transport needs 2..8 ranks on one NVSwitch domain")
        self.backend = backend
        self.nvls_blocks = int(nvls_blocks)
        self.nvls_threads = int(nvls_threads)
        self.exclusive_sms = bool(exclusive_sms) and self.nvls_blocks % 2 == 0
        self.fp32_multicast = bool(fp32_multicast)   # False: bf16 result in place + a separate bf16 -> fp32 pass
        if nvls_unroll not in (4, 8, 16):
            raise ValueError("nvls_unroll must be 4, 8 or 16")
        self.nvls_unroll = int(nvls_unroll)
        # False (bf16 exchange only): the averaged weight gradients STAY in the bf16 arena -- no bf16 -> fp32 pass
        # (0.95 GB of HBM traffic per step), the weights' `.grad` are not set; `BridgeAdamW` reads the arena directly
        # and `BridgeLite.materialize_grads()` produces fp32 `.grad` on demand
        self.materialize_fp32 = bool(materialize_fp32) or not self.wgrad_bf16
        # a rank may legitimately be seconds late (checkpoint write, validation, first-step initialisation): the
        # exchange waits as long as the process group would
        if timeout_s is None:
            try:
                timeout_s = int(dist.distributed_c10d._get_default_timeout(dist.get_backend(process_group)).total_seconds())
            except Exception:  # noqa: BLE001
                timeout_s = 600
        self.timeout_s = max(1, int(timeout_s))
        self._err_word: Optional[torch.Tensor] = None
        self.diag_skip_convert = False  # diagnostics only: leaves .grad of the weights unwritten (timing what the pass costs)
        self.diag_skip_exchange = False  # diagnostics only: no collective is launched (gradients stay local, in the same arenas)
        self._nvls = None               # (comm struct, symmetric byte buffer, flag buffer, handles)
        self.trace: Optional[list] = None   # set to [] to record per-bucket CUDA events (diagnostics)
        self._t0 = None
        self._epoch = 0
        self._post: Optional[torch.cuda.Stream] = None
        self._arena32: Optional[torch.Tensor] = None
        self._arena16: Optional[torch.Tensor] = None
        self._pending: Optional[tuple[int, int]] = None
        self._vec: Optional[tuple[int, int]] = None
        self._works: list = []
        self.bytes_reduced = 0          # since construction
        self.bytes_per_step = 0         # of the last begin()..finish()
        self.buckets_per_step = 0

    def trace_report(self) -> list[dict]:
        """After a traced backward (+ synchronize): per bucket, when its producer finished on the compute
        stream, when the exchange started / ended on the side stream (ms since backward began)."""
        out = []
        for lo, hi, nbytes, ev in self.trace or []:
            out.append({"MB": round(nbytes / 2 ** 20, 1), "ready_ms": round(self._t0.elapsed_time(ev[0]), 3),
                        "start_ms": round(self._t0.elapsed_time(ev[1]), 3), "end_ms": round(self._t0.elapsed_time(ev[2]), 3)})
        return out

    def describe(self) -> str:
        kind = "bf16 buckets + fp32 vectors" if self.wgrad_bf16 else "fp32 buckets"
        how = (f"own NVLS multimem kernel ({self.nvls_blocks} CTAs x {self.nvls_threads} threads"
               f"{' on SMs of their own' if self.exclusive_sms else ''}"
               f", {self.nvls_unroll} x 16 B in flight per thread"
               f"{', fp32 multicast into .grad' if self.fp32_multicast else (', bf16 in place + fp32 pass' if self.materialize_fp32 else ', bf16 in place, no fp32 pass (optimizer reads the bf16 arena)')})"
               if self.backend == "nvls" else ("NCCL" if self._nccl else dist.get_backend(self.group)))
        return (f"{how} all-reduce(avg), {kind}, >= {self.bucket_bytes >> 20} MiB per bucket, "
                f"{self.buckets_per_step} collectives per step")

    # -- nvls transport: symmetric buffers ---------------------------------------------------------
    def arenas(self, n_weights: int, total: int, device: torch.device):
        """(fp32 gradient arena [total], weight-gradient arena [n_weights] or None) the backward must
        write for this transport, or (None, None) if ordinary tensors will do (nccl). Both live in one
        symmetric allocation mapped on every rank: the fp32 arena is where `.grad` ends up (the exchange
        multicasts the averaged values into it), the bf16 arena is what the weight-gradient GEMMs write.
        Allocated once; reused by every step."""
        if self.backend != "nvls":
            return None, None
        esize = 2 if self.wgrad_bf16 else 4
        nv = self._nvls
        if nv is None or nv["n_weights"] != n_weights or nv["total"] != total or nv["esize"] != esize:
            nv = self._nvls = self._nvls_setup(n_weights, total, esize, device)
        return nv["arena32"], nv["weights16"]

    def _nvls_setup(self, n_weights: int, total: int, esize: int, device: torch.device) -> dict:
        import torch.distributed._symmetric_memory as symm

        from . import _lib

        group = self.group if self.group is not None else dist.group.WORLD
        bytes32 = (total * 4 + 255) // 256 * 256
        bytes16 = (n_weights * 2 + 255) // 256 * 256 if esize == 2 else 0
        buf = symm.empty(bytes32 + bytes16, dtype=torch.uint8, device=device)
        hbuf = symm.rendezvous(buf, group.group_name)
        nflag = _lib.lib().b200b_allreduce_nvls_flag_bytes() // 4
        flags = symm.empty(nflag, dtype=torch.int32, device=device)
        flags.zero_()
        hflags = symm.rendezvous(flags, group.group_name)
        if not hbuf.multicast_ptr:
            raise RuntimeError("symmetric memory has no multicast address on this system: use backend='nccl'")
        torch.cuda.synchronize(device)
        dist.barrier(group=self.group)            # every rank's flags are zero before anyone signals
        comm = _lib.NvlsComm()
        comm.multicast_base = hbuf.multicast_ptr
        comm.local_base = buf.data_ptr()
        for q in range(self.world_size):
            comm.flags[q] = hflags.buffer_ptrs[q]
        comm.rank, comm.world = dist.get_rank(self.group), self.world_size
        # pinned host word the kernel writes when a barrier wait expires (read by check_errors without a sync)
        self._err_word = torch.zeros(1, dtype=torch.int32).pin_memory()
        comm.timeout_s = self.timeout_s
        comm.error_word = self._err_word.data_ptr()
        return dict(comm=comm, buf=buf, flags=flags, handles=(hbuf, hflags), n_weights=n_weights, total=total,
                    esize=esize, mc=int(hbuf.multicast_ptr), off16=bytes32,
                    epoch_dev=torch.zeros(1, dtype=torch.int32, device=device),
                    arena32=buf[:total * 4].view(torch.float32),
                    weights16=buf[bytes32:bytes32 + n_weights * 2].view(torch.bfloat16) if esize == 2 else None)

    def _launch_nvls(self, byte_offset: int, nbytes: int, bf16: bool, out_multicast: int = 0) -> None:
        from . import _lib

        # collective number = (index within this step) + device counter advanced once per step by
        # finish(): identical for eager launches and for replays of a captured CUDA graph
        self._epoch += 1
        flags = ((_lib.NVLS_OUT_MULTICAST if out_multicast else 0) | (_lib.NVLS_EXCLUSIVE_SMS if self.exclusive_sms else 0)
                 | (self.nvls_unroll << 8))
        _lib.check(_lib.lib().b200b_allreduce_nvls(
            C.byref(self._nvls["comm"]), 0 if bf16 else 1, byte_offset, nbytes, 1.0 / self.world_size,
            C.c_void_p(out_multicast) if out_multicast else None, self._epoch, self._nvls["epoch_dev"].data_ptr(),
            self.nvls_blocks, self.nvls_threads, flags, self._post.cuda_stream),
            "allreduce_nvls")

    def check_errors(self) -> None:
        """Raises if an earlier collective of the nvls transport gave up waiting for a rank (`timeout_s`). The
        kernel reports that through a pinned host word, so this costs no synchronisation; it is called at the
        start and end of every backward, i.e. the failure surfaces one step late at the latest."""
        w = self._err_word
        if w is not None and int(w[0]) != 0:
            code = int(w[0])
            raise RuntimeError(f"gradient exchange (nvls): rank {dist.get_rank(self.group)} waited more than {self.timeout_s} s "
                               f"for rank {(code & 0xff) - 1} (phase {(code >> 8) & 0xf}, block {code >> 12}); the gradients "
                               "of that step are invalid. A rank is stalled or dead; raise `timeout_s` if ranks may "
                               "legitimately skew that long")

    # -- protocol ------------------------------------------------------------------------------------
    def begin(self, arena32: torch.Tensor, arena16: Optional[torch.Tensor], n_weights: int) -> None:
        self.check_errors()
        self._arena32, self._arena16, self._n_weights = arena32, arena16, n_weights
        self._pending, self._vec, self._works = None, None, []
        self._epoch = 0
        self.bytes_per_step = 0
        self.buckets_per_step = 0
        if arena32.is_cuda and self._post is None:
            # highest priority: a bucket's exchange CTAs are placed before the next compute kernel's
            self._post = torch.cuda.Stream(device=arena32.device, priority=-1)
        if self.trace is not None:
            self.trace.clear()
            self._t0 = torch.cuda.Event(enable_timing=True)
            self._t0.record()

    def weights_ready(self, start: int, end: int) -> None:
        """Elements [start, end) of the weight region are final (their kernel is enqueued)."""
        if self.world_size == 1 or end <= start:
            return
        if self._pending is None:
            self._pending = (start, end)
        elif end == self._pending[0]:
            self._pending = (start, self._pending[1])
        elif start == self._pending[1]:
            self._pending = (self._pending[0], end)
        else:
            self.flush()
            self._pending = (start, end)
        esize = self._arena16.element_size() if self._arena16 is not None else 4
        if (self._pending[1] - self._pending[0]) * esize >= self.bucket_bytes:
            self.flush()

    def flush(self) -> None:
        if self._pending is None:
            return
        lo, hi = self._pending
        self._pending = None
        src = self._arena16 if self._arena16 is not None else self._arena32
        self._launch(src[lo:hi], lo, hi, convert=src.dtype == torch.bfloat16 and self.materialize_fp32)

    def vectors_ready(self, start: int, end: int) -> None:
        """fp32 bias / LayerNorm gradient ranges; merged and sent as one bucket by finish()."""
        if self.world_size == 1 or end <= start:
            return
        self._vec = (start, end) if self._vec is None else (min(start, self._vec[0]), max(end, self._vec[1]))

    def finish(self) -> None:
        """Send what is left, then make the current stream wait for every bucket."""
        self.flush()
        if self._vec is not None:
            lo, hi = self._vec
            self._vec = None
            self._launch(self._arena32[lo:hi], lo, hi, convert=False)
        if self._post is not None:
            if self.backend == "nvls" and self._epoch:
                with torch.cuda.stream(self._post):
                    self._nvls["epoch_dev"].add_(self._epoch)      # after this step's last collective
            torch.cuda.current_stream().wait_stream(self._post)
        self._works.clear()
        self._arena32 = self._arena16 = None
        self.check_errors()

    # -- one bucket ----------------------------------------------------------------------------------
    def _launch(self, chunk: torch.Tensor, lo: int, hi: int, convert: bool) -> None:
        nbytes = chunk.numel() * chunk.element_size()
        if self.diag_skip_exchange:
            self.buckets_per_step += 1
            return
        if self.backend == "nvls":
            self.bytes_reduced += nbytes
            self.bytes_per_step += nbytes
            self.buckets_per_step += 1
            nv = self._nvls
            self._post.wait_stream(torch.cuda.current_stream())
            if self.trace is not None:
                ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
                ev[0].record(torch.cuda.current_stream())
                ev[1].record(self._post)
                self.trace.append((lo, hi, nbytes, ev))
            with torch.cuda.stream(self._post):
                if lo >= self._n_weights or chunk.dtype == torch.float32:
                    # fp32 ranges of the symmetric .grad arena (bias / LayerNorm gradients, or everything
                    # with grad_dtype=float32): averaged in place
                    self._launch_nvls(4 * lo, (nbytes + 15) // 16 * 16, False)
                elif self.fp32_multicast and convert:
                    # bf16 bucket: reduced in the switch, broadcast as fp32 into every rank's .grad arena
                    self._launch_nvls(nv["off16"] + 2 * lo, nbytes, True, nv["mc"] + 4 * lo)
                else:
                    # bf16 bucket averaged in place (half the broadcast bytes on the links), then one
                    # HBM-bound pass turns it into the fp32 .grad
                    from . import _lib

                    self._launch_nvls(nv["off16"] + 2 * lo, nbytes, True)
                    if convert and not self.diag_skip_convert:
                        _lib.check(_lib.lib().b200b_bf16_to_f32(chunk.data_ptr(), self._arena32[lo:hi].data_ptr(), hi - lo,
                                                                1.0, self._post.cuda_stream), "bf16_to_f32")
            if self.trace is not None:
                self.trace[-1][3][2].record(self._post)
            return
        op = dist.ReduceOp.AVG if self._nccl else dist.ReduceOp.SUM
        work = dist.all_reduce(chunk, op=op, group=self.group, async_op=True)
        self.bytes_reduced += nbytes
        self.bytes_per_step += nbytes
        self.buckets_per_step += 1
        if chunk.is_cuda:
            from . import _lib

            self._post.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self._post):
                work.wait()                                  # the side stream waits; the host does not
                if not self._nccl:
                    chunk.div_(self.world_size)
                if convert:
                    _lib.check(_lib.lib().b200b_bf16_to_f32(chunk.data_ptr(), self._arena32[lo:hi].data_ptr(), hi - lo, 1.0,
                                                            self._post.cuda_stream), "bf16_to_f32")
            self._works.append((work, chunk))
        else:                                                # CPU tensors (gloo): host-side logic tests
            work.wait()
            if not self._nccl:
                chunk.div_(self.world_size)
            if convert:
                self._arena32[lo:hi].copy_(chunk.float())


def broadcast_parameters(module: torch.nn.Module, src: int = 0, process_group=None) -> None:
    """Every rank starts from rank `src`'s parameters (one broadcast of the flat buffer when the
    module has been flattened, else one per tensor)."""
    flat = getattr(module, "_flat", None)
    if flat is not None:
        dist.broadcast(flat, src=src, group=process_group)
        module._w16_key = None
        return
    for p in module.parameters():
        dist.broadcast(p.data, src=src, group=process_group)


def _nvls_available() -> bool:
    """True when torch's symmetric memory (the allocator / mapper the nvls transport uses) is there;
    whether the fabric offers a multicast address is only known after the first rendezvous."""
    try:
        import torch.distributed._symmetric_memory as symm  # noqa: F401
    except Exception:  # noqa: BLE001
        return False
    return torch.cuda.is_available()


def enable_data_parallel(module, process_group=None, bucket_bytes: int = 32 << 20,
                         grad_dtype: torch.dtype = torch.bfloat16, backend: str = "auto",
                         nvls_blocks: int = 16, nvls_threads: int = 1024, exclusive_sms: bool = False,
                         fp32_multicast: bool = False, nvls_unroll: int = 8, materialize_fp32: bool = True,
                         timeout_s: Optional[int] = None) -> GradBucketReducer:
    """Attach a bucketed all-reduce to `module` (a B200 BridgeLite). Returns the reducer.

    Defaults are the fastest configuration measured on B200 (round 2: profiles/r02_dp_sweep6_2gpu.log,
    r02_dp_sweep7_4gpu.log, r02_dp_sweep8_8gpu.log): from four ranks on the own NVLS kernel with 16 CTAs x 1024
    threads x 8 units in flight sharing SMs with the backward (few, fat CTAs disturb fewer SMs: 2.01 ms per step at
    N = 8 against 2.18 for 32 x 512 x 4), between two ranks NCCL; bf16 result in place and a bf16 -> fp32 pass per
    bucket unless `materialize_fp32=False`. `fp32_multicast=True` broadcasts fp32 straight into `.grad` (no
    conversion pass, twice the broadcast bytes on the links); `exclusive_sms=True` sets `nvls_blocks`
    SMs aside for the exchange (CTA pairs claiming whole SMs, the persistent GEMMs limited to the
    rest via b200b_set_sm_limit until `disable_data_parallel`) -- interference drops to ~+10 % but a
    handful of SMs cannot issue multimem requests fast enough (~13 GB/s per SM at 4 units in flight per thread;
    `nvls_unroll` 8 / 16 deepens that).

    `materialize_fp32=False` (bf16 exchange): the averaged weight gradients stay in the bf16 arena, the weights'
    `.grad` stay None, `BridgeAdamW` consumes the arena directly (same arithmetic: the fp32 `.grad` would hold
    exactly these bf16 values) and `module.materialize_grads()` produces fp32 `.grad` tensors on demand -- for a
    loop that needs them every step (GradScaler.unscale_, clip_grad_norm_, torch.optim.AdamW) keep the default.

    `timeout_s`: how long a collective waits for a late rank before the step fails with a RuntimeError
    (default: the process group's timeout). The transport cannot wait forever the way a host-side NCCL watchdog
    does, because a spinning kernel cannot be cancelled; it reports through a pinned host word instead of
    trapping, so the CUDA context survives."""
    if not dist.is_initialized():
        raise RuntimeError("torch.distributed is not initialised")
    reducer = GradBucketReducer(process_group, bucket_bytes, grad_dtype, backend, nvls_blocks, nvls_threads,
                                exclusive_sms, fp32_multicast, nvls_unroll, materialize_fp32, timeout_s)
    module._bucket_hook = reducer
    if reducer.backend == "nvls" and reducer.exclusive_sms and reducer.world_size > 1 and torch.cuda.is_available():
        from . import _lib

        sms = torch.cuda.get_device_properties(torch.cuda.current_device()).multi_processor_count
        _lib.lib().b200b_set_sm_limit(sms - reducer.nvls_blocks)
    return reducer


def disable_data_parallel(module) -> None:
    reducer = getattr(module, "_bucket_hook", None)
    module._bucket_hook = None
    if reducer is not None and reducer.backend == "nvls" and reducer.exclusive_sms and torch.cuda.is_available():
        from . import _lib

        _lib.lib().b200b_set_sm_limit(0)
