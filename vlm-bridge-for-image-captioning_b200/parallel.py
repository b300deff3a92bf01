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
  own all-reduce kernel (csrc/allreduce_nvls.cu) on a symmetric gradient buffer -- in-switch
  reduction (multimem.ld_reduce / multimem.st), ~1x the bucket per link direction instead of a
  ring's 2(N-1)/N, fused with the bf16 -> fp32 conversion, and small enough (no shared memory) to
  share SMs with the backward GEMMs. torch.distributed's symmetric memory is used only to allocate
  and map the buffers.
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
                 grad_dtype: torch.dtype = torch.bfloat16, backend: str = "auto", nvls_blocks: int = 148,
                 nvls_threads: int = 128):
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
            backend = "nvls" if (self._nccl and self.world_size > 1 and _nvls_available()) else "nccl"
        if backend == "nvls" and not (2 <= self.world_size <= 8):
            raise RuntimeError("the nvls transport needs 2..8 ranks on one NVSwitch domain")
        self.backend = backend
        self.nvls_blocks = int(nvls_blocks)
        self.nvls_threads = int(nvls_threads)
        self._nvls = None               # (comm struct, symmetric byte buffer, flag buffer, handles)
        self.fuse_convert = False       # convert inside the exchange kernel (few CTAs) or as its own launch
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
        how = (f"own NVLS multimem kernel ({self.nvls_blocks} CTAs x {self.nvls_threads} threads)" if self.backend == "nvls"
               else ("NCCL" if self._nccl else dist.get_backend(self.group)))
        return (f"{how} all-reduce(avg), {kind}, >= {self.bucket_bytes >> 20} MiB per bucket, "
                f"{self.buckets_per_step} collectives per step")

    # -- nvls transport: symmetric buffers ---------------------------------------------------------
    def weight_arena(self, n_weights: int, n_vectors: int, device: torch.device) -> Optional[torch.Tensor]:
        """The arena the weight-gradient GEMMs must write for this transport, or None if any
        ordinary tensor will do (nccl). Allocated and mapped on every rank once."""
        if self.backend != "nvls":
            return None
        esize = 2 if self.wgrad_bf16 else 4
        if self._nvls is None or self._nvls["n_weights"] != n_weights or self._nvls["esize"] != esize:
            self._nvls = self._nvls_setup(n_weights, n_vectors, esize, device)
        return self._nvls["weights"]

    def _nvls_setup(self, n_weights: int, n_vectors: int, esize: int, device: torch.device) -> dict:
        import torch.distributed._symmetric_memory as symm

        from . import _lib

        group = self.group if self.group is not None else dist.group.WORLD
        wbytes = (n_weights * esize + 255) // 256 * 256
        vbytes = (n_vectors * 4 + 255) // 256 * 256
        buf = symm.empty(wbytes + vbytes, dtype=torch.uint8, device=device)
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
        return dict(comm=comm, buf=buf, flags=flags, handles=(hbuf, hflags), n_weights=n_weights, esize=esize,
                    epoch_dev=torch.zeros(1, dtype=torch.int32, device=device),
                    weights=buf[:n_weights * esize].view(torch.bfloat16 if esize == 2 else torch.float32),
                    vectors=buf[wbytes:wbytes + n_vectors * 4].view(torch.float32), voff=wbytes)

    def _launch_nvls(self, byte_offset: int, nbytes: int, bf16: bool, out_f32_ptr: int) -> None:
        from . import _lib

        # collective number = (index within this step) + device counter advanced once per step by
        # finish(): identical for eager launches and for replays of a captured CUDA graph
        self._epoch += 1
        _lib.check(_lib.lib().b200b_allreduce_nvls(
            C.byref(self._nvls["comm"]), 0 if bf16 else 1, byte_offset, nbytes, 1.0 / self.world_size,
            C.c_void_p(out_f32_ptr) if out_f32_ptr else None, self._epoch, self._nvls["epoch_dev"].data_ptr(),
            self.nvls_blocks, self.nvls_threads, self._post.cuda_stream),
            "allreduce_nvls")

    # -- protocol ------------------------------------------------------------------------------------
    def begin(self, arena32: torch.Tensor, arena16: Optional[torch.Tensor], n_weights: int) -> None:
        self._arena32, self._arena16, self._n_weights = arena32, arena16, n_weights
        self._pending, self._vec, self._works = None, None, []
        self._epoch = 0
        self.bytes_per_step = 0
        self.buckets_per_step = 0
        if arena32.is_cuda and self._post is None:
            self._post = torch.cuda.Stream(device=arena32.device)
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
        self._launch(src[lo:hi], lo, hi, convert=src.dtype == torch.bfloat16)

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

    # -- one bucket ----------------------------------------------------------------------------------
    def _launch(self, chunk: torch.Tensor, lo: int, hi: int, convert: bool) -> None:
        nbytes = chunk.numel() * chunk.element_size()
        if self.backend == "nvls":
            self.bytes_reduced += nbytes
            self.bytes_per_step += nbytes
            self.buckets_per_step += 1
            from . import _lib

            nv = self._nvls
            self._post.wait_stream(torch.cuda.current_stream())
            if self.trace is not None:
                ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
                ev[0].record(torch.cuda.current_stream())
                ev[1].record(self._post)
                self.trace.append((lo, hi, nbytes, ev))
            with torch.cuda.stream(self._post):
                if lo >= self._n_weights:             # fp32 vectors: stage through the symmetric buffer
                    n = hi - lo
                    nv["vectors"][:n].copy_(chunk)
                    self._launch_nvls(nv["voff"], (n * 4 + 15) // 16 * 16, False, 0)
                    chunk.copy_(nv["vectors"][:n])
                elif chunk.dtype == torch.bfloat16:   # bf16 bucket, averaged, then written as fp32 .grad
                    if self.fuse_convert:
                        self._launch_nvls(2 * lo, nbytes, True, self._arena32.data_ptr() + 4 * lo)
                    else:
                        # the exchange needs few CTAs (link bound), the conversion many (HBM bound)
                        self._launch_nvls(2 * lo, nbytes, True, 0)
                        _lib.check(_lib.lib().b200b_bf16_to_f32(chunk.data_ptr(), self._arena32[lo:hi].data_ptr(), hi - lo,
                                                                1.0, self._post.cuda_stream), "bf16_to_f32")
                else:                                 # fp32 bucket of the symmetric arena, in place, then copied out
                    self._launch_nvls(4 * lo, nbytes, False, 0)
                    self._arena32[lo:hi].copy_(chunk)
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
                         nvls_blocks: int = 148, nvls_threads: int = 128) -> GradBucketReducer:
    """Attach a bucketed all-reduce to `module` (a B200 BridgeLite). Returns the reducer."""
    if not dist.is_initialized():
        raise RuntimeError("torch.distributed is not initialised")
    reducer = GradBucketReducer(process_group, bucket_bytes, grad_dtype, backend, nvls_blocks, nvls_threads)
    module._bucket_hook = reducer
    return reducer


def disable_data_parallel(module) -> None:
    module._bucket_hook = None
