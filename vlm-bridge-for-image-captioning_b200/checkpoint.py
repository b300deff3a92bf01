"""Bridge-only checkpoint IO in the reference's two on-disk formats (SURVEY.md 8b "on-disk contract",
8f rank 4).

Format B -- training checkpoints, `training_orchestrator.save_checkpoint / load_checkpoint`
(training_orchestrator.py:104-194): a `torch.save`d dict with `epoch` (already +1),
`model_state_dict` (the bridge tensors under the `bridge_module.` prefix they have inside
`FullModel`, :114-121), `optimizer_state_dict`, `best_val_loss`, `config`, optional
`scaler_state_dict` / `scheduler_state_dict`, `early_stopping_counter`; written as
`latest_checkpoint.pth`, and for a new best also `best_model.pth` and the weights-only
`best_model_weights_only.pth` (:137-156).
Format A -- `FullModel.save_model / load_model` (full_model.py:443-472): `bridge_module_state_dict`
with bare keys plus `model_config`.

Files written here load in the reference (`torch.load` + `load_state_dict(strict=True)`) and files
the reference wrote load here: same keys, fp32 tensors, same optimizer `state_dict` layout
(`BridgeAdamW` keeps torch.optim.AdamW's).

What differs from the reference is how the bytes leave the GPU: the parameters (and AdamW moments)
live in flat arenas, so a snapshot is one device-to-host copy per arena into pinned memory, enqueued
on the current stream (the training loop continues at once); the pickling and the three file writes
run on a background thread from that host snapshot, each file written to a temporary name and
renamed, so a crash never leaves a truncated checkpoint. Under data parallelism every rank holds the
same state and only rank 0 writes.
"""
from __future__ import annotations

import os
import pickle
import threading
from collections import OrderedDict
from typing import Any, Dict, Optional

import torch

__all__ = ["save_checkpoint", "load_checkpoint", "save_model", "load_model", "CheckpointWriter",
           "BRIDGE_PREFIX"]

BRIDGE_PREFIX = "bridge_module."     # the attribute name of the bridge inside FullModel (full_model.py:68)


def _is_rank0() -> bool:
    import torch.distributed as dist

    return not (dist.is_available() and dist.is_initialized()) or dist.get_rank() == 0


class _Snapshot:
    """Host copy of the bridge parameters (and optimizer moments) taken asynchronously."""

    def __init__(self):
        self._pinned: Dict[str, torch.Tensor] = {}
        self.event: Optional[torch.cuda.Event] = None

    def _to_host(self, name: str, t: torch.Tensor) -> torch.Tensor:
        if t.device.type != "cuda":
            return t.detach().clone()
        buf = self._pinned.get(name)
        if buf is None or buf.shape != t.shape or buf.dtype != t.dtype:
            buf = self._pinned[name] = torch.empty(t.shape, dtype=t.dtype, device="cpu", pin_memory=True)
        buf.copy_(t.detach(), non_blocking=True)
        return buf

    def take(self, bridge, optimizer):
        """-> (bridge state_dict on the host, optimizer state_dict on the host or None). Tensors of a
        flattened bridge are views of ONE host buffer per arena."""
        flat = getattr(bridge, "_flat", None)
        named = bridge._named_params() if hasattr(bridge, "_named_params") else list(bridge.named_parameters())
        flattened = flat is not None and all(
            p.data_ptr() == flat.data_ptr() + 4 * bridge._layout.offsets[n] for n, p in named)
        sd = OrderedDict()
        if flattened:
            host = self._to_host("param", flat)
            for n, p in named:
                o = bridge._layout.offsets[n]
                sd[n] = host[o:o + p.numel()].view(p.shape)
        else:
            for n, t in bridge.state_dict().items():
                sd[n] = self._to_host("param/" + n, t)
        osd = None
        if optimizer is not None:
            osd = optimizer.state_dict()
            m_arena, v_arena = getattr(optimizer, "_m", None), getattr(optimizer, "_v", None)
            hosts = {}
            if flattened and m_arena is not None and v_arena is not None:
                hosts = {"exp_avg": (m_arena, self._to_host("exp_avg", m_arena)),
                         "exp_avg_sq": (v_arena, self._to_host("exp_avg_sq", v_arena))}
            state = {}
            for idx, st in osd["state"].items():
                new = {}
                for k, v in st.items():
                    if torch.is_tensor(v) and k in hosts and v.device.type == "cuda":
                        arena, host = hosts[k]
                        o = (v.data_ptr() - arena.data_ptr()) // 4
                        if 0 <= o and o + v.numel() <= arena.numel():
                            new[k] = host[o:o + v.numel()].view(v.shape)
                            continue
                    new[k] = self._to_host(f"opt/{idx}/{k}", v) if torch.is_tensor(v) else v
                state[idx] = new
            osd = {"state": state, "param_groups": osd["param_groups"]}
        if torch.cuda.is_available() and any(p.device.type == "cuda" for _, p in named):
            self.event = torch.cuda.Event()
            self.event.record()
        return sd, osd


def _atomic_save(obj: Any, path: str) -> None:
    tmp = f"{path}.tmp.{os.getpid()}"
    torch.save(obj, tmp)
    os.replace(tmp, path)


class CheckpointWriter:
    """Background writer: `submit` returns as soon as the device-to-host copies are enqueued; `wait()`
    blocks until the files of the last submit are on disk (and re-raises a failure of the writer)."""

    def __init__(self):
        self._snap = _Snapshot()
        self._thread: Optional[threading.Thread] = None
        self._error: Optional[BaseException] = None

    def wait(self) -> None:
        if self._thread is not None:
            self._thread.join()
            self._thread = None
        if self._error is not None:
            err, self._error = self._error, None
            raise err

    def submit(self, build_files, bridge, optimizer, asynchronous: bool) -> None:
        self.wait()                              # the pinned buffers are reused: one write at a time
        sd, osd = self._snap.take(bridge, optimizer)
        event = self._snap.event

        def work():
            try:
                if event is not None:
                    event.synchronize()
                for path, obj in build_files(sd, osd):
                    _atomic_save(obj, path)
            except BaseException as e:  # noqa: BLE001
                self._error = e

        if asynchronous:
            self._thread = threading.Thread(target=work, name="b200b-checkpoint", daemon=False)
            self._thread.start()
        else:
            work()
            self.wait()


_default_writer = CheckpointWriter()


def save_checkpoint(checkpoint_dir: str, bridge, optimizer=None, *, epoch: int, best_val_loss: float = float("inf"),
                    config: Optional[dict] = None, scaler=None, scheduler=None, early_stopping_counter: int = 0,
                    is_best: bool = False, asynchronous: bool = True,
                    writer: Optional[CheckpointWriter] = None) -> Optional[CheckpointWriter]:
    """Format B, same files and keys as `training_orchestrator.save_checkpoint`
    (training_orchestrator.py:104-156); `epoch` is the 0-based epoch just finished (stored +1, :114).
    Returns the writer (call `.wait()` before reading the files or exiting), or None on ranks != 0."""
    if not _is_rank0():
        return None
    os.makedirs(checkpoint_dir, exist_ok=True)
    meta = {"epoch": int(epoch) + 1, "best_val_loss": best_val_loss, "config": dict(config or {})}
    extra = {}
    if scaler is not None:
        extra["scaler_state_dict"] = scaler.state_dict()
    if scheduler is not None:
        extra["scheduler_state_dict"] = scheduler.state_dict()
    extra["early_stopping_counter"] = int(early_stopping_counter)

    def build_files(sd, osd):
        ckpt = {"epoch": meta["epoch"],
                "model_state_dict": OrderedDict((BRIDGE_PREFIX + k, v) for k, v in sd.items()),
                "optimizer_state_dict": osd, "best_val_loss": meta["best_val_loss"], "config": meta["config"]}
        ckpt.update(extra)
        files = [(os.path.join(checkpoint_dir, "latest_checkpoint.pth"), ckpt)]
        if is_best:
            files.append((os.path.join(checkpoint_dir, "best_model.pth"), ckpt))
            files.append((os.path.join(checkpoint_dir, "best_model_weights_only.pth"),
                          {"model_state_dict": ckpt["model_state_dict"], "config": ckpt["config"]}))
        return files

    w = writer or _default_writer
    w.submit(build_files, bridge, optimizer, asynchronous)
    return w


def _torch_load(path: str, map_location, trusted: bool):
    try:
        return torch.load(path, map_location=map_location, weights_only=True)
    except pickle.UnpicklingError as e:
        # reference checkpoints pickle `config.__dict__`, which may hold arbitrary objects; the
        # reference itself loads them with the unrestricted unpickler (training_orchestrator.py:166)
        if not trusted:
            raise RuntimeError(f"{path} holds objects the safe unpickler rejects ({e}); pass trusted=True to load "
                               "it the way the reference does (only for files you wrote yourself)") from e
        return torch.load(path, map_location=map_location, weights_only=False)


def _bridge_state(ckpt: dict) -> "OrderedDict[str, torch.Tensor]":
    if "bridge_module_state_dict" in ckpt:                                    # Format A
        return OrderedDict(ckpt["bridge_module_state_dict"])
    if "model_state_dict" in ckpt:                                            # Format B
        sd = OrderedDict((k[len(BRIDGE_PREFIX):], v) for k, v in ckpt["model_state_dict"].items()
                         if k.startswith(BRIDGE_PREFIX))
        if not sd:                                                            # bare keys (saved from a bare bridge)
            sd = OrderedDict(ckpt["model_state_dict"])
        return sd
    raise KeyError("not a bridge checkpoint: neither 'model_state_dict' (training_orchestrator.py:114) nor "
                   "'bridge_module_state_dict' (full_model.py:452)")


def load_checkpoint(path: str, bridge, optimizer=None, *, scaler=None, scheduler=None, map_location=None,
                    trusted: bool = False) -> dict:
    """Restore a Format B (or A) file into `bridge` (strict) and, when present and asked for, the
    optimizer / scaler / scheduler (training_orchestrator.py:159-194). Returns the training state the
    reference restores into its context: `start_epoch`, `best_val_loss`, `early_stopping_counter`, `config`."""
    if map_location is None:
        map_location = next(bridge.parameters()).device
    ckpt = _torch_load(path, map_location, trusted)
    bridge.load_state_dict(_bridge_state(ckpt), strict=True)
    if optimizer is not None and ckpt.get("optimizer_state_dict") is not None:
        optimizer.load_state_dict(ckpt["optimizer_state_dict"])
    if scaler is not None and "scaler_state_dict" in ckpt:
        scaler.load_state_dict(ckpt["scaler_state_dict"])
    if scheduler is not None and "scheduler_state_dict" in ckpt:
        scheduler.load_state_dict(ckpt["scheduler_state_dict"])
    return {"start_epoch": ckpt.get("epoch", 0), "best_val_loss": ckpt.get("best_val_loss", float("inf")),
            "early_stopping_counter": ckpt.get("early_stopping_counter", 0), "config": ckpt.get("config", {})}


def save_model(path: str, bridge, model_config: Optional[dict] = None, asynchronous: bool = False,
               writer: Optional[CheckpointWriter] = None) -> Optional[CheckpointWriter]:
    """Format A, as `FullModel.save_model` (full_model.py:443-461)."""
    if not _is_rank0():
        return None
    cfg = dict(model_config or {"vision_dim": bridge.vision_dim, "language_dim": bridge.language_dim})

    def build_files(sd, _osd):
        return [(path, {"bridge_module_state_dict": sd, "model_config": cfg})]

    w = writer or _default_writer
    w.submit(build_files, bridge, None, asynchronous)
    return w


def load_model(path: str, bridge, map_location=None, trusted: bool = False) -> None:
    """Format A (or B) weights into `bridge`, as `FullModel.load_model` (full_model.py:463-472)."""
    if map_location is None:
        map_location = next(bridge.parameters()).device
    bridge.load_state_dict(_bridge_state(_torch_load(path, map_location, trusted)), strict=True)
