"""Fused optimizer step for the B200 bridge.

The reference updates the bridge with `torch.optim.AdamW(filter(requires_grad, model.parameters()),
lr, weight_decay, betas=(0.9, 0.999), eps=1e-8)` (training_setup.py:228-257) after
`scaler.unscale_`, a per-parameter `grad.norm(2).item()` loop, and `clip_grad_norm_`
(core_training_loop.py:84-104): several passes over 632 MB of gradients, ~4.4 GB of optimizer traffic
and 50+ host synchronisations per step.

`BridgeAdamW` performs the same arithmetic over the module's flat parameter / gradient arenas in two
kernel launches (b200b_grad_sqnorm, b200b_adamw_fused), with no host synchronisation, and also
writes the bf16 operand copy of the updated weights, so the next forward skips its re-cast pass.

It is a `torch.optim.Optimizer`: `param_groups`, `state_dict()` / `load_state_dict()` have
`torch.optim.AdamW`'s layout (per parameter `step`, `exp_avg`, `exp_avg_sq`; parameters indexed
0..51 in registration order), so checkpoints written by the reference's `save_checkpoint`
(training_orchestrator.py:104-156) load here and vice versa, LR schedulers work unchanged, and
`GradScaler.step(optimizer)` hands it `grad_scale` / `found_inf` as it does for torch's fused Adam.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib

__all__ = ["BridgeAdamW"]


class BridgeAdamW(torch.optim.Optimizer):
    """optimizer = BridgeAdamW(bridge, lr=1e-5, weight_decay=0.01, max_grad_norm=0.3)

    `max_grad_norm` (the reference's `gradient_clip_val`) folds `clip_grad_norm_` into the step; the
    pre-clip global norm of the last step is kept on the device in `last_grad_norm` (read it with
    `.item()` only when it is logged)."""

    _step_supports_amp_scaling = True   # GradScaler.step passes grad_scale / found_inf as attributes

    def __init__(self, bridge, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 1e-2,
                 max_grad_norm: Optional[float] = None):
        if lr < 0 or eps < 0 or not (0 <= betas[0] < 1) or not (0 <= betas[1] < 1) or weight_decay < 0:
            raise ValueError("invalid AdamW hyper-parameters")
        self.bridge = bridge
        params = [p for p in bridge.parameters() if p.requires_grad]
        if len(params) != len(list(bridge.parameters())):
            raise RuntimeError("BridgeAdamW updates the whole flat arena: every bridge parameter must require grad")
        defaults = dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay, amsgrad=False, maximize=False,
                        foreach=None, capturable=False, differentiable=False, fused=True, decoupled_weight_decay=True)
        super().__init__(params, defaults)
        self.max_grad_norm = max_grad_norm
        self.last_grad_norm: Optional[torch.Tensor] = None
        self._steps = 0                  # host view of the step count (exact once `_sync_steps` has run)
        self._step_dev = None            # device float[2] = {steps applied, steps skipped}: the authoritative count
        self._steps_dirty = False
        self.gather_steps = 0
        self._m = self._v = self._ws = self._norm2 = self._gflat = None

    # -- flat state ----------------------------------------------------------------------------------
    def _ensure_state(self) -> None:
        b = self.bridge
        b._ensure_flat()
        flat = b._flat
        if self._m is not None and self._m.device == flat.device and self._m.numel() == flat.numel():
            return
        lay = b._layout
        old = {p: dict(self.state[p]) for p in self.state}      # e.g. loaded from a checkpoint
        self._m = torch.zeros_like(flat)
        self._v = torch.zeros_like(flat)
        nbytes = _lib.lib().b200b_grad_sqnorm_workspace_bytes()
        self._ws = torch.zeros(nbytes, dtype=torch.uint8, device=flat.device)
        self._norm2 = torch.zeros(2, dtype=torch.float32, device=flat.device)
        skipped = 0.0 if self._step_dev is None else float(self._step_dev[1].item())
        self._step_dev = torch.zeros(2, dtype=torch.float32, device=flat.device)
        for name, p in b._named_params():
            o = lay.offsets[name]
            m = self._m[o:o + p.numel()].view(p.shape)
            v = self._v[o:o + p.numel()].view(p.shape)
            st = old.get(p)
            step = torch.tensor(float(self._steps))
            if st:
                m.copy_(st["exp_avg"])
                v.copy_(st["exp_avg_sq"])
                step = st["step"].detach().clone().cpu() if torch.is_tensor(st["step"]) else torch.tensor(float(st["step"]))
            self.state[p] = {"step": step, "exp_avg": m, "exp_avg_sq": v}
        steps = {int(self.state[p]["step"]) for p in self.state}
        if len(steps) > 1:
            raise RuntimeError("BridgeAdamW needs one common step count for all parameters")
        self._steps = steps.pop() if steps else 0
        self._step_dev.copy_(torch.tensor([float(self._steps), skipped]))
        self._steps_dirty = False

    def _sync_steps(self) -> None:
        """Read the device-side step count back (one host sync; only where a host value is needed:
        state_dict(), `skipped_steps`). A step that GradScaler or a non-finite gradient norm skipped has not
        advanced it -- the behaviour of torch's fused AdamW, which the bias corrections depend on."""
        if self._steps_dirty and self._step_dev is not None:
            self._steps = int(self._step_dev[0].item())
            self._steps_dirty = False

    @property
    def skipped_steps(self) -> int:
        """Steps that were not applied (found_inf from GradScaler, or a non-finite gradient norm). Syncs."""
        return 0 if self._step_dev is None else int(self._step_dev[1].item())

    def load_state_dict(self, state_dict) -> None:
        super().load_state_dict(state_dict)
        self._m = None                   # re-flatten the loaded moments on the next step
        ps = [p for g in self.param_groups for p in g["params"]]
        if ps and ps[0] in self.state:
            s = self.state[ps[0]]["step"]
            self._steps = int(s.item() if torch.is_tensor(s) else s)
            self._steps_dirty = False

    def state_dict(self):
        self._sync_steps()
        for st in self.state.values():
            if "step" in st:
                st["step"] = torch.tensor(float(self._steps))
        return super().state_dict()

    def _flat_grads(self):
        """(flat fp32 gradients in arena order, bf16 weight-gradient arena or None). The bridge's own gradient
        arena when every `.grad` is still the view autograd installed (the normal case), else a gathered copy.
        Data parallel with `materialize_fp32=False`: the 2-D weights have no `.grad`; their averaged gradients
        are the bf16 arena the exchange left behind, which the kernels read in place of the fp32 values."""
        b = self.bridge
        lay = b._layout
        named = b._named_params()
        arena = b._last_grad_arena
        g16 = b.__dict__.get("_grad16")
        if g16 is not None and arena is not None:
            base = arena.data_ptr()
            if all((lay.offsets[n] < lay.n_weights and p.grad is None) or
                   (p.grad is not None and p.grad.data_ptr() == base + 4 * lay.offsets[n]) for n, p in named):
                return arena, g16
            b.materialize_grads()       # someone replaced a .grad: fall back to ordinary fp32 gradients
        g0 = named[0][1].grad
        if g0 is None:
            raise RuntimeError("BridgeAdamW.step(): parameters have no gradients")
        base = g0.data_ptr() - 4 * lay.offsets[named[0][0]]
        if arena is not None and arena.data_ptr() == base and all(
                p.grad is not None and p.grad.dtype == torch.float32 and p.grad.data_ptr() == base + 4 * lay.offsets[n]
                for n, p in named):
            return arena, None
        self.gather_steps += 1          # diagnostics: steps that could not use the arena in place
        if self._gflat is None or self._gflat.device != g0.device:
            self._gflat = torch.empty(lay.total, device=g0.device, dtype=torch.float32)
        for n, p in named:
            if p.grad is None:
                raise RuntimeError(f"BridgeAdamW.step(): {n} has no gradient")
            o = lay.offsets[n]
            self._gflat[o:o + p.numel()].view(p.shape).copy_(p.grad)
        return self._gflat, None

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        self._ensure_state()
        b = self.bridge
        lay = b._layout
        group = self.param_groups[0]
        g, g16 = self._flat_grads()
        g16_ptr = None if g16 is None else g16.data_ptr()
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        lib = _lib.lib()
        grad_scale = getattr(self, "grad_scale", None)     # set by GradScaler.step for the duration of this call
        found_inf = getattr(self, "found_inf", None)
        _lib.check(lib.b200b_grad_sqnorm(g.data_ptr(), lay.total, g16_ptr, lay.n_weights, self._ws.data_ptr(),
                                         self._ws.numel(), self._norm2.data_ptr(), st), "grad_sqnorm")
        self._steps_dirty = True      # the device counter advances only if the kernel applies the update
        lr = group["lr"]
        _lib.check(lib.b200b_adamw_fused(
            b._flat.data_ptr(), g.data_ptr(), g16_ptr, self._m.data_ptr(), self._v.data_ptr(), b._w16.data_ptr(), lay.total,
            lay.n_weights, self._norm2.data_ptr(), float(self.max_grad_norm or 0.0),
            None if grad_scale is None else grad_scale.data_ptr(), None if found_inf is None else found_inf.data_ptr(),
            float(lr.item() if torch.is_tensor(lr) else lr), group["betas"][0], group["betas"][1], group["eps"],
            group["weight_decay"], 0, self._step_dev.data_ptr(), st), "adamw_fused")
        inv = 1.0 if grad_scale is None else 1.0 / grad_scale
        self.last_grad_norm = self._norm2[0].sqrt() * inv
        # the parameters changed in place behind torch's back: bump their version counters (stale K/V
        # caches, saved-tensor checks) and tell the module its bf16 operand copy is already current
        params = [p for _, p in b._named_params()]
        torch.autograd.graph.increment_version(params)
        b._w16_key = tuple(p._version for p in params)
        return loss
