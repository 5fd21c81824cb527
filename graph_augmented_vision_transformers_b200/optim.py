"""The optimiser side of the reference's training step as libgvit multi-tensor kernels.

``/root/reference/src/training/trainer.py`` builds (``:47-56``) an AdamW over two parameter groups (the model; the loss's
mixing weights at 0.1 x lr), (``:77-87``) a per-step ``LambdaLR`` with linear warm-up followed by a cosine, and clips the
global gradient norm to 1.0 before every step (``:114-116``).  :class:`FusedAdamW` is that recipe with all state on the
device: one call = gradient norm + clip coefficient + schedule factor + AdamW update (``gvit_mt_adamw_step``), no host
synchronisation, capturable in a CUDA graph (``step.CapturedTrainStep``).

``state_dict`` / ``load_state_dict`` use ``torch.optim.AdamW``'s layout, so a reference checkpoint's
``optimizer_state_dict`` (trainer.py:188-197) loads here and vice versa; the scheduler state is the step counter.
"""
from __future__ import annotations

import math

import torch

from . import _lib
from .ops import _call, _check_cuda, _ptr, _stream

__all__ = ["FusedAdamW", "warmup_cosine_lambda"]


def warmup_cosine_lambda(step: int, warmup_steps: int, total_steps: int) -> float:
    """The reference's ``lr_lambda`` (trainer.py:81-85)."""
    if total_steps <= 0:
        return 1.0
    if step < warmup_steps:
        return float(step) / float(max(1, warmup_steps))
    progress = float(step - warmup_steps) / float(max(1, total_steps - warmup_steps))
    return 0.5 * (1.0 + math.cos(math.pi * progress))


class FusedAdamW:
    """``clip_grad_norm_`` + warm-up/cosine ``LambdaLR`` + ``AdamW`` in one device-side step.

    ``param_groups``: an iterable of parameters or of dicts ``{'params': [...], 'lr': ..., 'weight_decay': ...}`` exactly as
    ``torch.optim.AdamW`` takes them.  ``max_norm=None`` disables clipping; ``total_steps=0`` keeps the learning rate
    constant.  fp32 CUDA parameters only (the masters of an autocast model).
    """

    capturable = True

    def __init__(self, param_groups, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, *, max_norm=None,
                 warmup_steps=0, total_steps=0):
        groups = list(param_groups)
        if not groups:
            raise ValueError("FusedAdamW got an empty parameter list")
        if not isinstance(groups[0], dict):
            groups = [{"params": groups}]
        self.defaults = dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay)
        self.param_groups = []
        for g in groups:
            g = dict(g)
            g["params"] = list(g["params"])
            for k, v in self.defaults.items():
                g.setdefault(k, v)
            g.setdefault("capturable", True)
            self.param_groups.append(g)
        self.betas, self.eps = tuple(betas), float(eps)
        self.max_norm = None if max_norm is None else float(max_norm)
        self.warmup_steps, self.total_steps = int(warmup_steps), int(total_steps)
        self.params = [p for g in self.param_groups for p in g["params"]]
        if not self.params:
            raise ValueError("FusedAdamW got no parameters")
        _check_cuda(*self.params)
        for p in self.params:
            if p.dtype != torch.float32 or not p.is_contiguous():
                raise TypeError("FusedAdamW updates contiguous fp32 CUDA parameters (the masters of an autocast model)")
        dev = self.params[0].device
        self.device = dev
        self.exp_avg = [torch.zeros_like(p) for p in self.params]
        self.exp_avg_sq = [torch.zeros_like(p) for p in self.params]
        self.step_count = torch.zeros(1, dtype=torch.int64, device=dev)
        self.sched = torch.zeros(3, dtype=torch.float32, device=dev)      # norm, clip coefficient, lambda(step)
        self.tensor_steps = torch.zeros(len(self.params), dtype=torch.int32, device=dev)   # torch keeps `step` per parameter
        E = _lib.load().gvit_mt_chunk_elems()
        ct, ci = [], []
        for t, p in enumerate(self.params):
            for c in range((p.numel() + E - 1) // E):
                ct.append(t)
                ci.append(c)
        self.nchunks = len(ct)
        i64 = lambda xs: torch.tensor(xs, dtype=torch.int64, device=dev)
        self._p = i64([p.data_ptr() for p in self.params])
        self._m = i64([t.data_ptr() for t in self.exp_avg])
        self._v = i64([t.data_ptr() for t in self.exp_avg_sq])
        self._numel = i64([p.numel() for p in self.params])
        self._chunk_tensor = torch.tensor(ct, dtype=torch.int32, device=dev)
        self._chunk_index = torch.tensor(ci, dtype=torch.int32, device=dev)
        self._partial = torch.empty(self.nchunks, dtype=torch.float32, device=dev)
        self._g_host = torch.zeros(len(self.params), dtype=torch.int64).pin_memory()
        self._g = torch.zeros(len(self.params), dtype=torch.int64, device=dev)
        self._g_cached = None
        self._lr = self._wd = None
        self._refresh_hyper()

    # -- hyper-parameters live in param_groups (as torch's do); the device tables follow them ----------------------------
    def _refresh_hyper(self):
        lr = [float(g["lr"]) for g in self.param_groups for _ in g["params"]]
        wd = [float(g["weight_decay"]) for g in self.param_groups for _ in g["params"]]
        if self._lr is None or (lr, wd) != self._hyper_cached:
            self._lr = torch.tensor(lr, dtype=torch.float32, device=self.device)
            self._wd = torch.tensor(wd, dtype=torch.float32, device=self.device)
            self._hyper_cached = (lr, wd)

    def zero_grad(self, set_to_none: bool = True):
        for p in self.params:
            if p.grad is not None:
                if set_to_none:
                    p.grad = None
                else:
                    p.grad.zero_()

    @torch.no_grad()
    def step(self):
        """One optimisation step over the current ``.grad`` tensors (parameters without a gradient are skipped)."""
        ptrs = []
        for p in self.params:
            g = p.grad
            if g is None:
                ptrs.append(0)
                continue
            if g.dtype != torch.float32 or not g.is_contiguous() or g.device != p.device:
                raise TypeError("FusedAdamW needs contiguous fp32 gradients on the parameter's device")
            ptrs.append(g.data_ptr())
        if ptrs != self._g_cached:
            # addresses moved (zero_grad(set_to_none=True) lets autograd allocate fresh gradients): refresh the table.
            # Under CUDA-graph capture this copy becomes a graph node that re-uploads the (fixed) capture-pool addresses.
            self._g_host.copy_(torch.tensor(ptrs, dtype=torch.int64))
            self._g.copy_(self._g_host, non_blocking=True)
            self._g_cached = None if torch.cuda.is_current_stream_capturing() else ptrs
        if not torch.cuda.is_current_stream_capturing():
            self._refresh_hyper()
        _call("gvit_mt_adamw_step", _ptr(self._p), _ptr(self._g), _ptr(self._m), _ptr(self._v), _ptr(self._numel), _ptr(self._lr),
              _ptr(self._wd), _ptr(self._chunk_tensor), _ptr(self._chunk_index), _ptr(self.tensor_steps), len(self.params), self.nchunks,
              -1.0 if self.max_norm is None else self.max_norm, self.warmup_steps, self.total_steps, float(self.betas[0]),
              float(self.betas[1]), self.eps, _ptr(self.step_count), _ptr(self.sched), _ptr(self._partial), _stream())

    # -- introspection (host reads: not for the hot loop) -----------------------------------------------------------------
    @property
    def last_grad_norm(self) -> float:
        return float(self.sched[0])

    @property
    def last_lr_factor(self) -> float:
        return float(self.sched[2])

    def get_last_lr(self):
        lam = warmup_cosine_lambda(int(self.step_count), self.warmup_steps, self.total_steps)
        return [float(g["lr"]) * lam for g in self.param_groups]

    # -- torch.optim.AdamW-compatible state ---------------------------------------------------------------------------------
    def state_dict(self):
        steps = self.tensor_steps.detach().to(torch.float32).cpu()
        state, packed, i = {}, [], 0
        for g in self.param_groups:
            ids = list(range(i, i + len(g["params"])))
            for j in ids:
                state[j] = {"step": steps[j].clone(), "exp_avg": self.exp_avg[j].clone(), "exp_avg_sq": self.exp_avg_sq[j].clone()}
            packed.append({**{k: v for k, v in g.items() if k != "params"}, "params": ids})
            i += len(ids)
        return {"state": state, "param_groups": packed,
                "scheduler": {"last_epoch": int(self.step_count), "warmup_steps": self.warmup_steps, "total_steps": self.total_steps}}

    @torch.no_grad()
    def load_state_dict(self, sd):
        if len(sd["param_groups"]) != len(self.param_groups):
            raise ValueError("loaded state dict has a different number of parameter groups")
        step = None
        self.tensor_steps.zero_()
        for j, st in sd.get("state", {}).items():
            j = int(j)
            self.exp_avg[j].copy_(st["exp_avg"])
            self.exp_avg_sq[j].copy_(st["exp_avg_sq"])
            s = int(float(st["step"]))
            self.tensor_steps[j] = s
            step = s if step is None else max(step, s)
        for g, lg in zip(self.param_groups, sd["param_groups"]):
            if len(lg["params"]) != len(g["params"]):
                raise ValueError("loaded state dict contains a parameter group that doesn't match the size of optimizer's group")
            for k in ("lr", "weight_decay"):
                if k in lg:
                    g[k] = lg[k] if "initial_lr" not in lg or k != "lr" else lg["initial_lr"]   # LambdaLR stores the BASE lr there
        sch = sd.get("scheduler")
        if sch is not None:
            step = int(sch["last_epoch"])
        self.step_count.fill_(0 if step is None else step)
        self._refresh_hyper()
