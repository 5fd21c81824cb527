"""The training step of the reference's ``Trainer.train_epoch`` (/root/reference/src/training/trainer.py:96-120) as
ONE CUDA graph: ``zero_grad`` -> autocast forward -> loss -> backward -> [gradient all-reduce]
-> ``clip_grad_norm_`` -> ``optimizer.step()``, without the per-step host synchronisation of trainer.py:126-132.

A ViT-B + graph step is ~800 kernel launches of 20-300 us each; issued eagerly from Python the host needs about as
long to enqueue them as the B200 needs to run them, so the device idles between kernels.  Capturing the step once and
replaying it removes the host from the loop: one ``cudaGraphLaunch`` per step.

What makes the step capturable:
* every libgvit operator enqueues on the current stream only and allocates through PyTorch's caching allocator
  (which serves a private pool during capture);
* dropout masks come from Philox counters whose per-step offset lives in DEVICE memory
  (``ops.set_rng_offset_tensor``): the captured graph advances it itself, so every replay draws fresh masks;
* the optimizer must be ``capturable`` (e.g. ``torch.optim.AdamW(..., fused=True, capturable=True)``); schedule the
  learning rate by writing into a tensor-valued ``param_group['lr']`` in place;
* ``dp.GradSync`` launches its bucketed NCCL all-reduces from gradient hooks during the captured backward - they
  become graph nodes on the communicator's stream and keep overlapping the remaining backward kernels.
"""
from __future__ import annotations

import torch

from . import ops

__all__ = ["CapturedTrainStep", "CapturedForward"]


class CapturedTrainStep:
    """``loss = step(images, targets)`` - one optimisation step, replayed from a CUDA graph.

    ``criterion(logits, targets)`` returns the loss or a tuple whose first element is the loss (the reference's
    ``DynamicWeightedLoss`` returns ``(total, parts)``, losses.py:26-68).  The returned loss is a static device
    tensor that the next call overwrites; read it (``.item()``) or clone it before stepping again.

    ``images`` / ``targets`` may be any CUDA tensors of the captured shape: they are copied into the graph's static
    input buffers (``self.images`` / ``self.targets`` - a loader may also write into those directly and call
    ``step()`` without arguments).
    """

    def __init__(self, model, criterion, optimizer, *, max_norm: float | None = 1.0, clip_params=None, grad_sync=None,
                 amp_dtype: torch.dtype | None = torch.bfloat16, warmup: int = 3):
        for g in optimizer.param_groups:
            if not g.get("capturable", False):
                raise ValueError("CapturedTrainStep needs a capturable optimizer, e.g. "
                                 "torch.optim.AdamW(..., fused=True, capturable=True)")
        self.model, self.criterion, self.optimizer = model, criterion, optimizer
        self.max_norm, self.grad_sync, self.amp_dtype, self.warmup = max_norm, grad_sync, amp_dtype, max(1, int(warmup))
        self.clip_params = list(clip_params) if clip_params is not None else [p for g in optimizer.param_groups
                                                                              for p in g["params"]]
        self.graph = None
        self.images = self.targets = self.loss = None
        self.rng_offset = None
        self.replays = 0

    # one eager step; also the body that gets captured
    def _body(self):
        self.rng_offset.add_(1 << 40)            # fresh dropout masks on every replay (counters are 64-bit)
        if self.amp_dtype is not None:
            with torch.autocast("cuda", dtype=self.amp_dtype):
                logits = self.model(self.images)
        else:
            logits = self.model(self.images)
        out = self.criterion(logits, self.targets)
        loss = out[0] if isinstance(out, (tuple, list)) else out
        loss.backward()
        if self.grad_sync is not None:
            self.grad_sync.finish()
        if self.max_norm is not None:
            torch.nn.utils.clip_grad_norm_(self.clip_params, self.max_norm, foreach=True)
        self.optimizer.step()
        return loss.detach()

    def capture(self, images: torch.Tensor, targets: torch.Tensor):
        if not images.is_cuda:
            raise RuntimeError("CapturedTrainStep runs on CUDA tensors only")
        dev = images.device
        self.images, self.targets = images.clone(), targets.clone()
        self.rng_offset = torch.zeros(1, dtype=torch.int64, device=dev)
        ops.set_rng_offset_tensor(self.rng_offset)
        try:
            side = torch.cuda.Stream(dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                for _ in range(self.warmup):     # warm-up on a side stream: lazy initialisations, allocator, autotuning
                    self.optimizer.zero_grad(set_to_none=True)
                    self._body()
            torch.cuda.current_stream(dev).wait_stream(side)
            torch.cuda.synchronize(dev)
            self.optimizer.zero_grad(set_to_none=True)          # captured backward then WRITES fresh .grad tensors
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                self.loss = self._body()
            self.graph = graph
        except BaseException:
            ops.set_rng_offset_tensor(None)
            raise
        ops.invalidate_shadows()
        return self

    def __call__(self, images: torch.Tensor | None = None, targets: torch.Tensor | None = None):
        if self.graph is None:
            if images is None or targets is None:
                raise RuntimeError("the first call must pass example images and targets (they fix the captured shapes)")
            self.capture(images, targets)
        if images is not None and images.data_ptr() != self.images.data_ptr():
            self.images.copy_(images, non_blocking=True)
        if targets is not None and targets.data_ptr() != self.targets.data_ptr():
            self.targets.copy_(targets, non_blocking=True)
        self.graph.replay()
        self.replays += 1
        ops.invalidate_shadows()                 # the masters moved on the device: eager users must re-cast
        return self.loss

    def release(self):
        """Drop the graph and its memory pool; dropout goes back to host-drawn seeds.  With a ``grad_sync`` the graph
        holds NCCL kernels: release it (and synchronize) BEFORE ``torch.distributed.destroy_process_group()`` - tearing
        the communicator down under a live graph hung both ranks in testing."""
        self.graph = None
        self.loss = None
        ops.set_rng_offset_tensor(None)


class CapturedForward:
    """``probs = fwd(images)`` - the inference call of /root/reference/scripts/evaluate.py:104-115 (``model.eval()``,
    ``torch.no_grad()``, forward, sigmoid) replayed from ONE CUDA graph per batch shape.

    A ViT-B + graph forward is ~330 kernel launches; at batch 32 they take ~2.5 ms on the device but ~6.4 ms to issue from
    Python, so small-batch inference is launch-bound until the host is out of the loop.  Graphs are cached per input shape;
    the returned tensor is a static buffer that the next call with the same shape overwrites.  Inference is collective-free,
    so this is per-GPU state only.
    """

    def __init__(self, model, *, amp_dtype: torch.dtype | None = torch.bfloat16, activation=torch.sigmoid, warmup: int = 2):
        self.model, self.amp_dtype, self.activation, self.warmup = model, amp_dtype, activation, max(1, int(warmup))
        self._graphs: dict = {}

    def _body(self, x):
        if self.amp_dtype is not None:
            with torch.autocast("cuda", dtype=self.amp_dtype):
                y = self.model(x)
        else:
            y = self.model(x)
        return self.activation(y.float()) if self.activation is not None else y

    @torch.no_grad()
    def __call__(self, images: torch.Tensor) -> torch.Tensor:
        if not images.is_cuda:
            raise RuntimeError("CapturedForward runs on CUDA tensors only")
        if self.model.training:
            raise RuntimeError("CapturedForward is the inference path: call model.eval() first (dropout would be frozen into the graph)")
        key = (tuple(images.shape), images.dtype, images.device)
        entry = self._graphs.get(key)
        if entry is None:
            static_in = images.clone()
            side = torch.cuda.Stream(images.device)
            side.wait_stream(torch.cuda.current_stream(images.device))
            with torch.cuda.stream(side):
                for _ in range(self.warmup):
                    self._body(static_in)
            torch.cuda.current_stream(images.device).wait_stream(side)
            torch.cuda.synchronize(images.device)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                static_out = self._body(static_in)
            entry = self._graphs[key] = (graph, static_in, static_out)
        graph, static_in, static_out = entry
        if images.data_ptr() != static_in.data_ptr():
            static_in.copy_(images, non_blocking=True)
        graph.replay()
        return static_out

    def release(self):
        self._graphs.clear()
