"""Checkpoints in the reference's format, and a ``--resume`` that restores everything.

``Trainer.save_checkpoint`` (/root/reference/src/training/trainer.py:188-214) writes ``epoch``, ``model_state_dict``,
``optimizer_state_dict``, ``scheduler_state_dict``, ``scaler_state_dict``, ``best_val_auc``, ``metrics`` and ``config``;
``scripts/train.py:160-168`` then resumes from the model weights and the epoch ONLY - optimiser moments, the position in
the warm-up / cosine schedule, the loss-scaler state and the best metric are silently reset.  ``save_checkpoint`` writes the
same keys (plus the loss module's mixing weights), ``load_checkpoint`` restores all of them, for ``torch.optim.AdamW`` +
``LambdaLR`` as well as for :class:`~.optim.FusedAdamW` (whose scheduler state is its step counter).
"""
from __future__ import annotations

import os

import torch

__all__ = ["save_checkpoint", "load_checkpoint"]


def save_checkpoint(path, *, model, optimizer, epoch, scheduler=None, scaler=None, criterion=None, best_val_auc=0.0,
                    metrics=None, config=None):
    opt_sd = optimizer.state_dict()
    if scheduler is not None:
        sched_sd = scheduler.state_dict()
    else:                                        # FusedAdamW: the schedule position is the optimiser's step counter
        sched_sd = dict(opt_sd.get("scheduler", {}))
    ckpt = {
        "epoch": int(epoch),
        "model_state_dict": model.state_dict(),
        "optimizer_state_dict": opt_sd,
        "scheduler_state_dict": sched_sd,
        "scaler_state_dict": scaler.state_dict() if scaler is not None else {},
        "best_val_auc": float(best_val_auc),
        "metrics": metrics if metrics is not None else {},
        "config": config if config is not None else {},
    }
    if criterion is not None:
        ckpt["criterion_state_dict"] = criterion.state_dict()
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    tmp = f"{path}.tmp"
    torch.save(ckpt, tmp)
    os.replace(tmp, path)                        # a crash mid-write never leaves a truncated checkpoint behind
    return path


def load_checkpoint(path, *, model, optimizer=None, scheduler=None, scaler=None, criterion=None, strict=True,
                    map_location=None):
    """Restore a checkpoint written by ``save_checkpoint`` or by the reference's ``Trainer.save_checkpoint``.

    Returns ``{'start_epoch', 'best_val_auc', 'metrics', 'config'}``; ``start_epoch = epoch + 1`` as
    ``scripts/train.py:166`` computes it.  Objects passed as ``None`` are left alone.
    """
    if not os.path.isfile(path):
        raise FileNotFoundError(f"no checkpoint found at {path}")     # the reference logs an error and trains from scratch
    ckpt = torch.load(path, map_location=map_location, weights_only=False)
    model.load_state_dict(ckpt["model_state_dict"], strict=strict)
    if criterion is not None and "criterion_state_dict" in ckpt:
        criterion.load_state_dict(ckpt["criterion_state_dict"])
    if optimizer is not None and "optimizer_state_dict" in ckpt:
        sd = dict(ckpt["optimizer_state_dict"])
        if scheduler is None and "scheduler" not in sd and ckpt.get("scheduler_state_dict"):
            # a reference checkpoint loaded into FusedAdamW: LambdaLR's last_epoch is the number of completed steps
            sd["scheduler"] = {"last_epoch": int(ckpt["scheduler_state_dict"].get("last_epoch", 0))}
        optimizer.load_state_dict(sd)
    if scheduler is not None and ckpt.get("scheduler_state_dict"):
        scheduler.load_state_dict(ckpt["scheduler_state_dict"])
    if scaler is not None and ckpt.get("scaler_state_dict"):
        scaler.load_state_dict(ckpt["scaler_state_dict"])
    from . import ops
    ops.invalidate_shadows()                     # the masters changed under the 16-bit parameter shadows
    return {"start_epoch": int(ckpt.get("epoch", -1)) + 1, "best_val_auc": float(ckpt.get("best_val_auc", 0.0)),
            "metrics": ckpt.get("metrics", {}), "config": ckpt.get("config", {})}
