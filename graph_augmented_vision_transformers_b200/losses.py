"""Host-side training loss with the interface of the reference's ``DynamicWeightedLoss``
(/root/reference/src/training/losses.py:7-68): three multi-label terms - weighted BCE, focal (gamma 2) and
asymmetric (gamma+ 1, gamma- 4, clamp 1e-8) - mixed by a softmax over three learnable scalars.

This is (B, 14) element-wise work on the logits: outside the hot path, kept as plain
PyTorch so a training step on the device is complete.  State-dict keys match the reference
(``lambda_wbce``, ``lambda_focal``, ``lambda_asl``, buffer ``pos_weight``).
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F


class DynamicWeightedLoss(nn.Module):
    def __init__(self, num_classes, class_weights=None):
        super().__init__()
        self.num_classes, self.class_weights, self.gamma = num_classes, class_weights, 2.0
        for name in ("lambda_wbce", "lambda_focal", "lambda_asl"):
            setattr(self, name, nn.Parameter(torch.tensor(1.0)))
        self.register_buffer("pos_weight", torch.ones(num_classes) if class_weights is None else class_weights)

    def forward(self, logits, targets):
        logits = logits.float()
        mix = torch.softmax(torch.stack([self.lambda_wbce, self.lambda_focal, self.lambda_asl]), dim=0)
        elementwise = F.binary_cross_entropy_with_logits(logits, targets, reduction="none")
        wbce = F.binary_cross_entropy_with_logits(logits, targets, pos_weight=self.pos_weight)
        focal = ((1.0 - torch.exp(-elementwise)) ** self.gamma * elementwise).mean()
        prob = torch.sigmoid(logits)
        log_pos = torch.log(prob.clamp(min=1e-8))
        log_neg = torch.log((1.0 - prob).clamp(min=1e-8))
        asl = -(targets * log_pos * (1.0 - prob) + (1.0 - targets) * log_neg * prob.pow(4)).mean()
        total = mix[0] * wbce + mix[1] * focal + mix[2] * asl
        # exactly the reference's key set (losses.py:62-66): Trainer.train_epoch appends `.item()` of every entry to lists
        # pre-seeded with these three keys (trainer.py:129-130); the mixing weights are exposed by get_loss_weights()
        return total, {"wbce": wbce.detach(), "focal": focal.detach(), "asl": asl.detach()}

    def get_loss_weights(self):
        """The current mixing weights as a numpy array (losses.py:70-76)."""
        with torch.no_grad():
            return torch.softmax(torch.stack([self.lambda_wbce, self.lambda_focal, self.lambda_asl]), 0).cpu().numpy()
