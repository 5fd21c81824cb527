"""f2 / f3 - the optimiser side of the reference's training step (trainer.py:47-56,77-87,114-118) as libgvit multi-tensor kernels
(optim.FusedAdamW -> gvit_mt_adamw_step) against torch.optim.AdamW + LambdaLR + clip_grad_norm_, inside the captured
training step, and through a checkpoint / resume cycle in the reference's checkpoint format (trainer.py:188-214)."""
import math

import pytest
import torch

from gpu_util import DEV
from graph_augmented_vision_transformers_b200 import checkpoint, modules, optim
from graph_augmented_vision_transformers_b200.losses import DynamicWeightedLoss
from graph_augmented_vision_transformers_b200.step import CapturedTrainStep

pytestmark = pytest.mark.gpu

SHAPES = [(768, 768), (3072,), (1, 1, 768), (37,), (5, 13), (100000,), (3,)]


def _params(seed):
    g = torch.Generator().manual_seed(seed)
    return [torch.nn.Parameter(torch.randn(*s, generator=g).to(DEV)) for s in SHAPES]


def test_fused_adamw_matches_torch_adamw_lambdalr_clip():
    warm, total = 3, 10
    pa, pb = _params(0), _params(0)
    ga = [{"params": pa[:-1]}, {"params": pa[-1:], "lr": 1e-4}]          # trainer.py:47-56: second group at 0.1 x lr
    gb = [{"params": pb[:-1]}, {"params": pb[-1:], "lr": 1e-4}]
    ref = torch.optim.AdamW(ga, lr=1e-3, weight_decay=0.05, betas=(0.9, 0.999), eps=1e-8)
    sch = torch.optim.lr_scheduler.LambdaLR(ref, lambda s: optim.warmup_cosine_lambda(s, warm, total))
    ours = optim.FusedAdamW(gb, lr=1e-3, weight_decay=0.05, betas=(0.9, 0.999), eps=1e-8, max_norm=1.0, warmup_steps=warm,
                            total_steps=total)
    g = torch.Generator().manual_seed(1)
    for it in range(8):
        scale = 10.0 if it % 2 == 0 else 1e-3                              # clipped and un-clipped steps
        for a, b in zip(pa, pb):
            gr = (torch.randn(a.shape, generator=g) * scale).to(DEV)
            a.grad, b.grad = gr.clone(), gr.clone()
        if it == 5:
            pa[1].grad = pb[1].grad = None                                 # a parameter without gradient is skipped
        want_norm = torch.nn.utils.clip_grad_norm_(pa, 1.0)
        ref.step()
        sch.step()
        ours.step()
        assert abs(ours.last_grad_norm - float(want_norm)) < 1e-5 * float(want_norm)
        assert abs(ours.last_lr_factor - optim.warmup_cosine_lambda(it, warm, total)) < 1e-6
        for a, b in zip(pa, pb):
            assert float((a - b).abs().max()) <= 2e-6 * float(a.abs().max()) + 1e-9, it
    assert math.isclose(ours.get_last_lr()[0], sch.get_last_lr()[0], rel_tol=1e-6)
    # torch-compatible state: load ours into a fresh torch AdamW and vice versa
    sd = ours.state_dict()
    fresh = optim.FusedAdamW([{"params": _params(3)[:-1]}, {"params": _params(3)[-1:], "lr": 1e-4}], lr=1e-3, weight_decay=0.05,
                             max_norm=1.0, warmup_steps=warm, total_steps=total)
    fresh.load_state_dict(ref.state_dict())                                # torch -> ours (step from state['step'])
    assert int(fresh.step_count) == 8 and int(fresh.tensor_steps[1]) == 7 and int(fresh.tensor_steps[0]) == 8   # skipped once
    assert torch.allclose(fresh.exp_avg[0], ours.exp_avg[0], rtol=1e-5, atol=1e-8)
    t2 = torch.optim.AdamW([{"params": _params(4)[:-1]}, {"params": _params(4)[-1:], "lr": 1e-4}], lr=1e-3, weight_decay=0.05)
    t2.load_state_dict({k: v for k, v in sd.items() if k != "scheduler"})  # ours -> torch
    assert torch.equal(t2.state[t2.param_groups[0]["params"][0]]["exp_avg"], ours.exp_avg[0])


CFG = dict(img_size=64, patch_size=8, embed_dim=128, depth=2, num_heads=2, mlp_ratio=2.0, graph_mode="knn", graph_k=4)


def _trainer(seed=0, warm=2, total=12):
    torch.manual_seed(seed)
    model = modules.VisionTransformer(**CFG).to(DEV).train()
    crit = DynamicWeightedLoss(14).to(DEV)
    opt = optim.FusedAdamW([{"params": model.parameters()}, {"params": crit.parameters(), "lr": 1e-5}], lr=1e-4, weight_decay=0.05,
                           max_norm=1.0, warmup_steps=warm, total_steps=total)
    return model, crit, opt


def _eager_step(model, crit, opt, img, tgt):
    opt.zero_grad(set_to_none=True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        logits = model(img)
    loss, _ = crit(logits, tgt)
    loss.backward()
    opt.step()
    return loss.detach()


def test_captured_step_with_fused_adamw_equals_eager_steps():
    g = torch.Generator().manual_seed(5)
    img = torch.randn(4, 3, 64, 64, generator=g).to(DEV)
    tgt = (torch.rand(4, 14, generator=g) > 0.7).float().to(DEV)
    m1, c1, o1 = _trainer()
    m2, c2, o2 = _trainer()
    cap = CapturedTrainStep(m2, c2, o2, max_norm=None, warmup=1).capture(img, tgt)   # the clip lives inside the optimiser
    eager = [float(_eager_step(m1, c1, o1, img, tgt)) for _ in range(1 + 3)]          # warm-up body + 3 replays
    replay = [float(cap(img, tgt)) for _ in range(3)]
    assert int(o2.step_count) == 1 + 3 == int(o1.step_count)                          # capturing records the body, it does not run it
    for (n, a), b in zip(m1.named_parameters(), m2.parameters()):
        assert torch.allclose(a, b, rtol=1e-4, atol=1e-6), n
    assert all(math.isfinite(x) for x in eager + replay)
    cap.release()


def test_checkpoint_resume_restores_optimizer_schedule_and_weights(tmp_path):
    g = torch.Generator().manual_seed(6)
    img = torch.randn(4, 3, 64, 64, generator=g).to(DEV)
    tgt = (torch.rand(4, 14, generator=g) > 0.7).float().to(DEV)
    m1, c1, o1 = _trainer(seed=1)
    for _ in range(5):
        _eager_step(m1, c1, o1, img, tgt)                                  # the uninterrupted run: 5 steps
    m2, c2, o2 = _trainer(seed=1)
    for _ in range(3):
        _eager_step(m2, c2, o2, img, tgt)
    path = checkpoint.save_checkpoint(str(tmp_path / "ck.pt"), model=m2, optimizer=o2, criterion=c2, epoch=7, best_val_auc=0.8125,
                                      metrics={"mean_auc": 0.8125}, config={"model": CFG})
    ck = torch.load(path, weights_only=False)
    assert set(ck) >= {"epoch", "model_state_dict", "optimizer_state_dict", "scheduler_state_dict", "scaler_state_dict",
                       "best_val_auc", "metrics", "config"}               # trainer.py:189-198
    m3, c3, o3 = _trainer(seed=99)                                         # different init: everything must come from the file
    info = checkpoint.load_checkpoint(path, model=m3, optimizer=o3, criterion=c3)
    assert info["start_epoch"] == 8 and info["best_val_auc"] == 0.8125 and int(o3.step_count) == 3
    for _ in range(2):
        _eager_step(m3, c3, o3, img, tgt)
    for (n, a), b in zip(m1.named_parameters(), m3.parameters()):
        assert torch.equal(a, b), n                                        # bit-identical to the run that never stopped
    assert torch.equal(c1.lambdas if hasattr(c1, "lambdas") else next(c1.parameters()), next(c3.parameters()))
    with pytest.raises(FileNotFoundError):
        checkpoint.load_checkpoint(str(tmp_path / "missing.pt"), model=m3)
