"""N>1 path on CPU (gloo, world_size 2): bucketed, hook-launched gradient averaging reproduces the single-process
gradients over the concatenated batch, and replicas start from identical weights."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from graph_augmented_vision_transformers_b200 import dp
from oracle import vit_oracle

CFG = dict(img_size=32, patch_size=8, embed_dim=64, depth=2, num_heads=4, mlp_ratio=2.0, graph_mode="knn", graph_k=4)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _loss(model, img, tgt):
    return vit_oracle.multilabel_loss(model(img), tgt, torch.ones(3), torch.ones(14))


def _worker(rank, world, port, out_dir):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    torch.set_num_threads(1)
    r, w, _ = dp.init_from_env("gloo")
    assert (r, w) == (rank, world)
    torch.manual_seed(100 + rank)                       # deliberately different initial replicas
    model = vit_oracle.VisionTransformer(**CFG)
    dp.broadcast_parameters(model, src=0)
    sync = dp.GradSync(model, bucket_mb=0.05)
    g = torch.Generator().manual_seed(7)
    img, tgt = torch.randn(8, 3, 32, 32, generator=g), (torch.rand(8, 14, generator=g) > 0.7).float()
    lo, hi = dp.shard_batch(8, rank, world)
    for _ in range(2):                                  # two steps: the hooks must re-arm
        model.zero_grad(set_to_none=True)
        _loss(model, img[lo:hi], tgt[lo:hi]).backward()
        sync.finish()
    assert sync.collectives_issued == 2 * len(sync.buckets) and len(sync.buckets) > 2
    torch.save({n: p.grad.clone() for n, p in model.named_parameters()}, os.path.join(out_dir, f"g{rank}.pt"))
    if rank == 0:
        torch.save(model.state_dict(), os.path.join(out_dir, "w.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gradient_average_equals_single_process(tmp_path):
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    g0, g1 = torch.load(tmp_path / "g0.pt"), torch.load(tmp_path / "g1.pt")
    model = vit_oracle.VisionTransformer(**CFG)
    model.load_state_dict(torch.load(tmp_path / "w.pt"))
    g = torch.Generator().manual_seed(7)
    img, tgt = torch.randn(8, 3, 32, 32, generator=g), (torch.rand(8, 14, generator=g) > 0.7).float()
    _loss(model, img, tgt).backward()
    for n, p in model.named_parameters():
        assert torch.equal(g0[n], g1[n]), n             # both ranks hold the same averaged gradient
        err = (g0[n] - p.grad).abs().max() / p.grad.abs().max().clamp_min(1e-12)
        assert err < 1e-5, (n, float(err))
