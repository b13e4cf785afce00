"""GPU (>= 2 devices, NCCL): data-parallel equality on hardware - N ranks x one micro-batch each, gradients exchanged by the
bucketed all-reduce of pixelpick_b200.dist.GradAllReducer with the exact global-mean loss scaling, equal ONE rank running
the N micro-batches one after the other and averaging (last bullet; model.py:103-122 per micro-batch).
BatchNorm uses per-micro-batch statistics on both sides (the reference has no SyncBN), Dropout p = 0.  Skipped on a
single-GPU box."""
import os
import socket
from argparse import Namespace

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
C, H, W = 19, 64, 128


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _batch(seed, n_lab):
    rs = np.random.RandomState(seed)
    x = torch.from_numpy(rs.standard_normal((2, 3, H, W)).astype(np.float32))
    y = torch.from_numpy(rs.randint(0, C, size=(2, H, W)).astype(np.int64))
    q = np.zeros((2, H * W), dtype=bool)
    for i in range(2):
        q[i, rs.choice(H * W, n_lab, replace=False)] = True
    return x, y, torch.from_numpy(q.reshape(2, H, W))


def _model(backbone, dev):
    import torch.nn as nn
    from pixelpick_b200.deeplab import DeepLab
    torch.manual_seed(0)
    m = DeepLab(Namespace(use_mc_dropout=False, mc_dropout_p=0.2, n_classes=C), backbone=backbone)
    for mod in m.modules():
        if isinstance(mod, nn.Dropout):
            mod.p = 0.0
    return m.to(dev).train()


def _worker(rank, world, port, backbone, out):
    import torch.distributed as dist
    from pixelpick_b200 import dist as ppdist
    from pixelpick_b200.loss import sparse_cross_entropy
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    model = _model(backbone, dev)
    ppdist.broadcast_parameters(model)
    n_labs = [40 + 25 * r for r in range(world)]  # DIFFERENT labelled-pixel counts per rank: the global mean must weight them
    x, y, q = _batch(100 + rank, n_labs[rank])
    reducer = ppdist.GradAllReducer(model, bucket_mb=4.0)
    reducer.zero_grad()
    loss = sparse_cross_entropy(model.forward_lowres(x.to(dev)), y.to(dev), q.to(dev), C)
    scale = ppdist.global_mean_loss_scale(torch.tensor(float(2 * n_labs[rank]), device=dev))
    (loss * scale).backward()
    reducer()
    torch.cuda.synchronize()
    got = {n: p.grad.detach().float().cpu().clone() for n, p in model.named_parameters()}
    if rank == 0:
        # one rank, the same micro-batches in sequence: d/dw of sum_r n_r * loss_r / sum_r n_r
        ref_model = _model(backbone, dev)
        ref_model.load_state_dict({k: v for k, v in model.state_dict().items()})
        tot = float(sum(2 * n for n in n_labs))
        for r in range(world):
            xr, yr, qr = _batch(100 + r, n_labs[r])
            lr_ = sparse_cross_entropy(ref_model.forward_lowres(xr.to(dev)), yr.to(dev), qr.to(dev), C)
            (lr_ * (2 * n_labs[r] / tot)).backward()
        torch.cuda.synchronize()
        ref = {n: p.grad.detach().float().cpu().clone() for n, p in ref_model.named_parameters()}
        torch.save({"got": got, "ref": ref, "n_buckets": len(reducer.buckets)}, out)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("backbone", ["mobilenet", "resnet"])
def test_n_ranks_equal_one_rank_with_gradient_averaging(tmp_path, backbone):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (run under `gpurun --gpus 2`)")
    import torch.multiprocessing as mp
    world = 2
    out = str(tmp_path / "dp.pt")
    mp.spawn(_worker, args=(world, _free_port(), backbone, out), nprocs=world, join=True)
    r = torch.load(out)
    assert r["n_buckets"] >= 2
    worst = 1.0
    tot = sum(float(v.double().pow(2).sum()) for v in r["ref"].values()) ** 0.5
    for n, g in r["got"].items():
        ref = r["ref"][n]
        rn = ref.norm().item()
        if rn < 1e-5 * tot:
            continue
        cos = torch.dot(g.flatten(), ref.flatten()).item() / (g.norm().item() * rn + 1e-30)
        worst = min(worst, cos)
        # the same kernels on the same micro-batches: the only differences are the order of fp32 atomics (BatchNorm sums,
        # split-K weight gradients) and their amplification through the network
        assert cos > 0.995 and abs(g.norm().item() / rn - 1) < 2e-2, (n, cos, g.norm().item() / rn)
    print(f"{backbone}: worst gradient cosine between {world}-rank DP and the single-rank loop {worst:.6f}")
