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
C, H, W, B = 19, 128, 256, 4  # 4 images / rank: BatchNorm (incl. the pooled branch's, one value per image) sees enough samples


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _batch(seed, n_lab):
    rs = np.random.RandomState(seed)
    x = torch.from_numpy(rs.standard_normal((B, 3, H, W)).astype(np.float32))
    y = torch.from_numpy(rs.randint(0, C, size=(B, H, W)).astype(np.int64))
    q = np.zeros((B, H * W), dtype=bool)
    for i in range(B):
        q[i, rs.choice(H * W, n_lab, replace=False)] = True
    return x, y, torch.from_numpy(q.reshape(B, H, W))


def _model(backbone, dev):
    import torch.nn as nn
    from pixelpick_b200.deeplab import DeepLab
    torch.manual_seed(0)
    m = DeepLab(Namespace(use_mc_dropout=False, mc_dropout_p=0.2, n_classes=C), backbone=backbone)
    for mod in m.modules():
        if isinstance(mod, nn.Dropout):
            mod.p = 0.0
    # well-conditioned weights (see tests/test_train_parity_gpu.py): at the raw initialisation the network is chaotic and two
    # runs of the SAME kernels already disagree in the early layers (fp32 atomics order flips a few bf16 roundings, the
    # BatchNorm-ReLU stack amplifies them); with damped residual branches / mostly-open ReLUs the comparison measures the exchange
    with torch.no_grad():
        for n, p in m.named_parameters():
            if backbone == "resnet" and n.endswith("bn3.weight"):
                p.mul_(0.05)
            if backbone == "mobilenet" and n.endswith(".bias") and p.dim() == 1 and n.startswith(("backbone.", "aspp.", "low_level", "seg_head.segment")):
                p.add_(1.5)
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
    n_labs = [300 + 200 * r for r in range(world)]  # DIFFERENT labelled-pixel counts per rank: the global mean must weight them
    x, y, q = _batch(100 + rank, n_labs[rank])
    reducer = ppdist.GradAllReducer(model, bucket_mb=4.0)
    reducer.zero_grad()
    loss = sparse_cross_entropy(model.forward_lowres(x.to(dev)), y.to(dev), q.to(dev), C)
    scale = ppdist.global_mean_loss_scale(torch.tensor(float(B * n_labs[rank]), device=dev))
    (loss * scale).backward()
    reducer()
    torch.cuda.synchronize()
    got = {n: p.grad.detach().float().cpu().clone() for n, p in model.named_parameters()}
    if rank == 0:
        # one rank, the same micro-batches in sequence: d/dw of sum_r n_r * loss_r / sum_r n_r
        ref_model = _model(backbone, dev)
        ref_model.load_state_dict({k: v for k, v in model.state_dict().items()})
        tot = float(sum(B * n for n in n_labs))
        for r in range(world):
            xr, yr, qr = _batch(100 + r, n_labs[r])
            lr_ = sparse_cross_entropy(ref_model.forward_lowres(xr.to(dev)), yr.to(dev), qr.to(dev), C)
            (lr_ * (B * n_labs[r] / tot)).backward()
        torch.cuda.synchronize()
        ref = {n: p.grad.detach().float().cpu().clone() for n, p in ref_model.named_parameters()}
        # the same single-rank loop once more: how well does the pipeline reproduce ITSELF (fp32 atomics order)
        ref_model.zero_grad(set_to_none=True)
        for r in range(world):
            xr, yr, qr = _batch(100 + r, n_labs[r])
            lr_ = sparse_cross_entropy(ref_model.forward_lowres(xr.to(dev)), yr.to(dev), qr.to(dev), C)
            (lr_ * (B * n_labs[r] / tot)).backward()
        torch.cuda.synchronize()
        ref2 = {n: p.grad.detach().float().cpu().clone() for n, p in ref_model.named_parameters()}
        torch.save({"got": got, "ref": ref, "ref2": ref2, "n_buckets": len(reducer.buckets)}, out)
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
    tot = sum(float(v.double().pow(2).sum()) for v in r["ref"].values()) ** 0.5

    def cosines(a, b):
        rows = []
        for n, g in a.items():
            ref = b[n]
            rn = ref.norm().item()
            if rn < 1e-5 * tot:
                continue
            rows.append((torch.dot(g.flatten(), ref.flatten()).item() / (g.norm().item() * rn + 1e-30), g.norm().item() / rn, n))
        return sorted(rows)

    rows, floor = cosines(r["got"], r["ref"]), cosines(r["ref2"], r["ref"])
    grp = lambda rs, enc: [c for c, _, n in rs if n.startswith("backbone.") == enc]
    print(f"{backbone}: {world}-rank DP vs the single-rank loop: gradient cosine head min {min(grp(rows, False)):.5f} mean "
          f"{np.mean(grp(rows, False)):.5f} | encoder min {min(grp(rows, True)):.5f} mean {np.mean(grp(rows, True)):.5f}")
    print(f"   the single-rank loop vs ITSELF (second run)         : head min {min(grp(floor, False)):.5f} mean "
          f"{np.mean(grp(floor, False)):.5f} | encoder min {min(grp(floor, True)):.5f} mean {np.mean(grp(floor, True)):.5f}")
    for c, ratio, n in rows[:4]:
        print(f"   worst: {n}: cos {c:.5f} norm ratio {ratio:.4f}")
    # One train step does not reproduce itself bit for bit: BatchNorm sums and split-K weight gradients are fp32 atomics, their
    # order flips a few bf16 roundings and the BatchNorm-ReLU stack amplifies that (measured with library convs too,
    # scripts/repro_check.py: cosine 0.95-0.98 between two runs of the same step).  The exchange is correct when the N-rank
    # result is as close to the single-rank loop as that loop is to its own second run.
    for enc in (False, True):
        assert np.mean(grp(rows, enc)) >= np.mean(grp(floor, enc)) - 0.02, enc
        assert min(grp(rows, enc)) >= min(grp(floor, enc)) - 0.08, enc
    assert all(abs(ratio - 1) < 0.25 for _, ratio, _ in rows)
