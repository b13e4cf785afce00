"""CPU, world_size 2 over gloo: the multi-GPU plumbing of pixelpick_b200/dist.py (gradient all-reduce, exact
global-mean loss scaling, round-robin image sharding + all-gather of the picks)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pixelpick_b200 import dist as ppdist
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ReLU(), torch.nn.Linear(5, 3))
    if rank == 1:
        with torch.no_grad():
            for p in model.parameters():
                p.add_(1.0)
    ppdist.broadcast_parameters(model)
    # per-rank shards with DIFFERENT numbers of labelled samples
    g = torch.Generator().manual_seed(1)
    X, Y = torch.randn((10, 6), generator=g), torch.randint(0, 3, (10,), generator=g)
    cuts = {2: [0, 3, 10], 3: [0, 1, 4, 10], 4: [0, 1, 3, 6, 10]}[world]
    lo, hi = cuts[rank], cuts[rank + 1]
    loss = torch.nn.functional.cross_entropy(model(X[lo:hi]), Y[lo:hi])
    scale = ppdist.global_mean_loss_scale(torch.tensor(float(hi - lo)))
    # tiny buckets: several collectives are launched from the backward hooks, the call after backward waits for them
    reducer = ppdist.GradAllReducer(model, bucket_mb=64 / (1 << 20))
    assert len(reducer.buckets) >= 2
    reducer.zero_grad()
    (loss * scale).backward()
    reducer()
    grads = torch.cat([p.grad.flatten() for p in model.parameters()])
    assert all(p.grad.data_ptr() == reducer._view_ptr(p) for p in model.parameters())  # still views of the flat buffer
    # second step through the same reducer: zero_grad() re-arms the buckets; a dropped view is re-bound
    for p in list(model.parameters())[:1]:
        p.grad = None
    reducer.zero_grad()
    (torch.nn.functional.cross_entropy(model(X[lo:hi]), Y[lo:hi]) * scale).backward()
    reducer()
    assert torch.allclose(torch.cat([p.grad.flatten() for p in model.parameters()]), grads, atol=1e-6)
    # BatchNorm running statistics: per-rank during training, averaged before evaluation (dist.average_buffers)
    bn = torch.nn.BatchNorm1d(4)
    bn.running_mean.fill_(float(rank))
    bn.running_var.fill_(1.0 + rank)
    bn.num_batches_tracked.fill_(7)
    ppdist.average_buffers(bn)
    m_want = sum(range(world)) / world
    assert torch.allclose(bn.running_mean, torch.full((4,), m_want)) and torch.allclose(bn.running_var, torch.full((4,), 1.0 + m_want))
    assert int(bn.num_batches_tracked) == 7
    # sharded query bookkeeping
    idx = ppdist.shard_indices(7)
    rows = torch.tensor([[i, i * 10] for i in idx])
    allrows = ppdist.all_gather_rows(rows)
    # sharded train loader: disjoint per-rank index sets that cover the dataset, reshuffled by set_epoch
    from pixelpick_b200.utils import make_loader
    ds = torch.utils.data.TensorDataset(torch.arange(11))
    dl = make_loader(ds, batch_size=1, n_workers=0, shuffle=True, sharded=True, seed=3)
    epochs = []
    for e in (1, 2):
        dl.sampler.set_epoch(e)
        epochs.append(torch.cat([b[0] for b in dl]).tolist())
    seen = ppdist.all_gather_objects(epochs)
    whole = make_loader(ds, batch_size=4, n_workers=0, shuffle=False, sharded=False)  # query / val loaders stay whole
    assert torch.cat([b[0] for b in whole]).tolist() == list(range(11))
    if rank == 0:
        torch.save({"seen": seen, "grads": grads, "rows": allrows, "params": [p.detach().clone() for p in model.parameters()]}, out)
    dist.barrier()
    dist.destroy_process_group()


import pytest  # noqa: E402


@pytest.mark.parametrize("world", [2, 3, 4])
def test_dp_gradients_equal_single_process(tmp_path, world):
    """world 3 / 4: 7 images do not divide evenly, so the per-rank pick lists differ in length (all_gather_rows pads)."""
    out = str(tmp_path / "r0.pt")
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    got = torch.load(out)
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ReLU(), torch.nn.Linear(5, 3))
    for p, q in zip(model.parameters(), got["params"]):
        assert torch.equal(p, q)  # broadcast from rank 0
    g = torch.Generator().manual_seed(1)
    X, Y = torch.randn((10, 6), generator=g), torch.randint(0, 3, (10,), generator=g)
    torch.nn.functional.cross_entropy(model(X), Y).backward()
    ref = torch.cat([p.grad.flatten() for p in model.parameters()])
    assert torch.allclose(got["grads"], ref, atol=1e-6)
    assert np.array_equal(got["rows"].numpy(), np.array([[i, i * 10] for i in range(7)]))
    per_rank = -(-11 // world)
    for e in range(2):
        shards = [got["seen"][r][e] for r in range(world)]
        assert all(len(s) == per_rank for s in shards)                      # equal work per rank (sampler pads by wrapping)
        assert sorted(set(sum(shards, []))) == list(range(11))               # together they cover the dataset
        assert sum(len(set(s)) for s in shards) <= 11 + world                # overlap only from the padding
    assert got["seen"][0][0] != got["seen"][0][1]                           # set_epoch reshuffles
