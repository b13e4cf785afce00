"""Multi-GPU plumbing (no reference counterpart: the reference is single-process).
One process per GPU, `torch.distributed` over NCCL.  Both hot paths shard over images:
  train : per-rank micro-batches, ONE all-reduce of the flattened gradient per step (+ one scalar all-reduce of
          the labelled-pixel count so the loss is the exact global mean);
  query : image i -> rank i % world, ONE all-gather of the per-rank picks; rank order reproduces dataset order."""
import torch
import torch.distributed as dist


def world():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def rank():
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def broadcast_parameters(model, src=0):
    if world() == 1:
        return
    with torch.no_grad():
        for t in list(model.parameters()) + list(model.buffers()):
            dist.broadcast(t.data, src)


class GradAllReducer:
    """Flattens every gradient into one fp32 buffer, all-reduces it once and scatters the mean back."""

    def __init__(self, model):
        self.params = [p for p in model.parameters() if p.requires_grad]
        self.flat = None

    def __call__(self):
        if world() == 1:
            return
        grads = [p.grad if p.grad is not None else torch.zeros_like(p) for p in self.params]
        n = sum(g.numel() for g in grads)
        if self.flat is None or self.flat.numel() != n:
            self.flat = torch.empty(n, dtype=torch.float32, device=grads[0].device)
        off = 0
        views = []
        for g in grads:
            v = self.flat[off:off + g.numel()].view_as(g)
            views.append(v)
            off += g.numel()
        torch._foreach_copy_(views, grads)
        dist.all_reduce(self.flat)
        self.flat.div_(world())
        for p, v in zip(self.params, views):
            if p.grad is None:
                p.grad = v.clone()
        torch._foreach_copy_([p.grad for p in self.params], views)


def global_mean_loss_scale(n_local: torch.Tensor) -> torch.Tensor:
    """factor s so that mean over ranks of grad(s * local_mean_loss) == grad of the mean over ALL labelled pixels."""
    if world() == 1:
        return torch.ones((), device=n_local.device)
    tot = n_local.clone().float()
    dist.all_reduce(tot)
    return n_local.float() * world() / tot.clamp_min(1.0)


def all_gather_objects(obj):
    """[obj of rank 0, obj of rank 1, ...] on every rank (picklable Python objects: the per-rank pick dicts / statistics of
    one query round, KBs).  One collective."""
    if world() == 1:
        return [obj]
    out = [None] * world()
    dist.all_gather_object(out, obj)
    return out


def shard_indices(n_items):
    return list(range(rank(), n_items, world()))


def all_gather_rows(t):
    """[n_local, k] -> [n_total, k] in global (round-robin) order; n_local may differ by one between ranks."""
    w = world()
    if w == 1:
        return t
    n_local = torch.tensor([t.shape[0]], device=t.device)
    sizes = [torch.zeros_like(n_local) for _ in range(w)]
    dist.all_gather(sizes, n_local)
    mx = int(max(s.item() for s in sizes))
    pad = torch.zeros((mx,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    pad[:t.shape[0]] = t
    parts = [torch.empty_like(pad) for _ in range(w)]
    dist.all_gather(parts, pad)
    total = int(sum(s.item() for s in sizes))
    out = torch.empty((total,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    for r in range(w):
        out[r::w] = parts[r][: int(sizes[r].item())]
    return out
