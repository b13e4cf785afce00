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


def average_buffers(model):
    """BatchNorm running statistics drift apart between ranks (per-rank batches, no SyncBN as in the reference): average
    the floating-point buffers over the ranks - ONE all-reduce of the flattened statistics - so that validation, the saved
    checkpoint and the query round see the same eval-mode network on every rank.  Integer buffers (num_batches_tracked)
    advance identically on every rank and are left alone."""
    if world() == 1:
        return
    with torch.no_grad():
        bufs = [b for b in model.buffers() if b.is_floating_point()]
        if not bufs:
            return
        flat = torch.cat([b.detach().reshape(-1).float() for b in bufs])
        dist.all_reduce(flat)
        flat.div_(world())
        off = 0
        for b in bufs:
            n = b.numel()
            b.copy_(flat[off:off + n].view_as(b))
            off += n


class GradAllReducer:
    """Data-parallel gradient exchange of the train step: bucketed, overlapped with backward, no copies.

    * every parameter's `.grad` is a VIEW into one flat fp32 buffer, laid out in reverse parameter order (the order in
      which backward produces gradients), so autograd accumulates straight into the buffer the collective reduces and the
      optimiser reads the reduced values in place - no flatten / unflatten passes;
    * the buffer is cut into buckets of ~`bucket_mb`; a post-accumulate hook counts a bucket's parameters and, when the
      last one has its gradient, launches that bucket's all-reduce (`async_op`: NCCL's own stream) while backward keeps
      running on the compute stream.  NVSwitch makes the cost per bucket latency-, not link-bound, so buckets are few
      and large;
    * `reducer()` after backward launches whatever is left (parameters that received no gradient) and makes the compute
      stream wait for the collectives.  The whole sequence is capturable in the step's CUDA graph.
    Average = ReduceOp.AVG on NCCL (no separate division pass); SUM + in-place division on gloo (CPU tests).
    Usage per step:  reducer.zero_grad(); loss.backward(); reducer(); optimizer.step()."""

    def __init__(self, model, bucket_mb: float = 32.0, overlap: bool = True):
        import os
        # measurement knobs (A/B runs of bench.py): bucket size, hooks off = one exchange after backward, no exchange at all
        bucket_mb = float(os.environ.get("PP_DP_BUCKET_MB", bucket_mb))
        overlap = overlap and os.environ.get("PP_DP_OVERLAP", "1") != "0"
        self._nocomm = os.environ.get("PP_DP_NOCOMM", "0") == "1"
        self.params = [p for p in model.parameters() if p.requires_grad]
        order = list(reversed(self.params))
        n = sum(p.numel() for p in order)
        dev = order[0].device
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        self.overlap = overlap and world() > 1
        self._avg = dist.ReduceOp.AVG if (world() > 1 and dist.get_backend() == "nccl") else None
        cap = max(1, int(bucket_mb * (1 << 20) / 4))
        self.buckets = []      # [lo, hi) element ranges of the flat buffer
        self._bucket_of = {}   # parameter -> bucket index
        self._need = []        # parameters per bucket
        off = lo = 0
        cnt = 0
        for p in order:
            if p.dtype != torch.float32:
                raise TypeError("GradAllReducer expects fp32 master parameters")
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            self._bucket_of[p] = len(self.buckets)
            off += p.numel()
            cnt += 1
            if off - lo >= cap:
                self.buckets.append((lo, off))
                self._need.append(cnt)
                lo, cnt = off, 0
        if off > lo:
            self.buckets.append((lo, off))
            self._need.append(cnt)
        self._ready = [0] * len(self.buckets)
        self._works = [None] * len(self.buckets)
        self._handles = []
        if self.overlap:
            for p in order:
                self._handles.append(p.register_post_accumulate_grad_hook(self._hook))

    def zero_grad(self):
        """one memset of the flat buffer (the parameters' .grad views stay in place)"""
        self.flat.zero_()
        for p in self.params:  # an optimiser / user may have dropped the view (set_to_none): restore it
            if p.grad is None or p.grad.data_ptr() != self._view_ptr(p):
                self._rebind()
                break
        self._ready = [0] * len(self.buckets)
        self._works = [None] * len(self.buckets)

    def _view_ptr(self, p):
        if not hasattr(self, "_ptrs"):
            self._ptrs, off = {}, 0
            for q in reversed(self.params):
                self._ptrs[q] = self.flat.data_ptr() + 4 * off
                off += q.numel()
        return self._ptrs[p]

    def _rebind(self):
        off = 0
        for q in reversed(self.params):
            q.grad = self.flat[off:off + q.numel()].view_as(q)
            off += q.numel()

    def _launch(self, b):
        lo, hi = self.buckets[b]
        seg = self.flat[lo:hi]
        if self._nocomm:
            self._works[b] = False
            return
        if self._avg is not None:
            self._works[b] = dist.all_reduce(seg, op=self._avg, async_op=True)
        else:
            self._works[b] = dist.all_reduce(seg, async_op=True)

    def _hook(self, p):
        b = self._bucket_of[p]
        self._ready[b] += 1
        if self._ready[b] == self._need[b] and self._works[b] is None:
            self._launch(b)

    def __call__(self):
        if world() == 1:
            return
        for b in range(len(self.buckets)):
            if self._works[b] is None:
                self._launch(b)
        for b, w in enumerate(self._works):
            if w is False:
                continue
            w.wait()
            if self._avg is None:
                lo, hi = self.buckets[b]
                self.flat[lo:hi].div_(world())


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
