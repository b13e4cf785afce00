"""CUDA-graph capture of the whole train step (forward_lowres -> fused sparse CE -> backward -> [grad all-reduce]
-> optimiser), for the launch-bound regime of the reference batch size (B = 4: ~1000 kernels, CPU-bound in eager mode).

Static inputs: the image batch, a fixed-capacity labelled-pixel list + its device-side count (pp_sparse_ce reads the
count on the device), the dropout step counter (device int64 advanced inside the graph) and tensor learning rates.
"""
import os

import torch

from . import dist as ppdist
from .loss import LabelCapacityError, labelled_pixel_list_host, sparse_cross_entropy


def make_capturable_adam(param_groups):
    """torch.optim.Adam(fused, capturable) with TENSOR learning rates so a scheduler can change them between replays; its
    step is one launch of pp_adam_step_multi (optim.FusedAdam; PP_ADAM=torch keeps torch's multi-tensor kernel: A/B only)."""
    dev = param_groups[0]["params"][0].device if not hasattr(param_groups[0]["params"], "__next__") else None
    groups = []
    for g in param_groups:
        g = dict(g)
        g["params"] = list(g["params"])
        dev = g["params"][0].device
        g["lr"] = torch.tensor(float(g["lr"]), dtype=torch.float32, device=dev)
        groups.append(g)
    if os.environ.get("PP_ADAM", "ours") == "torch":
        return torch.optim.Adam(groups, fused=True, capturable=True)
    from .optim import FusedAdam
    return FusedAdam(groups)


class GraphedTrainStep:
    def __init__(self, model, optimizer, batch_shape, ignore_index, capacity, device, reducer=None, warmup=3,
                 n_classes=None):
        B, H, W = batch_shape
        self.model, self.opt, self.ignore_index, self.capacity, self.size = model, optimizer, ignore_index, capacity, (H, W)
        self.reducer = reducer
        self.x = torch.zeros((B, 3, H, W), dtype=torch.float32, device=device)
        # the three labelled-pixel lists and their count live in ONE int32 buffer: one H2D copy instead of four (every
        # copy on the B200 pod pays a fixed ~0.2-0.9 ms before its first byte, scripts/bench_h2d.py)
        self.meta = torch.zeros(3 * capacity + 1, dtype=torch.int32, device=device)
        self.meta[3 * capacity] = 1
        self.px = [self.meta[i * capacity:(i + 1) * capacity] for i in range(3)]
        self.n_valid = self.meta[3 * capacity:3 * capacity + 1]
        self.loss_scale = torch.ones((), dtype=torch.float32, device=device)
        # double buffering: the NEXT batch is uploaded into staging buffers on a copy stream while the graph of the
        # current batch runs; commit() moves it into the graph's static inputs with two device-to-device copies
        self.copy_stream = torch.cuda.Stream(device=device)
        self.x_stage = torch.empty_like(self.x)
        self.meta_stage = torch.empty_like(self.meta)
        self.x_host = torch.empty(self.x.shape, dtype=torch.float32).pin_memory()
        self.meta_host = torch.zeros(3 * capacity + 1, dtype=torch.int32).pin_memory()
        self.stage_ready = torch.cuda.Event()
        self.commit_done = torch.cuda.Event()
        self.commit_done.record()
        self._staged = False
        # running metrics accumulated INSIDE the captured step (confusion matrix of the labelled pixels, loss sum): the
        # loop needs no per-step device->host read at all
        from . import _lib
        self.metrics = _lib.DeviceMetrics(n_classes, device) if n_classes else None
        self.graph = None
        self._warm = warmup

    def _zero_grad(self):
        if self.reducer is not None:
            self.reducer.zero_grad()  # gradients live in the reducer's flat buffer: one memset, views stay bound
        else:
            self.opt.zero_grad(set_to_none=True)

    def _step(self):
        if self.reducer is not None:
            self.reducer.zero_grad()  # inside the captured step: every replay starts from a zeroed flat buffer
        lowres = self.model.forward_lowres(self.x)
        loss, pred, _ = sparse_cross_entropy(lowres, None, None, self.ignore_index, size=self.size, return_pred=True,
                                             px=self.px, n_valid=self.n_valid)
        (loss * self.loss_scale).backward()
        if self.reducer is not None:
            self.reducer()
        self.opt.step()
        if self.metrics is not None:
            self.metrics.accumulate(self.px[2], pred, loss=loss.detach().reshape(1), n_valid=self.n_valid)
        return loss, pred

    def prefetch(self, x, y, queries):
        """host batch (CPU tensors from the dataloader) -> device STAGING buffers, asynchronously on the copy stream (it
        overlaps whatever the main stream is running); returns the host label list.  Follow with commit().
        Returns None (nothing staged) when the batch holds more labelled pixels than the captured capacity: the caller
        runs that batch through the eager step."""
        if x.is_cuda:
            return self._prefetch_device(x, y, queries)
        try:
            pi, px, pl, n = labelled_pixel_list_host(y, queries, self.ignore_index, self.capacity,
                                                     n_classes=self.metrics.n_classes if self.metrics is not None else None)
        except LabelCapacityError:
            self._staged = False
            return None
        self.commit_done.synchronize()  # the previous commit has consumed the staging / pinned buffers
        cap = self.capacity
        mh = self.meta_host
        mh[:cap], mh[cap:2 * cap], mh[2 * cap:3 * cap], mh[3 * cap] = pi, px, pl, int(n)
        src = x if x.is_pinned() else self.x_host.copy_(x)  # pageable batches go through the pinned bounce buffer
        with torch.cuda.stream(self.copy_stream):
            self.x_stage.copy_(src, non_blocking=True)
            self.meta_stage.copy_(mh, non_blocking=True)
            self.stage_ready.record(self.copy_stream)
        self._staged = True
        return pl[: int(n)].clone()

    def _prefetch_device(self, x, y, queries):
        """The batch is already on the device (augmented there, Model._device_augment): the labelled-pixel list is built with
        device ops (one host sync for its length) and the staging buffers are filled on the copy stream."""
        from .loss import labelled_pixel_list
        self.commit_done.synchronize()
        cap = self.capacity
        cur = torch.cuda.current_stream()
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_stream(cur)  # no-op when the caller already works on the copy stream
            pi, px, pl = labelled_pixel_list(y, queries, self.ignore_index)
            n = int(pi.numel())
            if n > cap:
                self._staged = False
                return None
            n_classes = self.metrics.n_classes if self.metrics is not None else None
            if n and n_classes is not None and (int(pl.min()) < 0 or int(pl.max()) >= n_classes):
                raise IndexError(f"Target {int(pl[(pl < 0) | (pl >= n_classes)][0])} is out of bounds.")
            ms = self.meta_stage
            ms.zero_()
            ms[:n], ms[cap:cap + n], ms[2 * cap:2 * cap + n] = pi, px, pl
            ms[3 * cap] = n
            self.x_stage.copy_(x, non_blocking=True)
            self.stage_ready.record(self.copy_stream)
        for t in (x, y, queries):
            if t is not None:
                t.record_stream(self.copy_stream)
        self._staged = True
        return pl

    def drop_staged(self):
        """forget a prefetched batch (the loop decided not to run it through the graph)."""
        self._staged = False

    def commit(self):
        """staging -> the graph's static inputs (device-to-device, on the current stream)."""
        assert self._staged, "commit() without a prefetch()"
        cur = torch.cuda.current_stream()
        cur.wait_event(self.stage_ready)
        self.x.copy_(self.x_stage, non_blocking=True)
        self.meta.copy_(self.meta_stage, non_blocking=True)
        self.commit_done.record(cur)
        self._staged = False
        if ppdist.world() > 1:
            self.loss_scale.copy_(ppdist.global_mean_loss_scale(self.n_valid.float().reshape(())))

    def load(self, x, y, queries):
        """prefetch + commit: host batch -> static device buffers; returns the host label list."""
        labels = self.prefetch(x, y, queries)
        self.commit()
        return labels

    def _snapshot(self):
        """clones of everything the warm-up steps mutate: parameters, buffers (BatchNorm running statistics, counters),
        the optimiser state and the dropout step counter — restored after capture so that capturing is free of side
        effects on the training trajectory (the first replay is the first optimisation step)."""
        snap = {"model": [t.detach().clone() for t in list(self.model.parameters()) + list(self.model.buffers())],
                "rng": None if self.model._rng_step is None else self.model._rng_step.clone(), "opt": []}
        for st in self.opt.state.values():
            snap["opt"].append({k: (v.clone() if torch.is_tensor(v) else v) for k, v in st.items()})
        snap["opt_empty"] = len(self.opt.state) == 0
        return snap

    def _restore(self, snap):
        with torch.no_grad():
            for t, c in zip(list(self.model.parameters()) + list(self.model.buffers()), snap["model"]):
                t.copy_(c)
            if snap["rng"] is not None and self.model._rng_step is not None:
                self.model._rng_step.copy_(snap["rng"])
            elif self.model._rng_step is not None:
                self.model._rng_step.zero_()
            if snap["opt_empty"]:  # state was created by the warm-up: reset it in place (addresses are captured)
                for st in self.opt.state.values():
                    for v in st.values():
                        if torch.is_tensor(v):
                            v.zero_()
            else:
                for st, c in zip(self.opt.state.values(), snap["opt"]):
                    for k, v in st.items():
                        if torch.is_tensor(v):
                            v.copy_(c[k])

    def capture(self, restore_state=False):
        self.model.train()
        snap = self._snapshot() if restore_state else None
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(self._warm):
                self._zero_grad()
                self._step()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        self._zero_grad()
        with torch.cuda.graph(self.graph):
            self.loss, self.pred = self._step()
        if snap is not None:
            self._restore(snap)
        if self.metrics is not None:
            self.metrics.reset()  # the warm-up steps are not part of the epoch
        return self

    def __call__(self):
        """replay on the current contents of the static buffers; returns (loss [0-d], pred_at [capacity]) device tensors."""
        if self.graph is None:
            self.capture()
        self.graph.replay()
        return self.loss, self.pred
