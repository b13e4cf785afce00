"""Mirror of the reference `model.py:Model` (active-learning round loop, model.py:14-239) on the B200 hot paths.

Same constructor/`__call__` contract, files written (log_train.txt, log_val.txt, best_miou_model.pt,
{n}_query/queries.pkl, query_stats.pkl) and round structure: every round re-initialises the model (model.py:163),
trains n_epochs, then queries with the LAST-epoch model (model.py:80-83).  The train step is
  forward_lowres (tcgen05 head) -> fused upsample + sparse CE kernel -> custom backward -> optimiser
and the running metrics are computed from the labelled pixels only (identical confusion matrix, see
utils.RunningScore.update_pairs) instead of copying two full int64 maps to the host every step (model.py:125).
Not mirrored: the PNG visualiser (model.py:150-158, host-side matplotlib-style drawing, out of scope)."""
import os
from copy import deepcopy
from math import ceil

import numpy as np
import torch
import torch.nn.functional as F

from . import dist as ppdist
from .loss import sparse_cross_entropy
from .query import QuerySelector
from .utils import AverageMeter, RunningScore, get_dataloader, get_lr_scheduler, get_model, get_optimizer, optimizer_kind, write_log


class Model:
    def __init__(self, args, dataloaders=None):
        self.args = args
        self.best_miou = -1.0
        self.dataset_name = args.dataset_name
        self.debug = args.debug
        if not torch.cuda.is_available():
            raise RuntimeError("pixelpick_b200.Model needs a CUDA device (no CPU fallback)")
        self.device = torch.device("cuda", torch.cuda.current_device())
        self.dir_checkpoints = f"{args.dir_root}/checkpoints/{args.experim_name}"
        self.experim_name = args.experim_name
        self.ignore_index = args.ignore_index
        self.init_n_pixels = args.n_init_pixels
        self.max_budget = args.max_budget
        self.n_classes = args.n_classes
        self.n_epochs = args.n_epochs
        self.n_pixels_by_us = args.n_pixels_by_us
        self.network_name = args.network_name
        self.nth_query = -1
        self.stride_total = args.stride_total
        if dataloaders is None:
            self.dataloader = get_dataloader(deepcopy(args), val=False, query=False, shuffle=True,
                                             batch_size=args.batch_size, n_workers=args.n_workers)
            self.dataloader_query = get_dataloader(deepcopy(args), val=False, query=True, shuffle=False, batch_size=1,
                                                   n_workers=args.n_workers)
            self.dataloader_val = get_dataloader(deepcopy(args), val=True, query=False, shuffle=False, batch_size=1,
                                                 n_workers=args.n_workers)
        else:
            self.dataloader, self.dataloader_query, self.dataloader_val = dataloaders
        self.lr_scheduler_type = args.lr_scheduler_type
        self.query_selector = QuerySelector(args, self.dataloader_query, device=self.device)
        self._use_graph, self._graph, self._graph_shape, self._graph_labels = False, None, None, None
        self._gpu_aug = None  # --gpu_augment: augment.GpuAugment, built for the first raw batch
        self.running_loss, self.running_score = AverageMeter(), RunningScore(args.n_classes)

    def __call__(self):
        if self.n_pixels_by_us == 0:  # fully-supervised model (model.py:55-64)
            d = f"{self.dir_checkpoints}/fully_sup"
            os.makedirs(d, exist_ok=True)
            self.log_train, self.log_val = f"{d}/log_train.txt", f"{d}/log_val.txt"
            write_log(self.log_train, header=["epoch", "mIoU", "pixel_acc", "loss"])
            write_log(self.log_val, header=["epoch", "mIoU", "pixel_acc"])
            self._train()
            return
        n_stages = self.max_budget // self.n_pixels_by_us
        n_stages += 1 if self.init_n_pixels > 0 else 0
        print("n_stages:", n_stages)
        for nth_query in range(n_stages):
            d = f"{self.dir_checkpoints}/{nth_query}_query"
            os.makedirs(d, exist_ok=True)
            self.log_train, self.log_val = f"{d}/log_train.txt", f"{d}/log_val.txt"
            write_log(self.log_train, header=["epoch", "mIoU", "pixel_acc", "loss"])
            write_log(self.log_val, header=["epoch", "mIoU", "pixel_acc"])
            self.nth_query = nth_query
            model = self._train()
            queries = self.query_selector(nth_query, model)
            # every rank keeps its dataset object current; only rank 0 writes {nth_query+1}_query/queries.pkl
            self.dataloader.dataset.label_queries(queries, nth_query + 1 if ppdist.rank() == 0 else None)
            if nth_query == n_stages - 1:
                break

    def train_step(self, model, optimizer, dict_data, reducer=None):
        """model.py:103-129 for one batch; returns (loss tensor, labels, predictions at the labelled pixels)."""
        if "_ready" in dict_data:  # a batch made on the device by _device_augment, possibly on the copy stream
            cur = torch.cuda.current_stream()
            cur.wait_event(dict_data["_ready"])
            for k in ("x", "y", "queries"):
                dict_data[k].record_stream(cur)
        x = dict_data["x"].to(self.device, non_blocking=True)
        y = dict_data["y"].to(self.device, non_blocking=True)
        mask = dict_data["queries"].to(self.device, torch.bool) if self.n_pixels_by_us != 0 else None
        lowres = model.forward_lowres(x)
        loss, pred_at, (_, _, px_label) = sparse_cross_entropy(lowres, y, mask, self.ignore_index, return_pred=True)
        if reducer is not None:
            reducer.zero_grad()  # gradients are views of the reducer's flat buffer
        else:
            optimizer.zero_grad(set_to_none=True)
        if ppdist.world() > 1:
            n_local = torch.tensor(float(px_label.numel()), device=self.device)
            (loss * ppdist.global_mean_loss_scale(n_local)).backward()
            reducer()  # bucketed all-reduce, launched from backward hooks; this waits for the tail
        else:
            loss.backward()
        optimizer.step()
        return loss.detach(), px_label, pred_at

    def _device_augment(self, dict_data):
        """--gpu_augment: a batch of RAW samples ({'x_raw' uint8 [B,H,W,3], 'y_raw' uint8 [B,H,W], 'queries_raw' uint8}) ->
        the batch the reference's dataset would have delivered (base_dataset.py:174-183: joint geometric augmentation,
        photometric augmentation, to_tensor + normalize), made on the DEVICE (augment.GpuAugment; the draws come from the
        streams the reference uses, per sample in its order).  Runs on the graph's copy stream when there is one, so it
        overlaps the step that is replaying.  Other batches pass through unchanged."""
        if dict_data is None or "x_raw" not in dict_data:
            return dict_data
        xr = dict_data["x_raw"]
        if self._gpu_aug is None:
            from .augment import GpuAugment
            aug = getattr(self.args, "augmentations", None) or {}
            self._gpu_aug = GpuAugment(tuple(xr.shape[1:3]), self.args.mean, self.args.std, self.ignore_index,
                                       geometric_flags=aug.get("geometric"), photometric_flags=aug.get("photometric"))
        stream = self._graph.copy_stream if self._graph is not None else torch.cuda.current_stream()
        with torch.cuda.stream(stream):
            up = lambda t: None if t is None else t.to(self.device, non_blocking=True).contiguous()
            x, y, q, _ = self._gpu_aug(up(xr), up(dict_data.get("y_raw")), up(dict_data.get("queries_raw")), None)
            ready = torch.cuda.Event()
            ready.record(stream)
        out = {k: v for k, v in dict_data.items() if not k.endswith("_raw")}
        out.update(x=x, y=y, queries=q if q is not None else torch.ones_like(y), _ready=ready)
        return out

    def _graphed_step(self, model, optimizer, dict_data, reducer):
        """The same step replayed from ONE captured CUDA graph (graph.py): at the reference batch of 4 the eager step is
        ~900 kernel launches and CPU-bound (4.8x slower).  Captured lazily for the loop's batch shape; other shapes (a
        ragged last batch) take the eager step.  Returns None when this batch cannot use the graph."""
        x, y, q = dict_data["x"], dict_data["y"], dict_data["queries"]
        shape = (x.shape[0], x.shape[2], x.shape[3])
        if "_ready" in dict_data:  # made by _device_augment on some stream: order the current one (and through it the copy stream) after it
            torch.cuda.current_stream().wait_event(dict_data["_ready"])
        gs = self._graph
        if gs is None:
            from .graph import GraphedTrainStep
            # labelled pixels per crop: the reference's NEAREST rescale (x0.5-2.0) + crop of the `queries` mask
            # (base_dataset.py:55-98) can multiply a crop's count by up to 4 -> capacity with that factor, capped at the
            # crop size; a batch that still overflows takes the eager step (below) instead of aborting the epoch
            per_img = min(shape[1] * shape[2], 4 * (self.init_n_pixels + self.max_budget + self.n_pixels_by_us) + 16)
            gs = GraphedTrainStep(model, optimizer, shape, self.ignore_index, capacity=shape[0] * per_img,
                                  device=self.device, reducer=reducer, n_classes=self.n_classes)
            if gs.prefetch(x, y, q) is None:
                return None  # nothing captured yet: try again with the next batch
            gs.commit()
            gs.capture(restore_state=True)  # warm-up steps leave no trace: the first replay is the first update
            self._graph, self._graph_shape = gs, shape
            self._graph_labels = None
            gs()
            return True
        if shape != self._graph_shape:
            gs.drop_staged()
            self._graph_labels = None
            return None
        if self._graph_labels is None and gs.prefetch(x, y, q) is None:
            return None  # more labelled pixels than the captured capacity: eager step for this batch
        self._graph_labels = None
        gs.commit()
        gs()  # loss / predictions / confusion matrix stay on the device until the end of the epoch
        return True

    def _prefetch_next(self, dict_data):
        """upload the NEXT batch while the graph of the current one runs (copy stream, double-buffered staging)."""
        gs = self._graph
        if gs is None or dict_data is None:
            return
        x = dict_data["x"]
        if "_ready" in dict_data:
            torch.cuda.current_stream().wait_event(dict_data["_ready"])
        if (x.shape[0], x.shape[2], x.shape[3]) == self._graph_shape:
            # None = the batch overflows the captured capacity: nothing staged, _graphed_step will send it to the eager step
            self._graph_labels = True if gs.prefetch(x, dict_data["y"], dict_data["queries"]) is not None else None

    def _train_epoch(self, epoch, model, optimizer, lr_scheduler, reducer=None):
        if self.n_pixels_by_us != 0:
            print(f"training an epoch {epoch} of {self.nth_query}th query "
                  f"({self.dataloader.dataset.n_pixels_total} labelled pixels)")
        model.train()
        miou = pixel_acc = float("nan")
        sampler = getattr(self.dataloader, "sampler", None)
        if hasattr(sampler, "set_epoch"):  # multi-GPU: per-rank shards, reshuffled every epoch
            sampler.set_epoch(epoch + 1000 * max(self.nth_query, 0))
        it = iter(self.dataloader)
        dict_data = self._device_augment(next(it, None))
        while dict_data is not None:
            graphed = self._graphed_step(model, optimizer, dict_data, reducer) if self._use_graph else None
            if graphed is None:  # eager step: metrics from the labelled pixels, read back per step (model.py:124-136)
                loss, labels, preds = self.train_step(model, optimizer, dict_data, reducer)
                self.running_score.update_pairs(labels.cpu().numpy(), preds.cpu().numpy())
                self.running_loss.update(loss.item())
            nxt = self._device_augment(next(it, None))
            if graphed:
                self._prefetch_next(nxt)  # H2D of the next batch overlaps this batch's graph
            dict_data = nxt
            if self.lr_scheduler_type == "Poly":
                lr_scheduler.step(epoch=epoch - 1)
            if self.debug:
                break
        self._graph_labels = None
        if self._graph is not None:
            self._graph._staged = False  # a batch prefetched before a debug break is dropped
            if self._graph.metrics is not None:  # ONE device->host read per epoch
                conf, loss_sum, n_steps = self._graph.metrics.read()
                self._graph.metrics.reset()
                self.running_score.update_confusion(conf)
                if n_steps:
                    self.running_loss.update(loss_sum / n_steps, weight=n_steps)
        scores = self.running_score.get_scores()[0]
        miou, pixel_acc = scores["Mean IoU"], scores["Pixel Acc"]
        if self.lr_scheduler_type == "MultiStepLR":
            lr_scheduler.step(epoch=epoch - 1)
        print(f"({self.experim_name}) Epoch {epoch} | mIoU.: {miou:.3f} | pixel acc.: {pixel_acc:.3f} | "
              f"avg loss: {self.running_loss.avg:.3f}")
        if ppdist.rank() == 0:
            write_log(self.log_train, list_entities=[epoch, miou, pixel_acc, self.running_loss.avg])
        self._reset_meters()
        return model, optimizer, lr_scheduler

    def _train(self):
        print(f"\n({self.experim_name}) training...\n")
        model = get_model(self.args).to(self.device)
        model.base_seed = self.args.seed * 131 + self.nth_query + 1
        ppdist.broadcast_parameters(model)
        # whole-step CUDA graph (Adam configs, sparse labels); --no_cuda_graph or SGD / fully-supervised runs stay eager
        self._use_graph = (getattr(self.args, "cuda_graph", True) and optimizer_kind(self.args) == "Adam"
                           and self.n_pixels_by_us != 0)
        self._graph, self._graph_labels = None, None
        optimizer = get_optimizer(self.args, model, capturable=self._use_graph)
        lr_scheduler = get_lr_scheduler(self.args, optimizer=optimizer, iters_per_epoch=len(self.dataloader))
        reducer = ppdist.GradAllReducer(model) if ppdist.world() > 1 else None
        for e in range(1, 1 + self.n_epochs):
            model, optimizer, lr_scheduler = self._train_epoch(e, model, optimizer, lr_scheduler, reducer)
            # BatchNorm running statistics are per-rank during the epoch (the reference has no SyncBN); average them
            # before anything evaluates the model (validation, best-mIoU checkpoint, the query round that scores image
            # i on rank i % world) so every rank holds the SAME eval-mode network
            ppdist.average_buffers(model)
            self._val(e, model)
            if self.debug:
                break
        self.best_miou = -1.0
        self._graph = None  # the graph holds this round's model / optimiser storage
        return model

    @torch.no_grad()
    def _val(self, epoch, model):
        """model.py:177-239: eval over the validation set, full-map confusion matrix, best-mIoU checkpoint.  The
        confusion matrix is accumulated on the device (eval.confusion_over_loader: micro-batched forward, fused
        upsample + argmax + binning kernel) and read back once."""
        from .eval import confusion_over_loader
        self.running_score.update_confusion(confusion_over_loader(
            model, self.dataloader_val, self.n_classes, self.device, self.dataset_name, self.stride_total, self.debug))
        scores = self.running_score.get_scores()[0]
        miou, pixel_acc = scores["Mean IoU"], scores["Pixel Acc"]
        print(f"({self.experim_name}) Epoch {epoch} | val mIoU: {miou:.3f} | pixel acc.: {pixel_acc:.3f}")
        if ppdist.rank() == 0:
            if miou > self.best_miou:
                d = f"{self.dir_checkpoints}/{self.nth_query}_query" if self.n_pixels_by_us != 0 else \
                    f"{self.dir_checkpoints}/fully_sup"
                os.makedirs(d, exist_ok=True)
                torch.save({"model": model.state_dict()}, f"{d}/best_miou_model.pt")
                self.best_miou = miou
            write_log(self.log_val, list_entities=[epoch, miou, pixel_acc])
        self._reset_meters()

    def _reset_meters(self):
        self.running_loss.reset()
        self.running_score.reset()
