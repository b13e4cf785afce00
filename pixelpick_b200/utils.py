"""Host glue mirrored from the reference `utils/utils.py`, `utils/lr_scheduler.py`, `utils/metrics.py` — the parts
the hot paths' callers need (model factory, optimiser groups, schedulers, meters, CSV log) plus a synthetic
dataset with the reference dataset interface (no real datasets exist offline)."""
import os
import pickle as pkl
from typing import Dict, List

import numpy as np
import torch
from torch.optim.lr_scheduler import _LRScheduler
from torch.utils.data import DataLoader, Dataset

from .deeplab import DeepLab
from .query import QuerySelector


def get_model(args):
    """utils/utils.py:15-51.  network_name 'deeplab' -> MobileNetV2-DeepLabv3+ (reference) ; 'deeplab_rn50' -> the
    ResNet-50(dilated-8) + ASPP(2048, OS8) composition of BASELINE configs 3-5.  'FPN' is out of scope."""
    if args.network_name == "deeplab":
        return DeepLab(args)
    if args.network_name == "deeplab_rn50":
        return DeepLab(args, backbone="resnet")
    raise NotImplementedError(f"network_name={args.network_name}: only the DeepLabv3+ hot path is implemented")


def optimizer_kind(args) -> str:
    """The optimizer class utils/utils.py:112-306 ends up building: `cs` is always Adam, `voc` always SGD."""
    return {"cs": "Adam", "voc": "SGD"}.get(args.dataset_name, args.optimizer_type)


def get_optimizer(args, model, capturable=False):
    """utils/utils.py:112-306 for the deeplab branch.  Mirrored quirks: `cs` is always Adam and `voc` always SGD whatever
    `optimizer_type` says (utils.py:114,141,208-233); the SGD groups are hard-coded (backbone 1e-3, rest 1e-2, weight decay
    5e-4, momentum 0.9 - NOT `optimizer_params`, so VOC's declared 1e-4 never applies); Adam takes lr / weight decay from
    `optimizer_params` with the backbone at lr/10, but its declared eps/betas are not forwarded (torch defaults apply).
    capturable: Adam with tensor learning rates, usable inside a captured CUDA graph (pixelpick_b200/graph.py)."""
    kind = optimizer_kind(args)
    parts = (model.backbone, model.aspp, model.low_level_conv, model.seg_head)
    if kind == "Adam":
        op = args.optimizer_params
        groups = [{"params": m.parameters(), "lr": op["lr"] / 10 if i == 0 else op["lr"], "weight_decay": op["weight_decay"]}
                  for i, m in enumerate(parts)]
        if capturable:
            from .graph import make_capturable_adam
            return make_capturable_adam(groups)
        return torch.optim.Adam(groups, fused=next(model.parameters()).is_cuda)
    if kind == "SGD":
        groups = [{"params": m.parameters(), "lr": 1e-3 if i == 0 else 1e-2, "weight_decay": 5e-4, "momentum": 0.9}
                  for i, m in enumerate(parts)]
        return torch.optim.SGD(groups)
    raise ValueError(args.optimizer_type)


class Poly(_LRScheduler):
    """utils/lr_scheduler.py:4-21 — per-iteration polynomial decay, called as step(epoch=epoch-1)."""

    def __init__(self, optimizer, num_epochs, iters_per_epoch, warmup_epochs=0, last_epoch=-1):
        self.iters_per_epoch = iters_per_epoch
        self.cur_iter = 0
        self.N = num_epochs * iters_per_epoch
        self.warmup_iters = warmup_epochs * iters_per_epoch
        # capturable Adam keeps its learning rates as DEVICE tensors: read the base values back ONCE here instead of one
        # blocking .item() per parameter group per iteration (the graph path exists to remove per-step syncs)
        self._host_base_lrs = [float(g.get("initial_lr", g["lr"])) for g in optimizer.param_groups]
        super().__init__(optimizer, last_epoch)

    def get_lr(self):
        T = self.last_epoch * self.iters_per_epoch + self.cur_iter
        factor = pow((1 - 1.0 * T / self.N), 0.9)
        if self.warmup_iters > 0 and T < self.warmup_iters:
            factor = 1.0 * T / self.warmup_iters
        self.cur_iter %= self.iters_per_epoch
        self.cur_iter += 1
        assert factor >= 0, "error in lr_scheduler"
        return [base_lr * factor for base_lr in self._host_base_lrs]


def get_lr_scheduler(args, optimizer, iters_per_epoch=-1):
    """utils/utils.py:309-335 (`voc` is always Poly, utils.py:323-325)."""
    if args.lr_scheduler_type == "MultiStepLR" and args.dataset_name != "voc":
        return torch.optim.lr_scheduler.MultiStepLR(optimizer, milestones=[20, 40], gamma=0.1)
    return Poly(optimizer, args.n_epochs, iters_per_epoch)


def write_log(fp, list_entities=None, header=None):
    """utils/utils.py:66-72: a header truncates the file, rows append; plain comma-joined str() values, "\n" line ends."""
    with open(fp, "w" if header is not None else "a") as f:
        for row in (header, list_entities):
            if row is not None:
                f.write(",".join(str(e) for e in row) + "\n")


class AverageMeter:
    """utils/metrics.py:85-126."""

    def __init__(self):
        self.reset()

    def reset(self):
        self.val = self.avg = self.sum = self.count = 0

    def update(self, val, weight=1):
        self.val = val
        self.sum += val * weight
        self.count += weight
        self.avg = self.sum / self.count


class RunningScore:
    """utils/metrics.py:162-207: confusion matrix over pixels whose label is < n_classes.  `update_pairs` takes the
    (label, prediction) pairs of the labelled pixels only — identical counts, since `_fast_hist` drops every
    ignore_index pixel (utils/metrics.py:168-173) and the train loop sets unlabelled pixels to ignore_index."""

    def __init__(self, n_classes):
        self.n_classes = n_classes
        self.confusion_matrix = np.zeros((n_classes, n_classes))

    def _fast_hist(self, lt, lp):
        n = self.n_classes
        mask = (lt >= 0) & (lt < n)
        return np.bincount(n * lt[mask].astype(int) + lp[mask], minlength=n ** 2).reshape(n, n)

    def update(self, label_trues, label_preds):
        for lt, lp in zip(label_trues, label_preds):
            self.confusion_matrix += self._fast_hist(lt.flatten(), lp.flatten())

    def update_pairs(self, labels, preds):
        self.confusion_matrix += self._fast_hist(np.asarray(labels).flatten(), np.asarray(preds).flatten())

    def update_confusion(self, confusion):
        """add a confusion matrix accumulated elsewhere (the on-device accumulator of the captured train step)."""
        self.confusion_matrix += np.asarray(confusion, dtype=np.float64)

    def get_scores(self):
        hist = self.confusion_matrix
        with np.errstate(divide="ignore", invalid="ignore"):
            acc = np.diag(hist).sum() / hist.sum()
            acc_cls = np.nanmean(np.diag(hist) / hist.sum(axis=1))
            iu = np.diag(hist) / (hist.sum(axis=1) + hist.sum(axis=0) - np.diag(hist))
            freq = hist.sum(axis=1) / hist.sum()
        return ({"Pixel Acc": acc, "Mean Acc": acc_cls, "FreqW Acc": (freq[freq > 0] * iu[freq > 0]).sum(),
                 "Mean IoU": np.nanmean(iu)}, dict(zip(range(self.n_classes), iu)))

    def reset(self):
        self.confusion_matrix = np.zeros((self.n_classes, self.n_classes))


class SyntheticDataset(Dataset):
    """Reference dataset interface (datasets/base_dataset.py:18-46,151-189) over seeded synthetic tensors:
    items {'x','y','queries','p_img'}; `.queries` (list of bool [H,W]), `.n_pixels_total`, `.label_queries`."""

    def __init__(self, args, n_images, size, val=False, query=False, seed=0):
        self.H, self.W = size
        self.n, self.val, self.query = n_images, val, query
        self.n_classes, self.ignore_index = args.n_classes, args.ignore_index
        self.dir_checkpoints = f"{args.dir_root}/checkpoints/{args.experim_name}"
        self.seed = seed + (10_000 if val else 0)
        # --gpu_augment: TRAIN items are the raw sample (uint8 image HWC, label map, query mask as 0 / 255) - what the reference's
        # __getitem__ holds before base_dataset.py:174-183; the loop augments and normalises the batch on the device
        self.raw = bool(getattr(args, "gpu_augment", False)) and not val and not query
        self.mean, self.std = getattr(args, "mean", None), getattr(args, "std", None)
        self.list_labelled_queries = None
        self.list_inputs = None  # optional re-targeting: paths "synthetic/<id>.png" -> the images with those ids, in that order
        rs = np.random.RandomState(self.seed)
        n_init = args.n_pixels_by_us if args.n_pixels_by_us > 0 else 0
        self.queries: List[np.ndarray] = []
        for _ in range(n_images):  # initial random queries (camvid.py:50-96)
            q = np.zeros(self.H * self.W, dtype=bool)
            if n_init:
                q[rs.choice(self.H * self.W, n_init, replace=False)] = True
            self.queries.append(q.reshape(self.H, self.W))
        self.n_pixels_total = int(sum(q.sum() for q in self.queries))

    def __len__(self):
        return self.n if self.list_inputs is None else len(self.list_inputs)

    def update_labelled_queries(self, labelled_queries):
        """datasets/base_dataset.py:143-149: human-labelled maps (ignore_index where unlabelled), one per entry of
        `list_inputs` (the human-in-the-loop flow of query.py:389-412 / train.py:214-226 re-targets the dataset at the
        annotated images first)."""
        self.list_labelled_queries = labelled_queries

    def _image_id(self, i):
        if self.list_inputs is None:
            return i
        return int(os.path.splitext(os.path.basename(self.list_inputs[i]))[0])

    def _xy(self, i):
        g = torch.Generator().manual_seed(self.seed * 1_000_003 + i)
        y = torch.randint(0, self.n_classes, (self.H // 16, self.W // 16), generator=g)
        y = y.repeat_interleave(16, 0).repeat_interleave(16, 1)  # blocky "segments"
        x = torch.randn((3, self.H, self.W), generator=g) * 0.5 + (y.float() / self.n_classes - 0.5)[None]
        void = torch.rand((self.H, self.W), generator=g) < 0.01
        y = y.clone()
        y[void] = self.ignore_index
        return x, y

    def __getitem__(self, i):
        j = self._image_id(i)
        x, y = self._xy(j)
        if self.raw:
            mean = torch.tensor(self.mean if self.mean is not None else [0.5] * 3)[:, None, None]
            std = torch.tensor(self.std if self.std is not None else [0.25] * 3)[:, None, None]
            x_u8 = ((x * std + mean).clamp(0, 1) * 255).round().to(torch.uint8).permute(1, 2, 0).contiguous()  # de-normalised
            return {"x_raw": x_u8, "y_raw": y.to(torch.uint8), "queries_raw": torch.from_numpy(self.queries[j].astype(np.uint8) * 255),
                    "p_img": f"synthetic/{j:06d}.png"}
        d = {"x": x, "y": y, "p_img": f"synthetic/{j:06d}.png"}
        if not self.val:
            d["queries"] = torch.from_numpy(self.queries[j].astype(np.uint8))
        if self.list_labelled_queries is not None:  # base_dataset.py:165-170
            d["labelled_queries"] = torch.from_numpy(np.asarray(self.list_labelled_queries[i]))
        return d

    def label_queries(self, dict_queries: Dict[str, dict], nth_query=None):
        """datasets/base_dataset.py:24-46: OR-merge the new picks, persist the wire-format dict."""
        assert len(dict_queries) == len(self.queries), f"{len(dict_queries)} != {len(self.queries)}"
        new_q = QuerySelector.decode_queries(dict_queries)
        previous = self.n_pixels_total
        self.queries = [np.logical_or(p, q) for p, q in zip(self.queries, new_q)]
        self.n_pixels_total = int(sum(q.sum() for q in self.queries))
        print(f"# labelled pixels is changed from {previous} to {self.n_pixels_total} (delta: {self.n_pixels_total - previous})")
        if isinstance(nth_query, int):
            os.makedirs(f"{self.dir_checkpoints}/{nth_query}_query", exist_ok=True)
            pkl.dump(dict_queries, open(f"{self.dir_checkpoints}/{nth_query}_query/queries.pkl", "wb"))


def get_dataloader(args, batch_size, n_workers, shuffle, val=False, query=False, generate_init_queries=True):
    """utils/utils.py:75-109.  The real datasets (CamVid / Cityscapes / VOC files) are not part of the hot path and
    do not exist offline; `args.synthetic = (n_images, H, W)` selects the synthetic stand-in."""
    if getattr(args, "synthetic", None) is None:
        raise NotImplementedError("dataset readers are out of scope (SURVEY.md §2 #9): pass args.synthetic=(n, H, W) "
                                  "or plug a dataset with the reference interface into Model(dataloaders=...)")
    n, H, W = args.synthetic
    ds = SyntheticDataset(args, n if not val else max(2, n // 4), (H, W), val=val, query=query, seed=args.seed)
    return make_loader(ds, batch_size, n_workers, shuffle, sharded=not val and not query, seed=args.seed)


def make_loader(ds, batch_size, n_workers, shuffle, sharded=False, seed=0):
    """DataLoader with the reference's settings (utils/utils.py:100-108).  Under torch.distributed the TRAIN loader is sharded
    over ranks (disjoint per-rank batches of `batch_size`, reshuffled per epoch through `sampler.set_epoch`); the query and
    validation loaders stay whole - the selector shards the images itself and needs every rank to walk all of them."""
    from . import dist as ppdist
    if sharded and ppdist.world() > 1:
        from torch.utils.data.distributed import DistributedSampler
        sampler = DistributedSampler(ds, num_replicas=ppdist.world(), rank=ppdist.rank(), shuffle=shuffle, seed=seed)
        per_rank = len(sampler)
        return DataLoader(ds, batch_size=batch_size, num_workers=n_workers, sampler=sampler, drop_last=per_rank % batch_size == 1)
    return DataLoader(ds, batch_size=batch_size, num_workers=n_workers, shuffle=shuffle,
                      drop_last=len(ds) % batch_size == 1)
