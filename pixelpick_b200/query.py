"""Host-side mirror of the reference `query.py` (NoelShin/PixelPick @ 43c2981) for the Q path.

Same class names, constructor arguments, return values, files written and NumPy-RNG consumption as
the reference (`QuerySelector`, `UncertaintySampler`, `QueryStats`, `gather/merge_previous_query_files`)
— the per-pixel work (softmax -> uncertainty -> mask fill -> sorted top-k -> random sub-selection)
runs in the hand-written sm_100a kernels of `libpixelpick_b200.so` via `_lib` (no CPU fallback).

Differences that do not change results:
  * images are scored in batches of `batch_imgs` (reference: one at a time, query.py:159); the global
    NumPy stream is still consumed once per image in dataloader order (query.py:40,64);
  * `np.random.choice(ind, n, False)` is evaluated as `ind[np.random.permutation(len(ind))[:n]]`
    (bit-identical, incl. the RNG state afterwards) so only n positions per image leave the device;
  * `QueryStats` evaluates the entropy only at the selected pixels (reference recomputes the full map,
    query.py:260-264);
  * under `torch.distributed` (one process per GPU) image i is scored by rank i % world; every rank still walks the
    whole dataloader and draws every image's random numbers, so for the SAME model the picks are those of a
    single-process run (the caller makes the model identical on every rank: gradients are all-reduced and
    `dist.average_buffers` equalises the BatchNorm running statistics before the round), and the round ends with ONE
    all-gather of the per-rank picks / statistics.
Tie rule of the top-k: equal scores -> lower flat index first (CPU `torch.topk` leaves it unspecified).
"""
import os
import pickle as pkl
from math import ceil
from pathlib import Path
from typing import Dict, List, Tuple, Union

import numpy as np
import torch
import torch.nn.functional as F

from . import _lib
from . import dist as ppdist


class UncertaintySampler:
    """query.py:224-247.  `prob` is a [b, c, h, w] probability map (softmax output)."""

    def __init__(self, query_strategy):
        self.query_strategy = query_strategy

    @staticmethod
    def _from_prob(prob, strategy):
        # softmax(log p) == p, so the fused kernel applied to log-probabilities gives the reference
        # quantities; log(0) = -inf reproduces the entropy NaN of query.py:230.
        return _lib.acq_score(torch.log(prob.float()), strategy)

    @staticmethod
    def _entropy(prob):
        return UncertaintySampler._from_prob(prob, "entropy")

    @staticmethod
    def _least_confidence(prob):
        return UncertaintySampler._from_prob(prob, "least_confidence")

    @staticmethod
    def _margin_sampling(prob):
        return UncertaintySampler._from_prob(prob, "margin_sampling")

    @staticmethod
    def _random(prob):
        b, _, h, w = prob.shape
        return torch.rand((b, h, w))  # torch CPU generator, as query.py:242-244

    def __call__(self, prob):
        return getattr(self, f"_{self.query_strategy}")(prob)


class QueryStats:
    """query.py:250-308 — label histogram, entropy, unique labels and spatial spread of the picks."""

    def __init__(self, args):
        self.dir_checkpoints = f"{args.dir_root}/checkpoints/{args.experim_name}"
        self.list_entropy, self.list_n_unique_labels, self.list_spatial_coverage = list(), list(), list()
        self.dict_label_cnt = {l: 0 for l in range(args.n_classes)}

    def begin_round(self):
        """The lists are cumulative over query rounds, as in the reference (never reset): remember where this round starts."""
        self._round_start = (len(self.list_entropy), len(self.list_n_unique_labels), dict(self.dict_label_cnt))

    def round_payload(self, image_ids: List[int]):
        """Multi-GPU: this rank's records of the current round - `image_ids` are the dataloader positions of the images it
        updated since `begin_round`, in update order.  Travels in the round's all-gather next to the picks."""
        e0, u0, cnt0 = self._round_start
        ent, uniq, cov = self.list_entropy[e0:], self.list_n_unique_labels[u0:], self.list_spatial_coverage[u0:]
        n = len(image_ids)
        assert n == len(uniq) == len(cov) and (len(ent) % n == 0 if n else not ent)
        per = len(ent) // n if n else 0
        return ([(i, ent[j * per:(j + 1) * per], uniq[j], cov[j]) for j, i in enumerate(image_ids)],
                {l: c - cnt0[l] for l, c in self.dict_label_cnt.items()})

    def absorb_round(self, payloads):
        """Replace this round's local records by every rank's, in dataloader order: `save` then writes what a single process
        would have written."""
        e0, u0, cnt0 = self._round_start
        records, counts = [], dict(cnt0)
        for recs, delta in payloads:
            records.extend(recs)
            for l, c in delta.items():
                counts[l] += c
        records.sort(key=lambda r: r[0])
        self.list_entropy[e0:] = [e for r in records for e in r[1]]
        self.list_n_unique_labels[u0:] = [r[2] for r in records]
        self.list_spatial_coverage[u0:] = [r[3] for r in records]
        self.dict_label_cnt = counts

    def _count_labels(self, labels):
        for l in labels:
            self.dict_label_cnt[l] += 1

    @staticmethod
    def _spatial_coverage(ys, xs):
        a, b = ys[:, None].astype(np.int64), xs[:, None].astype(np.int64)
        dist = np.sqrt((a - a.T) ** 2 + (b - b.T) ** 2)
        try:
            dist = dist[~np.eye(dist.shape[0], dtype=bool)].reshape(dist.shape[0], -1).mean()
        except ValueError:
            return np.nan
        return dist

    def update_selected(self, flat_idx_sorted, width, y, entropies):
        """`flat_idx_sorted`: row-major sorted picks of one image; y: [h, w] labels (host);
        entropies: entropy at those pixels in the same order."""
        labels = y.reshape(-1)[flat_idx_sorted]
        self._count_labels(labels)
        self.list_entropy.extend([float(e) for e in entropies])
        self.list_n_unique_labels.append(len(set(labels.tolist())))
        self.list_spatial_coverage.append(self._spatial_coverage(flat_idx_sorted // width, flat_idx_sorted % width))

    def update_from_device(self, labels_at, n_unique, coverage, entropies, label_hist):
        """the same bookkeeping from what pp_query_stats_at computed for a whole batch: labels_at [b, n], n_unique [b],
        coverage [b] (float64, NumPy's summation order), entropies [b, n], label_hist [n_classes] (this batch's counts)"""
        n_classes = len(self.dict_label_cnt)
        if (labels_at < 0).any() or (labels_at >= n_classes).any():
            raise KeyError(int(labels_at[(labels_at < 0) | (labels_at >= n_classes)][0]))  # as dict_label_cnt[l] would (query.py:268)
        for l, c in enumerate(label_hist.tolist()):
            self.dict_label_cnt[l] += int(c)
        for i in range(labels_at.shape[0]):
            self.list_entropy.extend([float(e) for e in entropies[i]])
            self.list_n_unique_labels.append(int(n_unique[i]))
            self.list_spatial_coverage.append(np.float64(coverage[i]))

    def save(self, nth_query):
        dict_stats = {
            "label_distribution": self.dict_label_cnt,
            "avg_entropy": np.mean(self.list_entropy),
            "avg_n_unique_labels": np.mean(self.list_n_unique_labels),
            "avg_spatial_coverage": np.mean(self.list_spatial_coverage),
        }
        for k, v in dict_stats.items():
            print(f"{k}: {v}")
        os.makedirs(f"{self.dir_checkpoints}/{nth_query}_query", exist_ok=True)
        pkl.dump(dict_stats, open(f"{self.dir_checkpoints}/{nth_query}_query/query_stats.pkl", "wb"))


class _AsyncDraws:
    """The random ranks of EVERY image of a query round, drawn in dataloader order from the global NumPy stream (query.py:40,64)
    on a background thread while the main thread loads and scores this rank's images.  The sizes the draws depend on are known
    up-front on every rank: `dataset.queries` holds one [h, w] mask per image.  NumPy's legacy shuffle releases the GIL, so the
    ~60 us per image (a full Fisher-Yates pass over k = 5 % of the pixels, needed to leave the stream where the reference
    leaves it) overlap the host loop instead of adding world x n_own x 60 us to every rank's round."""

    def __init__(self, selector, shapes):
        import threading
        self.shapes, self.pos, self.done, self.err = shapes, [None] * len(shapes), 0, None
        self.cv = threading.Condition()
        self.selector = selector
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def _run(self):
        try:
            for i, (h, w) in enumerate(self.shapes):
                _, p = self.selector._draw_positions(1, h, w)
                with self.cv:
                    self.pos[i] = p
                    self.done = i + 1
                    self.cv.notify_all()
        except BaseException as e:  # surfaced by get() / join() on the main thread
            with self.cv:
                self.err = e
                self.cv.notify_all()

    def get(self, i):
        with self.cv:
            while self.done <= i and self.err is None:
                self.cv.wait()
            if self.err is not None:
                raise self.err
            return self.pos[i]

    def join(self):
        self.thread.join()
        if self.err is not None:
            raise self.err


class _Slot:
    """Host staging of one batch (pinned when a GPU is present): images and masks are written straight into these buffers as
    the loader yields them (no torch.cat, no pageable copies) and go up with ONE async copy each.  Two slots alternate, so
    the host fills batch i+1 while the GPU works on batch i."""

    def __init__(self, n, x_shape, hw, pin):
        def mk(shape, dt):
            return torch.empty(shape, dtype=dt, pin_memory=pin)
        self.x = mk((n,) + tuple(x_shape), torch.float32)
        self.lab = mk((n,) + tuple(hw), torch.uint8)
        self.void = mk((n,) + tuple(hw), torch.uint8)
        self.event = None


class _Batch:
    """one batch in flight: its items (host bookkeeping), staging slot, device intermediates and pinned results"""

    def __init__(self, slot, hw, x_shape):
        self.items, self.slot, self.hw, self.x_shape = [], slot, hw, x_shape
        self.score = self.handle = self.ws = None
        self.n_top = 0
        self.stats_on = False
        self.sel_host = self.ent_host = self.done = None
        self.lab8 = self.dev_stats = self.pos_host = None


class QuerySelector:
    """Drop-in for the reference QuerySelector (query.py:12-221)."""

    def __init__(self, args, dataloader, device=torch.device("cuda:0"), batch_imgs: int = 32):
        self.dataset_name = args.dataset_name
        self.dataloader = dataloader
        self.debug = args.debug
        self.device = torch.device(device)
        self.dir_checkpoints = f"{args.dir_root}/checkpoints/{args.experim_name}"
        self.ignore_index = args.ignore_index
        self.mc_n_steps = args.mc_n_steps
        self.n_classes = args.n_classes
        self.n_pixels_by_us = args.n_pixels_by_us
        self.network_name = args.network_name
        self.query_stats = QueryStats(args)
        self.query_strategy = args.query_strategy
        self.reverse_order = args.reverse_order
        self.stride_total = args.stride_total
        self.top_n_percent = args.top_n_percent
        self.uncertainty_sampler = UncertaintySampler(args.query_strategy)
        self.use_mc_dropout = args.use_mc_dropout
        self.vote_type = args.vote_type
        self.batch_imgs = batch_imgs
        self._ws = {}
        # label maps travel as uint8 when every label (and ignore_index) fits: void mask + QueryStats come from them on the device
        self._labels_u8 = 0 <= args.ignore_index <= 255 and args.n_classes <= 256
        if self.device.type != "cuda":
            raise _lib.PixelPickError("QuerySelector needs a CUDA device (no CPU fallback)")

    # ---- wire format (query.py:72-142), unchanged -------------------------------------------------
    @staticmethod
    def encode_query(p_img: str, size: Tuple[int, int], query: np.ndarray) -> Dict[str, dict]:
        y_coords, x_coords = np.where(query)
        return {p_img: {"height": size[0], "width": size[1], "x_coords": x_coords, "y_coords": y_coords}}

    @staticmethod
    def decode_queries(encoded_query: Dict[str, dict], ignore_index: int = 255, return_as_dict: bool = False
                       ) -> Union[List[np.ndarray], Dict[str, np.ndarray]]:
        def decode_query(info: dict) -> np.ndarray:
            labels = info.get("category_id", None)
            if labels is None:
                q = np.zeros((info["height"], info["width"]), dtype=bool)
                q[info["y_coords"], info["x_coords"]] = True
            else:
                q = ignore_index * np.ones((info["height"], info["width"]), dtype=np.int64)
                for i, loc in enumerate(zip(info["y_coords"], info["x_coords"])):  # later duplicates win
                    q[loc] = labels[i]
            return q

        if len(encoded_query) == 0:
            raise ValueError(len(encoded_query))
        items = sorted(encoded_query.items()) if len(encoded_query) > 1 else list(encoded_query.items())
        if return_as_dict:
            return {p: decode_query(info) for p, info in items}
        return [decode_query(info) for _, info in items]

    # ---- selection of one batch of score maps ------------------------------------------------------
    def _k(self, h, w):
        return int(h * w * self.top_n_percent) if self.top_n_percent > 0.0 else self.n_pixels_by_us

    def _workspace(self, n, hw, k):
        key = (n, hw, k)
        if key not in self._ws:
            self._ws = {key: _lib.TopKWorkspace(n, hw, k, self.device)}  # keep only the current shape
        return self._ws[key]

    def _select_queries(self, uc_map) -> np.ndarray:
        """query.py:33-69 for ONE [h, w] score map that is already masked; returns bool [h, w]."""
        h, w = uc_map.shape[-2:]
        uc = uc_map.reshape(1, h * w).to(self.device, torch.float32)
        k = self._k(h, w)
        largest = _lib.LARGEST[self.query_strategy]
        if self.reverse_order:
            assert self.top_n_percent > 0.0
            ind = np.random.permutation(h * w)[:k]  # == np.random.choice(range(h*w), k, False)
            keep = torch.zeros(h * w, dtype=torch.bool)
            keep[torch.from_numpy(ind)] = True
            uc = uc.clone()
            uc[0, ~keep.to(self.device)] = _lib.FILL[self.query_strategy]
            ind_queries = _lib.acq_topk(uc, self.n_pixels_by_us, largest).cpu().numpy()[0]
        else:
            topk = _lib.acq_topk(uc, k, largest)
            if self.top_n_percent > 0.0:
                pos = torch.from_numpy(np.random.permutation(k)[: self.n_pixels_by_us].astype(np.int32))[None]
                ind_queries = _lib.acq_gather(topk, pos).cpu().numpy()[0]
            else:
                ind_queries = topk.cpu().numpy()[0]
        query = np.zeros(h * w, dtype=bool)
        query[ind_queries] = True
        return query.reshape(h, w)

    def _draw_positions(self, n_img, h, w):
        """Consume the global NumPy stream exactly like query.py:40,64 — once per image, in order."""
        k = self._k(h, w)
        keep, pos = None, None
        if self.reverse_order:
            assert self.top_n_percent > 0.0
            keep = np.zeros((n_img, h * w), dtype=bool)
            for i in range(n_img):
                keep[i, np.random.permutation(h * w)[:k]] = True
        elif self.top_n_percent > 0.0:
            pos = np.stack([np.random.permutation(k)[: self.n_pixels_by_us] for _ in range(n_img)]).astype(np.int32)
        return keep, pos

    def _score_batch(self, model, x, h, w, labelled, void, keep, ws):
        """model forward + fused scoring for a [b, 3, H', W'] batch -> (score [b, h*w], logits handle)."""
        st = self.query_strategy
        lowres = getattr(model, "forward_lowres", None)
        # the fused upsample+score kernel is instantiated for the reference datasets' class counts (cv 11, cs 19, voc 21);
        # any other n_classes (a --p_dataset_config dataset) takes the full-resolution path below, whose scalar scoring
        # kernel handles every C
        if lowres is not None and x.shape[2] == h and x.shape[3] == w and self.n_classes in _lib.UPSAMPLED_SCORE_CLASSES:
            lr = lowres(x)  # [b, C, h/4, w/4] fp32: the ×4 upsample is fused into the kernel
            score = _lib.acq_score_upsampled(lr, (h, w), st, labelled, void, keep, hist0_ws=ws)
            return score, ("lowres", lr)
        pred = model(x)["pred"][:, :, :h, :w]  # query.py:190
        if pred.dtype not in (torch.float32, torch.bfloat16):
            pred = pred.float()
        score = _lib.acq_score(pred, st, labelled, void, keep, hist0_ws=ws)
        return score, ("full", pred)

    # ---- batches: pinned staging, asynchronous launch, deferred collection --------------------------------------
    def _slot_for(self, x_shape, hw):
        """one of two alternating staging slots for batches of this geometry (host fills one while the GPU reads the other)"""
        key = (self.batch_imgs, tuple(x_shape), tuple(hw), self.n_pixels_by_us)
        if getattr(self, "_slot_key", None) != key:
            pin = self.device.type == "cuda" and torch.cuda.is_available()
            self._slots = [_Slot(self.batch_imgs, x_shape, hw, pin) for _ in range(2)]
            self._slot_key, self._slot_turn = key, 0
        slot = self._slots[self._slot_turn & 1]
        self._slot_turn += 1
        if slot.event is not None:
            slot.event.synchronize()  # the upload that last used this slot (two batches ago) has read it
        return slot

    def _add(self, batch, item, x, mask, human_labels):
        """write one image's inputs straight into the batch's staging slot"""
        j = len(batch.items)
        s = batch.slot
        s.x[j].copy_(x[0])
        lab = s.lab[j].numpy()
        if human_labels:
            np.not_equal(mask, self.ignore_index, out=lab.view(bool))
        else:
            lab[...] = mask
        if item["y"] is not None:
            if self._labels_u8:  # the label map itself (uint8): void mask AND the labels at the picks come from it on the device
                np.copyto(s.void[j].numpy(), item["y"], casting="unsafe")
            else:
                np.equal(item["y"], self.ignore_index, out=s.void[j].numpy().view(bool))  # query.py:196-201: void pixels
        batch.items.append(item)

    def _launch(self, model, batch, pick):
        """H2D + forward + fused scoring (+ select / pick when the drawn ranks are known) for one batch, all asynchronous."""
        n = len(batch.items)
        h, w = batch.hw
        hw = h * w
        st = self.query_strategy
        s = batch.slot
        dev = self.device
        xs = s.x[:n].to(dev, non_blocking=True)
        labelled = s.lab[:n].to(dev, non_blocking=True).view(torch.bool)
        has_y = batch.items[0]["y"] is not None
        void = None
        batch.lab8 = None
        if has_y and self._labels_u8:
            batch.lab8 = s.void[:n].to(dev, non_blocking=True)
            void = batch.lab8 == self.ignore_index
        elif has_y:
            void = s.void[:n].to(dev, non_blocking=True).view(torch.bool)
        if dev.type == "cuda":
            if s.event is None:
                s.event = torch.cuda.Event()
            s.event.record()
        if self.dataset_name == "voc":  # query.py:171-174
            pad_h = ceil(h / self.stride_total) * self.stride_total - h
            pad_w = ceil(w / self.stride_total) * self.stride_total - w
            xs = F.pad(xs, pad=(0, pad_w, 0, pad_h), mode="reflect")
        keep_np = None if batch.items[0]["keep"] is None else np.concatenate([b["keep"] for b in batch.items])
        keep = None if keep_np is None else torch.from_numpy(keep_np).to(dev)
        batch.n_top = self.n_pixels_by_us if self.reverse_order else self._k(h, w)
        if st == "random":
            uc = torch.stack([b["rand"] for b in batch.items]).to(dev)
            excl = labelled if void is None else (labelled | void)
            if keep is not None:
                excl = excl | ~keep.view(n, h, w)
            uc[excl] = _lib.FILL[st]
            batch.score, batch.handle, batch.ws = uc.view(n, hw), None, None
        else:
            # immediate mode: one workspace per geometry (stream order protects it); deferred picks: one per batch, its
            # level-0 histogram must survive until the ranks are drawn
            batch.ws = self._workspace(n, hw, batch.n_top) if pick else _lib.TopKWorkspace(n, hw, batch.n_top, dev)
            batch.ws.prepare()
            score, batch.handle = self._score_batch(model, xs, h, w, labelled, void, keep, batch.ws)
            batch.score = score.view(n, hw)
        if pick:
            self._pick(batch)

    def _pick(self, batch):
        """select + order statistics at the drawn ranks, entropy at the picks, results -> pinned host buffers (async)"""
        n = len(batch.items)
        h, w = batch.hw
        st = self.query_strategy
        largest = _lib.LARGEST[st]
        pos_np = None if batch.items[0]["pos"] is None else np.concatenate([b["pos"] for b in batch.items])
        pos_t = None
        if pos_np is not None:
            if batch.score.is_cuda:  # through pinned memory: a pageable upload would wait for the forward queued ahead of it
                batch.pos_host = torch.empty(pos_np.shape, dtype=torch.int32, pin_memory=True)
                batch.pos_host.numpy()[...] = pos_np
                pos_t = batch.pos_host.to(batch.score.device, non_blocking=True)
            else:
                pos_t = torch.from_numpy(pos_np)
        # only the n drawn ranks of the sorted top-k are needed (query.py:63-64): radix pick, no sort
        if batch.ws is None:
            sel = _lib.acq_select_pick(batch.score, batch.n_top, largest, pos_t, n=self.n_pixels_by_us)
        else:
            sel = _lib.acq_select_pick(batch.score, batch.n_top, largest, pos_t, n=self.n_pixels_by_us, ws=batch.ws, hist0_valid=True)
        sel, _ = torch.sort(sel.long(), dim=1)  # np.where order: row-major ascending (query.py:77)
        ent = None
        if batch.stats_on and batch.handle is not None:
            kind, t = batch.handle
            ent = _lib.acq_entropy_at_upsampled(t, (h, w), sel) if kind == "lowres" else _lib.acq_entropy_at(t, sel)
        pin = sel.is_cuda
        batch.dev_stats = None
        if batch.stats_on and sel.is_cuda and batch.lab8 is not None:
            # QueryStats + wire coordinates for the whole batch in one launch (pp_query_stats_at), read back with the picks
            hist = torch.zeros(self.n_classes, dtype=torch.int64, device=sel.device)
            outs = _lib.query_stats_at(sel.contiguous(), w, h * w, batch.lab8.view(n, -1), self.n_classes, hist)
            batch.dev_stats = []
            for t in outs + (hist,):
                th = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
                th.copy_(t, non_blocking=True)
                batch.dev_stats.append(th)
        batch.lab8 = None
        batch.sel_host = torch.empty(sel.shape, dtype=sel.dtype, pin_memory=pin)
        batch.sel_host.copy_(sel, non_blocking=True)
        batch.ent_host = None
        if ent is not None:
            batch.ent_host = torch.empty(ent.shape, dtype=ent.dtype, pin_memory=pin)
            batch.ent_host.copy_(ent, non_blocking=True)
        batch.done = torch.cuda.Event() if pin else None
        if batch.done is not None:
            batch.done.record()
        batch.score = batch.handle = batch.ws = None  # device buffers go back to the allocator (stream-ordered)

    def _collect(self, batch, dict_queries):
        """wait for one launched batch and do its host bookkeeping: wire-format entries + QueryStats"""
        if batch.done is not None:
            batch.done.synchronize()
        h, w = batch.hw
        sel_np = batch.sel_host.numpy()
        ent = None if batch.ent_host is None else batch.ent_host.numpy()
        n_new = 0
        if batch.dev_stats is not None:
            xs, ys, lab_at, uniq, cov, hist = [t.numpy() for t in batch.dev_stats]
            for i, b in enumerate(batch.items):
                dict_queries[b["p_img"]] = {"height": h, "width": w, "x_coords": xs[i].copy(), "y_coords": ys[i].copy()}
                n_new += xs[i].size
            self.query_stats.update_from_device(lab_at, uniq, cov, ent if ent is not None else np.full(lab_at.shape, np.nan), hist)
            return n_new
        for i, b in enumerate(batch.items):
            idx = sel_np[i]
            dict_queries[b["p_img"]] = {"height": h, "width": w, "x_coords": idx % w, "y_coords": idx // w}
            n_new += idx.size
            if batch.stats_on:
                e = ent[i] if ent is not None else np.full(idx.size, np.nan)
                self.query_stats.update_selected(idx, w, b["y"], e)
        return n_new

    @staticmethod
    def _batches_of_one(dl):
        """Iterate a batch_size-1 DataLoader.  For the plain case - map-style dataset, sequential order, default collate, no
        workers, items that are dicts of tensors / strings - the items are taken from the dataset directly and given their
        batch axis as a VIEW: the default collate's copy of every tensor (a 1.5 MB image and a 1 MB label map per Cityscapes
        crop, ~0.2 ms of host memcpy per image on a path that is host-bound) is the only thing skipped.  The iterator's one
        draw from torch's global stream (its base seed) is made all the same, so the streams stay where the reference's are."""
        from torch.utils.data import DataLoader, IterableDataset, SequentialSampler
        from torch.utils.data._utils.collate import default_collate
        plain = (isinstance(dl, DataLoader) and not isinstance(dl.dataset, IterableDataset) and dl.batch_size == 1
                 and dl.num_workers == 0 and isinstance(dl.sampler, SequentialSampler) and dl.collate_fn is default_collate
                 and not dl.pin_memory and not dl.drop_last)
        if not plain:
            yield from dl
            return
        torch.empty((), dtype=torch.int64).random_(generator=dl.generator)  # _BaseDataLoaderIter.__init__: the base seed
        ds = dl.dataset
        for i in range(len(ds)):
            d = ds[i]
            if isinstance(d, dict) and all(torch.is_tensor(v) or isinstance(v, str) for v in d.values()):
                yield {k: (v.unsqueeze(0) if torch.is_tensor(v) else [v]) for k, v in d.items()}
            else:
                yield default_collate([d])

    def _own_loader(self, world, rank):
        """multi-GPU: a loader over THIS rank's images only (image i -> rank i % world), built from the caller's loader, so a
        rank reads 1/world of the dataset instead of all of it.  None when the loader cannot be re-targeted (not a map-style
        DataLoader in sequential order with batch_size 1) - the caller then walks the whole loader."""
        from torch.utils.data import DataLoader, IterableDataset, SequentialSampler, Subset
        dl = self.dataloader
        if not isinstance(dl, DataLoader) or isinstance(dl.dataset, IterableDataset) or dl.batch_size != 1 \
                or not isinstance(dl.sampler, SequentialSampler):
            return None
        idx = list(range(rank, len(dl.dataset), world))
        sub = DataLoader(Subset(dl.dataset, idx), batch_size=1, shuffle=False, num_workers=dl.num_workers,
                         collate_fn=dl.collate_fn, pin_memory=dl.pin_memory)
        return sub, idx

    def __call__(self, nth_query, model, human_labels: bool = False):
        if human_labels:
            prev_queries = self.dataloader.dataset.list_labelled_queries
        else:
            prev_queries = self.dataloader.dataset.queries
        model.eval()
        if self.use_mc_dropout:
            # the reference branch is dead code: `up_map` NameError at query.py:186
            raise NotImplementedError("use_mc_dropout: the reference implementation raises NameError (query.py:186)")
        print(f"Choosing pixels by {self.query_strategy}")
        n_pixels, n_imgs = 0, 0
        dict_queries: dict = dict()
        world, rank = ppdist.world(), ppdist.rank()
        order: List[str] = []      # every image path in dataloader order
        mine: List[int] = []       # dataloader positions of the images this rank scores
        self.query_stats.begin_round()
        # Multi-GPU fast path: every rank loads and scores ONLY its own images; the random ranks - which the reference draws
        # from the global NumPy stream once per image in dataloader order (query.py:40,64) and which depend on every image's
        # size - are drawn for ALL images on a background thread (_AsyncDraws: the sizes come from the per-image masks every
        # rank holds).  The draws are the same numbers in the same order as in a single process (a query dataset draws nothing
        # itself: base_dataset.py:172-181).  reverse_order / random need their draws BEFORE scoring and keep the
        # walk-everything path.
        own = None
        if world > 1 and not self.reverse_order and self.query_strategy != "random":
            own = self._own_loader(world, rank)
        any_y = [False]
        launched: List[_Batch] = []   # launched, not yet collected (immediate mode: at most one in flight)
        scored: List[_Batch] = []     # deferred mode: scored, waiting for their ranks
        cur = [None]

        drawer = [None]

        def close_batch(pick):
            b = cur[0]
            cur[0] = None
            if b is None or not b.items:
                return
            if drawer[0] is not None:
                for it in b.items:
                    it["pos"] = drawer[0].get(it["gpos"])
            b.stats_on = not human_labels and b.items[0]["y"] is not None
            self._launch(model, b, pick)
            if pick:
                launched.append(b)
                while len(launched) > 1:  # collect the PREVIOUS batch while this one runs
                    nonlocal_counts[0] += self._collect(launched.pop(0), dict_queries)
            else:
                scored.append(b)

        nonlocal_counts = [0]

        def feed(batch_ind, dict_data, keep, pos, rand, pick):
            x = dict_data["x"]
            y = dict_data.get("y", None)
            h, w = tuple(x.shape[2:])
            if y is not None:
                y = y.squeeze(dim=0).numpy()
                any_y[0] = True
            item = {"y": y, "hw": (h, w), "p_img": dict_data["p_img"][0], "keep": keep, "pos": pos, "rand": rand, "gpos": batch_ind}
            b = cur[0]
            if b is not None and (b.hw != (h, w) or b.x_shape != tuple(x.shape[1:]) or len(b.items) == self.batch_imgs
                                  or (b.items[0]["y"] is None) != (y is None)):
                close_batch(pick)
                b = None
            if b is None:
                b = cur[0] = _Batch(self._slot_for(x.shape[1:], (h, w)), (h, w), tuple(x.shape[1:]))
            self._add(b, item, x, np.asarray(prev_queries[batch_ind]), human_labels)
            mine.append(batch_ind)

        with torch.no_grad():
            if own is None:
                for batch_ind, dict_data in enumerate(self._batches_of_one(self.dataloader)):
                    h, w = tuple(dict_data["x"].shape[2:])
                    # the random numbers of EVERY image are drawn here, in dataloader order, on every rank (query.py:40,64
                    # and UncertaintySampler._random): the streams, hence the picks, do not depend on the world size
                    keep, pos = self._draw_positions(1, h, w)
                    rand = self.uncertainty_sampler(torch.empty(1, 1, h, w))[0] if self.query_strategy == "random" else None
                    order.append(dict_data["p_img"][0])
                    n_imgs += 1
                    if batch_ind % world != rank:
                        continue
                    feed(batch_ind, dict_data, keep, pos, rand, pick=True)
                close_batch(pick=True)
            else:
                loader, positions = own
                shapes = [tuple(np.shape(q)[-2:]) for q in prev_queries]
                drawer[0] = _AsyncDraws(self, shapes) if self.top_n_percent > 0.0 else None
                for j, dict_data in enumerate(self._batches_of_one(loader)):
                    h, w = tuple(dict_data["x"].shape[2:])
                    if (h, w) != shapes[positions[j]]:
                        raise _lib.PixelPickError(f"image {positions[j]} is {h}x{w} but its query mask is {shapes[positions[j]]}")
                    feed(positions[j], dict_data, None, None, None, pick=True)
                close_batch(pick=True)
                if drawer[0] is not None:
                    drawer[0].join()  # the stream ends where a single process leaves it: every image's draw is consumed
                n_imgs = len(shapes)
            for b in launched:
                nonlocal_counts[0] += self._collect(b, dict_queries)
        n_pixels = nonlocal_counts[0]
        assert n_imgs > 0, "no queries are chosen!"
        stats_on = not human_labels and any_y[0]
        if world > 1:
            # the round's ONE exchange: per-rank picks with their dataloader positions (and statistics) -> every rank
            local_pos = {it_p: g for g, it_p in zip(mine, list(dict_queries))}
            gathered = ppdist.all_gather_objects((dict_queries, local_pos, stats_on,
                                                  self.query_stats.round_payload(mine) if stats_on else None))
            stats_on = any(g[2] for g in gathered)
            merged, where = dict(), dict()
            for d, lp, _, _ in gathered:
                merged.update(d)
                where.update(lp)
            if own is not None:
                order = sorted(merged, key=lambda p_: where[p_])
            dict_queries = {p: merged[p] for p in order}
            n_pixels = sum(len(info["x_coords"]) for info in dict_queries.values())
            if stats_on:
                self.query_stats.absorb_round([g[3] for g in gathered if g[3] is not None])
        if stats_on:
            if rank == 0:
                self.query_stats.save(nth_query)
            print(f"{n_pixels} labelled pixels  are chosen by {self.query_strategy} strategy")
            # every rank merges the picks into its own dataset object; only rank 0 writes {nth_query}_query/queries.pkl
            self.dataloader.dataset.label_queries(dict_queries, nth_query if rank == 0 else None)
        return dict_queries


def gather_previous_query_files(dir_base: str, ext="pkl") -> List[str]:
    """query.py:311-313."""
    return [str(p) for p in Path(dir_base).rglob(f"*/queries.{ext}" if ext is not None else "*")]


def merge_previous_query_files(list_previous_query_files: List[str], ignore_index: int, verbose: bool = True
                               ) -> Dict[str, np.ndarray]:
    """query.py:316-351 — later files overwrite earlier labels of the same pixel."""
    merged: Dict[str, np.ndarray] = dict()
    cnt = 0
    for p in list_previous_query_files:
        decoded = QuerySelector.decode_queries(pkl.load(open(p, "rb")), ignore_index=ignore_index, return_as_dict=True)
        for p_img, q in decoded.items():
            if p_img not in merged:
                merged[p_img] = ignore_index * np.ones_like(q, dtype=np.int64)
            sel = q != ignore_index
            merged[p_img][sel] = q[sel]
            cnt += int(sel.sum())
    if verbose:
        print(f"# merged pixels: {cnt}")
    return merged


def main(argv=None):
    """`python -m pixelpick_b200.query` - the human-in-the-loop query step of the reference (query.py:354-437,
    scripts/query.sh): with --p_state_dict, load the trained model, merge every `*/queries.pkl` under --dir_checkpoints
    (the annotated pixels so far, query.py:311-351), point the query dataset at exactly those images, choose the next
    pixels with `human_labels=True` (labelled pixels = where the merged map differs from ignore_index) and write
    `{dir_checkpoints}/{nth_query}_query/queries.pkl`.  Without --p_state_dict: build the query dataloader with fresh initial
    queries (the dataset constructor writes `0_query/label.npy`, nothing else happens - as in the reference)."""
    from copy import deepcopy
    from torch.utils.data import DataLoader
    from .args import Arguments
    from .utils import get_dataloader, get_model
    parser = Arguments()
    parser.parser.add_argument("--p_state_dict", type=str, default="", help="path to a state_dict file")
    args = parser.parse_args(argv=argv, verbose=True) if argv is not None else parser.parse_args(verbose=True)
    if not torch.cuda.is_available():
        raise _lib.PixelPickError("pixelpick_b200.query needs a CUDA device (no CPU fallback)")
    device = torch.device("cuda", torch.cuda.current_device())
    if args.p_state_dict == "":
        get_dataloader(deepcopy(args), query=True, val=False, generate_init_queries=True, shuffle=False, batch_size=1,
                       n_workers=args.n_workers)
        return None
    model = get_model(args).to(device)
    model.load_state_dict(torch.load(args.p_state_dict, map_location=device)["model"])
    print(f"pretrained model is loaded from {args.p_state_dict}")
    list_prev = gather_previous_query_files(args.dir_checkpoints)
    merged = merge_previous_query_files(list_prev, ignore_index=args.ignore_index)
    dataset = get_dataloader(deepcopy(args), query=True, val=False, generate_init_queries=False, shuffle=False, batch_size=1,
                             n_workers=args.n_workers).dataset
    list_inputs, list_merged = [], []
    for p_img, q in sorted(merged.items()):  # query.py:389-397: the images that carry annotations, by file name
        if getattr(args, "synthetic", None) is None:
            p_img = f"{args.dir_dataset}/train/{p_img.split('/')[-1]}"
            assert os.path.exists(p_img), p_img
        list_inputs.append(p_img)
        list_merged.append(q)
    dataset.list_inputs = list_inputs
    dataset.update_labelled_queries(list_merged)
    dataloader = DataLoader(dataset, batch_size=1, num_workers=args.n_workers, shuffle=False)
    nth_query = len(list_prev)
    qs = QuerySelector(args, dataloader, device=device)
    dict_queries = qs(nth_query=nth_query, model=model, human_labels=True)
    os.makedirs(f"{args.dir_checkpoints}/{nth_query}_query", exist_ok=True)
    pkl.dump(dict_queries, open(f"{args.dir_checkpoints}/{nth_query}_query/queries.pkl", "wb"))
    print(f"Queries are saved at {args.dir_checkpoints}/{nth_query}_query/queries.pkl")
    return dict_queries


if __name__ == "__main__":
    main()
