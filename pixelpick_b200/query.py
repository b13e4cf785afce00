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

    def _flush(self, model, batch, human_labels, dict_queries, stats_on):
        xs = torch.cat([b["x"] for b in batch], dim=0).to(self.device, non_blocking=True)
        h, w = batch[0]["hw"]
        n = len(batch)
        hw = h * w
        k = self._k(h, w)
        st = self.query_strategy
        if self.dataset_name == "voc":  # query.py:171-174
            pad_h = ceil(h / self.stride_total) * self.stride_total - h
            pad_w = ceil(w / self.stride_total) * self.stride_total - w
            xs = F.pad(xs, pad=(0, pad_w, 0, pad_h), mode="reflect")
        lab = np.stack([b["mask"] for b in batch])
        labelled = torch.from_numpy((lab != self.ignore_index) if human_labels else lab.astype(bool)).to(self.device)
        void = None
        if batch[0]["y"] is not None:
            void = torch.from_numpy(np.stack([b["y"] == self.ignore_index for b in batch])).to(self.device)
        # drawn per image when the loop saw it (dataloader order, on every rank): query.py:40,64
        keep_np = None if batch[0]["keep"] is None else np.concatenate([b["keep"] for b in batch])
        pos_np = None if batch[0]["pos"] is None else np.concatenate([b["pos"] for b in batch])
        keep = None if keep_np is None else torch.from_numpy(keep_np).to(self.device)
        n_top = self.n_pixels_by_us if self.reverse_order else k
        largest = _lib.LARGEST[st]
        pos_t = None if pos_np is None else torch.from_numpy(pos_np)
        if st == "random":
            uc = torch.stack([b["rand"] for b in batch]).to(self.device)
            excl = labelled if void is None else (labelled | void)
            if keep is not None:
                excl = excl | ~keep.view(n, h, w)
            uc[excl] = _lib.FILL[st]
            sel = _lib.acq_select_pick(uc.view(n, hw), n_top, largest, pos_t, n=self.n_pixels_by_us)
            handle = None
        else:
            ws = self._workspace(n, hw, n_top)
            ws.prepare()
            score, handle = self._score_batch(model, xs, h, w, labelled, void, keep, ws)
            # only the n drawn ranks of the sorted top-k are needed (query.py:63-64): radix pick, no sort
            sel = _lib.acq_select_pick(score.view(n, hw), n_top, largest, pos_t, n=self.n_pixels_by_us, ws=ws, hist0_valid=True)
        sel, _ = torch.sort(sel.long(), dim=1)  # np.where order: row-major ascending (query.py:77)
        ent = None
        if stats_on and handle is not None:
            kind, t = handle
            ent = (_lib.acq_entropy_at_upsampled(t, (h, w), sel) if kind == "lowres"
                   else _lib.acq_entropy_at(t, sel)).cpu().numpy()
        sel_np = sel.cpu().numpy()
        n_new = 0
        for i, b in enumerate(batch):
            idx = sel_np[i]
            info = {"height": h, "width": w, "x_coords": idx % w, "y_coords": idx // w}
            dict_queries[b["p_img"]] = info
            n_new += idx.size
            if stats_on:
                e = ent[i] if ent is not None else np.full(idx.size, np.nan)
                self.query_stats.update_selected(idx, w, b["y"], e)
        return n_new

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
        y = None
        batch: List[dict] = []
        world, rank = ppdist.world(), ppdist.rank()
        order: List[str] = []      # every image path in dataloader order (all ranks walk the whole loader)
        mine: List[int] = []       # dataloader positions of the images this rank scores
        self.query_stats.begin_round()
        with torch.no_grad():
            for batch_ind, dict_data in enumerate(self.dataloader):
                x = dict_data["x"]
                y = dict_data.get("y", None)
                h, w = tuple(x.shape[2:])
                # the random numbers of EVERY image are drawn here, in dataloader order, on every rank (query.py:40,64 and
                # UncertaintySampler._random): the streams, hence the picks, do not depend on the world size
                keep, pos = self._draw_positions(1, h, w)
                rand = self.uncertainty_sampler(torch.empty(1, 1, h, w))[0] if self.query_strategy == "random" else None
                order.append(dict_data["p_img"][0])
                n_imgs += 1
                if batch_ind % world != rank:
                    continue
                if y is not None:
                    y = y.squeeze(dim=0).numpy()
                item = {"x": x, "y": y, "mask": np.asarray(prev_queries[batch_ind]), "hw": (h, w), "p_img": order[-1],
                        "keep": keep, "pos": pos, "rand": rand}
                if batch and (item["hw"] != batch[0]["hw"] or len(batch) == self.batch_imgs):
                    n_pixels += self._flush(model, batch, human_labels, dict_queries, not human_labels and batch[0]["y"] is not None)
                    batch = []
                batch.append(item)
                mine.append(batch_ind)
            if batch:
                n_pixels += self._flush(model, batch, human_labels, dict_queries, not human_labels and batch[0]["y"] is not None)
        assert n_imgs > 0, "no queries are chosen!"
        stats_on = not human_labels and y is not None
        if world > 1:  # the round's ONE exchange: per-rank picks (and statistics) -> every rank, back in dataloader order
            gathered = ppdist.all_gather_objects((dict_queries, self.query_stats.round_payload(mine) if stats_on else None))
            merged: dict = dict()
            for d, _ in gathered:
                merged.update(d)
            dict_queries = {p: merged[p] for p in order}
            n_pixels = sum(len(info["x_coords"]) for info in dict_queries.values())
            if stats_on:
                self.query_stats.absorb_round([payload for _, payload in gathered])
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
