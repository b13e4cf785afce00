"""Generate golden vectors for the Q path by running the UNMODIFIED reference in this container.

    python tests/golden/make_golden.py            # needs /root/reference (build container only)

The reference (NoelShin/PixelPick @ 43c2981) is imported from /root/reference; nothing is copied.
Outputs `tests/golden/query_golden.npz` (small) which travels with the repo; the GPU box has no
/root/reference, so tests only ever read the .npz.

What is pinned
  scores_*      UncertaintySampler._entropy/_least_confidence/_margin_sampling on softmax(logits)
  uc_*          + mask fills of QuerySelector.__call__ (query.py:195-201)
  sel_*         QuerySelector._select_queries under np.random.seed (top-5 % + np.random.choice,
                top_n_percent == 0, reverse_order)
  call_*        QuerySelector.__call__ end to end with a stub dataloader/model (x carries the logits)
  big_*         a 256x512, C=19 image regenerated from a seed: only the selected coordinates are stored
"""
import os
import sys
import tempfile
from argparse import Namespace

import numpy as np
import torch

REF = "/root/reference"
sys.dont_write_bytecode = True
sys.path.insert(0, REF)
import query as refq  # noqa: E402  (the reference module)

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "query_golden.npz")
STRATS = ["entropy", "least_confidence", "margin_sampling"]


def make_args(strategy, n_classes, ignore_index, top_n_percent=0.05, n_pixels_by_us=10, reverse_order=False,
              dir_root="/tmp"):
    return Namespace(dataset_name="cs", debug=False, dir_root=dir_root, experim_name="golden",
                     ignore_index=ignore_index, mc_n_steps=20, n_classes=n_classes, n_pixels_by_us=n_pixels_by_us,
                     network_name="deeplab", query_strategy=strategy, reverse_order=reverse_order, stride_total=8,
                     top_n_percent=top_n_percent, use_mc_dropout=False, vote_type="soft")


def make_inputs(seed, n, C, h, w, scale=3.0, void_frac=0.02, n_lab=10):
    g = torch.Generator().manual_seed(seed)
    logits = (torch.randn((n, C, h, w), generator=g) * scale).float()
    rs = np.random.RandomState(seed)
    y = rs.randint(0, C, size=(n, h, w)).astype(np.int64)
    y[rs.rand(n, h, w) < void_frac] = C  # ignore_index == n_classes (cs / cv convention)
    lab = np.zeros((n, h * w), dtype=bool)
    for i in range(n):
        lab[i, rs.choice(h * w, n_lab, replace=False)] = True
    return logits, y, lab.reshape(n, h, w)


class StubDataset:
    def __init__(self, logits, y, lab):
        self.logits, self.y = logits, y
        self.queries = [m.copy() for m in lab]
        self.labelled = None

    def label_queries(self, dict_queries, nth_query=None):
        self.labelled = (dict_queries, nth_query)


class StubLoader:
    """Yields what `for batch_ind, dict_data in enumerate(dataloader)` sees with batch_size=1."""

    def __init__(self, ds):
        self.dataset = ds

    def __iter__(self):
        for i in range(self.dataset.logits.shape[0]):
            yield {"x": self.dataset.logits[i:i + 1], "y": torch.from_numpy(self.dataset.y[i:i + 1]),
                   "p_img": [f"img_{i:04d}.png"]}


class StubModel:
    def eval(self):
        return self

    def __call__(self, x):
        return {"pred": x}


def main():
    out = {}
    # ---- scores + masks on three class counts --------------------------------------------------
    for C, (h, w) in [(11, (24, 40)), (19, (16, 32)), (21, (20, 28))]:
        logits, y, lab = make_inputs(100 + C, 2, C, h, w)
        out[f"logits_c{C}"] = logits.numpy()
        out[f"y_c{C}"] = y
        out[f"lab_c{C}"] = lab
        prob = torch.softmax(logits, dim=1)
        for s in STRATS:
            uc = refq.UncertaintySampler(s)(prob)
            out[f"scores_{s}_c{C}"] = uc.numpy()
            fill = 0.0 if s in ["entropy", "least_confidence"] else 1.0
            ucm = uc.clone()
            for i in range(2):
                ucm[i][torch.from_numpy(lab[i])] = fill
                ucm[i][torch.from_numpy(y[i] == C)] = fill
            out[f"uc_{s}_c{C}"] = ucm.numpy()
            # _select_queries, three flavours
            for tag, kw in [("top5", dict(top_n_percent=0.05)), ("topn", dict(top_n_percent=0.0)),
                            ("rev", dict(top_n_percent=0.05, reverse_order=True))]:
                qs = refq.QuerySelector(make_args(s, C, C, **kw), None, device=torch.device("cpu"))
                np.random.seed(7)
                sel = np.stack([qs._select_queries(ucm[i].clone()) for i in range(2)])
                out[f"sel_{tag}_{s}_c{C}"] = sel
    # ---- entropy NaN semantics ----------------------------------------------------------------
    logits, y, lab = make_inputs(5, 1, 11, 8, 16)
    logits[0, 3, 2, 5] = -200.0  # p underflows to 0 -> 0 * log 0 = NaN
    logits[0, 0, 7, 1] = 150.0
    out["logits_nan"] = logits.numpy()
    uc = refq.UncertaintySampler("entropy")(torch.softmax(logits, dim=1))
    out["scores_entropy_nan"] = uc.numpy()
    qs = refq.QuerySelector(make_args("entropy", 11, 11, top_n_percent=0.0, n_pixels_by_us=4), None,
                            device=torch.device("cpu"))
    out["sel_nan"] = qs._select_queries(uc[0].clone())
    # ---- QuerySelector.__call__ end to end ------------------------------------------------------
    for s in STRATS:
        logits, y, lab = make_inputs(300, 3, 19, 32, 48)
        ds = StubDataset(logits, y, lab)
        with tempfile.TemporaryDirectory() as td:
            qs = refq.QuerySelector(make_args(s, 19, 19, dir_root=td), StubLoader(ds), device=torch.device("cpu"))
            np.random.seed(0)
            d = qs(0, StubModel())
        for i, (p, info) in enumerate(sorted(d.items())):
            out[f"call_{s}_{i}_xy"] = np.stack([info["x_coords"], info["y_coords"]])
        if s == "entropy":
            out["call_logits"], out["call_y"], out["call_lab"] = logits.numpy(), y, lab
    # ---- a full-size Cityscapes-shape image, inputs regenerated from the seed in the test -------
    # logits from NumPy's legacy RandomState (bit-identical on every platform; torch.randn's CPU stream depends on the
    # vectorised code path of the host, which made this case skip on some GPU boxes)
    _, y, lab = make_inputs(900, 1, 19, 256, 512)
    logits = torch.from_numpy((np.random.RandomState(900).standard_normal((1, 19, 256, 512)) * 3.0).astype(np.float32))
    # exact, order-independent checksum: sums of the raw bit patterns (a float sum depends on the host's vector width)
    bits = logits.numpy().view(np.uint32).astype(np.uint64)
    out["big_logits_checksum"] = np.array([bits.sum(), (bits * (np.arange(bits.size, dtype=np.uint64).reshape(bits.shape) % 65521)).sum()],
                                          dtype=np.uint64)
    for s in STRATS:
        qs = refq.QuerySelector(make_args(s, 19, 19), None, device=torch.device("cpu"))
        uc = refq.UncertaintySampler(s)(torch.softmax(logits, dim=1))[0]
        fill = 0.0 if s in ["entropy", "least_confidence"] else 1.0
        uc[torch.from_numpy(lab[0])] = fill
        uc[torch.from_numpy(y[0] == 19)] = fill
        largest = s in ["entropy", "least_confidence"]
        top = uc.flatten().topk(k=int(256 * 512 * 0.05), largest=largest)
        out[f"big_topk_idx_{s}"] = top.indices.numpy().astype(np.int32)
        out[f"big_topk_val_{s}"] = top.values.numpy()
        np.random.seed(3)
        out[f"big_sel_{s}"] = np.flatnonzero(qs._select_queries(uc.clone())).astype(np.int32)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
