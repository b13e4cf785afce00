"""ORACLE (test infrastructure, NOT product code) — CPU restatement of PixelPick's query path.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may
import this module; the product path (`pixelpick_b200/`) never does.

Each function restates one piece of the reference `query.py` (NoelShin/PixelPick @ 43c2981) with the
same torch-CPU / NumPy primitives the reference itself executes (the reference has no native code:
its arithmetic *is* torch's CPU kernels and NumPy's legacy RandomState).

Pinning: PINNED.  The reference ships no tests or golden vectors of its own, so this oracle is pinned against
outputs of the UNMODIFIED reference imported and run in the build container: `tests/golden/make_golden.py`
(generator, committed) -> `tests/golden/query_golden.npz` -> `tests/test_oracle_golden.py` (checker, bit-exact).
"""
from typing import Dict, Optional

import numpy as np
import torch
import torch.nn.functional as F

# fill value of excluded pixels and direction of the top-k (query.py:45-61,195-201)
FILL = {"entropy": 0.0, "least_confidence": 0.0, "margin_sampling": 1.0, "random": 1.0}
LARGEST = {"entropy": True, "least_confidence": True, "margin_sampling": False, "random": False}


def probabilities(logits: torch.Tensor) -> torch.Tensor:
    """query.py:190 — softmax over the class axis of [b, c, h, w] logits (fp32)."""
    return F.softmax(logits, dim=1)


def uncertainty(prob: torch.Tensor, strategy: str) -> torch.Tensor:
    """query.py:224-247 UncertaintySampler: [b, c, h, w] -> [b, h, w]."""
    if strategy == "entropy":  # query.py:229-230 (NaN where a probability underflows to 0)
        return (-prob * torch.log(prob)).sum(dim=1)
    if strategy == "least_confidence":  # query.py:233-234
        return 1.0 - prob.max(dim=1)[0]
    if strategy == "margin_sampling":  # query.py:237-239 (best-vs-second-best)
        top2 = prob.topk(k=2, dim=1).values
        return (top2[:, 0] - top2[:, 1]).abs()
    if strategy == "random":  # query.py:242-244 (torch CPU generator)
        b, _, h, w = prob.shape
        return torch.rand((b, h, w))
    raise ValueError(strategy)


def apply_masks(uc_map: torch.Tensor, strategy: str, labelled: Optional[np.ndarray],
                void: Optional[np.ndarray]) -> torch.Tensor:
    """query.py:195-201 — already-labelled and void pixels get the 'never pick me' value."""
    uc_map = uc_map.clone()
    if labelled is not None:
        uc_map[torch.from_numpy(np.asarray(labelled, dtype=bool))] = FILL[strategy]
    if void is not None:
        uc_map[torch.from_numpy(np.asarray(void, dtype=bool))] = FILL[strategy]
    return uc_map


def score_map(logits: torch.Tensor, strategy: str, labelled=None, void=None) -> torch.Tensor:
    """query.py:190-201 for one image: logits [1, c, h, w] -> masked uncertainty map [h, w]."""
    uc = uncertainty(probabilities(logits), strategy).squeeze(dim=0)
    return apply_masks(uc, strategy, labelled, void)


def topk_indices_torch(uc_flat: torch.Tensor, k: int, largest: bool) -> np.ndarray:
    """query.py:57-61 — exactly what the reference runs (tie order is whatever torch CPU does)."""
    return uc_flat.topk(k=k, dim=0, largest=largest).indices.cpu().numpy()


def ord_key(scores: np.ndarray, largest: bool) -> np.ndarray:
    """uint32 key whose ascending order is the selection order: NaN ranks as the largest value
    (torch.topk semantics), -0.0 == +0.0.  Mirrors pp::ord_key in csrc/pp_common.cuh."""
    s = np.asarray(scores, dtype=np.float32) + np.float32(0.0)
    u = s.view(np.uint32).copy()
    neg = (u & np.uint32(0x80000000)) != 0
    u = np.where(neg, ~u, u | np.uint32(0x80000000)).astype(np.uint32)
    u[np.isnan(s)] = np.uint32(0xFFFFFFFF)
    return (~u).astype(np.uint32) if largest else u


def topk_indices_spec(uc_flat, k: int, largest: bool) -> np.ndarray:
    """The order contract of the CUDA path: sorted by score, ties -> lower flat index.
    Equals `topk_indices_torch` whenever the top k+1 scores are pairwise distinct."""
    s = uc_flat.detach().cpu().numpy() if isinstance(uc_flat, torch.Tensor) else np.asarray(uc_flat)
    key = ord_key(s.reshape(-1), largest).astype(np.uint64)
    comp = (key << np.uint64(32)) | np.arange(key.size, dtype=np.uint64)
    part = np.argpartition(comp, k - 1)[:k] if k < comp.size else np.arange(comp.size)
    return part[np.argsort(comp[part], kind="stable")].astype(np.int64)


def select_queries(uc_map: torch.Tensor, strategy: str, n_pixels_by_us: int, top_n_percent: float,
                   reverse_order: bool = False, topk=topk_indices_torch) -> np.ndarray:
    """query.py:33-69 QuerySelector._select_queries: [h, w] scores -> bool [h, w] mask of new queries.
    Consumes the GLOBAL NumPy RNG exactly like the reference (np.random.choice, legacy RandomState)."""
    h, w = uc_map.shape[-2:]
    uc_flat = uc_map.flatten().clone()
    k = int(h * w * top_n_percent) if top_n_percent > 0.0 else n_pixels_by_us
    largest = LARGEST[strategy]
    if reverse_order:  # query.py:38-54: random 5 % subsample first, then the n most uncertain of it
        assert top_n_percent > 0.0
        ind = np.random.choice(range(h * w), k, False)
        sampling = np.zeros(h * w, dtype=bool)
        sampling[ind] = True
        uc_flat[torch.from_numpy(~sampling)] = FILL[strategy]
        ind_queries = topk(uc_flat, n_pixels_by_us, largest)
    else:
        ind_queries = topk(uc_flat, k, largest)
        if top_n_percent > 0.0:  # query.py:63-64
            ind_queries = np.random.choice(ind_queries, n_pixels_by_us, False)
    query = np.zeros(h * w, dtype=bool)
    query[ind_queries] = True
    return query.reshape(h, w)


def encode_query(p_img: str, size, query: np.ndarray) -> Dict[str, dict]:
    """query.py:72-87 — wire format of one image's queries."""
    y_coords, x_coords = np.where(query)
    return {p_img: {"height": size[0], "width": size[1], "x_coords": x_coords, "y_coords": y_coords}}


def query_images(logits_per_image, strategy: str, labelled_masks, void_masks, p_imgs, n_pixels_by_us: int = 10,
                 top_n_percent: float = 0.05, reverse_order: bool = False, topk=topk_indices_torch):
    """query.py:159-212 loop body over pre-computed logits (one [1, c, h, w] tensor per image)."""
    out: Dict[str, dict] = {}
    for i, logits in enumerate(logits_per_image):
        h, w = logits.shape[-2:]
        uc = score_map(logits, strategy, None if labelled_masks is None else labelled_masks[i],
                       None if void_masks is None else void_masks[i])
        q = select_queries(uc, strategy, n_pixels_by_us, top_n_percent, reverse_order, topk)
        out.update(encode_query(p_imgs[i], (h, w), q))
    return out


def entropy_at(logits: torch.Tensor, query: np.ndarray) -> list:
    """query.py:260-264 QueryStats._get_entropy: entropy of the selected pixels of one image."""
    prob = probabilities(logits)
    ent = (-prob * torch.log(prob)).sum(dim=1).cpu().numpy()
    return ent.flatten()[query.flatten()].tolist()
