"""CPU: host-side logic of pixelpick_b200/query.py that needs no GPU (wire format, merge, RNG identity)."""
import pickle
from argparse import Namespace

import numpy as np
import pytest
import torch

from pixelpick_b200 import _lib
from pixelpick_b200 import query as q


def test_choice_equals_permutation_prefix_including_rng_state():
    for k, n in [(6553, 10), (8640, 10), (38, 10), (10, 10)]:
        a = np.random.RandomState(1).permutation(1 << 20)[:k]
        np.random.seed(5)
        r1 = np.random.choice(a, n, False)
        s1 = np.random.get_state()
        np.random.seed(5)
        r2 = a[np.random.permutation(k)[:n]]
        s2 = np.random.get_state()
        assert np.array_equal(r1, r2) and np.array_equal(s1[1], s2[1]) and s1[2] == s2[2]


def test_encode_decode_roundtrip():
    rs = np.random.RandomState(0)
    masks = {f"im_{i:03d}.png": rs.rand(12, 20) < 0.05 for i in range(4)}
    enc = {}
    for p, m in masks.items():
        enc.update(q.QuerySelector.encode_query(p, m.shape, m))
    dec = q.QuerySelector.decode_queries(enc)
    for got, (_, m) in zip(dec, sorted(masks.items())):
        assert got.dtype == bool and np.array_equal(got, m)
    dec = q.QuerySelector.decode_queries(enc, return_as_dict=True)
    assert list(dec) == sorted(masks)
    one = q.QuerySelector.decode_queries({"a": enc["im_000.png"]})
    assert isinstance(one, list) and len(one) == 1
    with pytest.raises(ValueError):
        q.QuerySelector.decode_queries({})


def test_decode_with_category_ids_and_merge(tmp_path):
    info = {"height": 4, "width": 5, "x_coords": np.array([1, 2]), "y_coords": np.array([0, 3]), "category_id": [7, 2]}
    m = q.QuerySelector.decode_queries({"a.png": info}, ignore_index=255)[0]
    assert m.dtype == np.int64 and m[0, 1] == 7 and m[3, 2] == 2 and (m == 255).sum() == 18
    info2 = {"height": 4, "width": 5, "x_coords": np.array([1]), "y_coords": np.array([0]), "category_id": [9]}
    for i, d in enumerate([{"a.png": info}, {"a.png": info2, "b.png": info}]):
        (tmp_path / f"{i}_query").mkdir()
        pickle.dump(d, open(tmp_path / f"{i}_query" / "queries.pkl", "wb"))
    files = sorted(q.gather_previous_query_files(str(tmp_path)))
    assert len(files) == 2
    merged = q.merge_previous_query_files(files, ignore_index=255, verbose=False)
    assert merged["a.png"][0, 1] == 9 and merged["a.png"][3, 2] == 2 and merged["b.png"][0, 1] == 7


def test_query_selector_refuses_cpu_device():
    args = Namespace(dataset_name="cs", debug=False, dir_root="/tmp", experim_name="t", ignore_index=19,
                     mc_n_steps=20, n_classes=19, n_pixels_by_us=10, network_name="deeplab",
                     query_strategy="entropy", reverse_order=False, stride_total=8, top_n_percent=0.05,
                     use_mc_dropout=False, vote_type="soft")
    with pytest.raises(_lib.PixelPickError):
        q.QuerySelector(args, None, device=torch.device("cpu"))


def test_kernels_refuse_cpu_tensors():
    with pytest.raises(_lib.PixelPickError):
        _lib.acq_score(torch.zeros(1, 19, 8, 8), "entropy")
    with pytest.raises(_lib.PixelPickError):
        _lib.sparse_ce(torch.zeros(1, 19, 8, 8), (32, 32), torch.zeros(1, dtype=torch.int32),
                       torch.zeros(1, dtype=torch.int32), torch.zeros(1, dtype=torch.int32))


def test_spatial_coverage_matches_reference_formula():
    ys, xs = np.array([0, 3, 7]), np.array([1, 5, 2])
    got = q.QueryStats._spatial_coverage(ys, xs)
    pts = np.stack([ys, xs], 1).astype(float)
    d = [np.linalg.norm(pts[i] - pts[j]) for i in range(3) for j in range(3) if i != j]
    assert abs(got - np.mean(d)) < 1e-12
