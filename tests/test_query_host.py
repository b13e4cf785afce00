"""CPU: host-side logic of pixelpick_b200/query.py that needs no GPU (wire format, merge, RNG identity)."""
import os
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


def test_select_ordering_maps_are_consistent_on_the_host():
    """The radix select is exact iff (a) pp_host_ord_key orders scores like torch.topk (descending for `largest` with NaN
    first, ascending otherwise with NaN last, -0.0 == +0.0) and (b) the level-0 bucket is monotone non-decreasing in that
    key.  Both maps are host functions of the library: checked here without a GPU on specials, dense sweeps of the score
    ranges and random bit patterns."""
    import ctypes
    import numpy as np
    from pixelpick_b200 import _lib
    lib = _lib.lib()
    rs = np.random.RandomState(0)
    specials = np.array([0.0, -0.0, 1.0, -1.0, np.inf, -np.inf, np.nan, 1e-45, -1e-45, 1e-38, 3.4e38, -3.4e38, 0.5, 0.999999,
                         2.9444, 4.0, 4.0000005, 3.9999998, 0.0019550342, 0.001955], dtype=np.float32)
    dense = np.concatenate([np.linspace(0, 1, 4001), np.linspace(0, 3.05, 4001), np.linspace(-0.01, 4.2, 3001)]).astype(np.float32)
    rand_bits = rs.randint(0, 2 ** 32, size=4000, dtype=np.uint64).astype(np.uint32).view(np.float32)
    xs = np.concatenate([specials, dense, rand_bits])
    for largest in (0, 1):
        key = np.array([lib.pp_host_ord_key(ctypes.c_float(float(x)), largest) for x in xs], dtype=np.uint64)
        bkt = np.array([lib.pp_host_bucket0(ctypes.c_float(float(x)), largest) for x in xs], dtype=np.int64)
        assert bkt.min() >= 0 and bkt.max() <= 2047
        # (a) key order == the selection order of torch.topk / sort with NaN as the largest value
        nan = np.isnan(xs)
        assert len(set(key[nan])) == 1 and (key[nan][0] == (0 if largest else 2 ** 32 - 1))
        fin = ~nan
        order = np.argsort(key[fin], kind="stable")
        v = xs[fin][order].astype(np.float64)
        assert np.all(np.diff(v) <= 0) if largest else np.all(np.diff(v) >= 0)
        z = key[(xs == 0)]
        assert len(set(z)) == 1  # -0.0 and +0.0 share a key
        # (b) monotone: sort by key, buckets must be non-decreasing
        o = np.argsort(key, kind="stable")
        assert np.all(np.diff(bkt[o]) >= 0), "bucket0 is not monotone in the ordering key"
        # equal keys -> equal buckets
        for k_ in np.unique(key):
            assert len(set(bkt[key == k_])) == 1
    # resolution where the strategies live: largest-first scores in [0.5, 1) must spread over many buckets
    b_lo, b_hi = lib.pp_host_bucket0(ctypes.c_float(0.5), 1), lib.pp_host_bucket0(ctypes.c_float(0.999), 1)
    assert abs(int(b_lo) - int(b_hi)) > 200


def test_eval_loader_batching_and_voc_padding_on_the_host_path(tmp_path):
    """eval.confusion_over_loader with a model that has no fused path (plain `model(x)["pred"]`): same-sized images are
    micro-batched, size changes flush, the VOC branch reflect-pads to a stride multiple and crops (eval.py:49-55); counts
    equal RunningScore.update image by image."""
    import numpy as np
    import torch
    from pixelpick_b200.eval import confusion_over_loader, evaluate
    from pixelpick_b200.utils import RunningScore

    class Stub(torch.nn.Module):
        calls = []

        def forward(self, x):
            Stub.calls.append(tuple(x.shape))
            # 5 "classes" from simple functions of the input, defined for any (padded) size
            return {"pred": torch.stack([x[:, 0], -x[:, 0], x[:, 1], x[:, 2], x.sum(1) * 0.3], dim=1)}

    g = torch.Generator().manual_seed(0)
    sizes = [(20, 28)] * 5 + [(17, 23)] * 2 + [(20, 28)]
    items = [{"x": torch.randn((1, 3) + s, generator=g), "y": torch.randint(0, 6, (1,) + s, generator=g)} for s in sizes]

    class DS:
        n_classes, dataset_name = 5, "voc"

    class Loader(list):
        dataset = DS()

    for name in ("cs", "voc"):
        Stub.calls = []
        conf = confusion_over_loader(Stub(), Loader(items), 5, torch.device("cpu"), dataset_name=name, stride_total=8, batch_imgs=4)
        ref = RunningScore(5)
        m = Stub()
        for it in items:
            ref.update(it["y"].numpy(), m(it["x"])["pred"].argmax(1).numpy())
        assert np.array_equal(conf, ref.confusion_matrix)
        batches = [c for c in Stub.calls if c[0] > 1 or True][:4]
        assert [c[0] for c in Stub.calls[:4]] == [4, 1, 2, 1]  # 4 + 1 of the first size, 2 of the second, 1 again
        if name == "voc":
            assert Stub.calls[0][2:] == (24, 32) and Stub.calls[2][2:] == (24, 24)  # padded to multiples of 8
    DS.dataset_name = "cs"
    miou = evaluate(Stub(), Loader(items), "stub", epoch=3, dir_ckpt=str(tmp_path), device=torch.device("cpu"), batch_imgs=4)
    assert 0.0 <= miou <= 1.0 and (tmp_path / "e03" / "val" / "log_val.txt").read_text().startswith("epoch,miou,pixel_acc")


def test_train_cli_detects_human_label_files(tmp_path):
    """train.py:199-203: earlier `*/queries.pkl` files that carry `category_id` switch the run to human labels; query files
    written by the selector itself (no category_id) do not."""
    import pickle
    import numpy as np
    from pixelpick_b200.query import gather_previous_query_files, merge_previous_query_files
    from pixelpick_b200.train import _has_human_labels
    own = {"a.png": {"height": 4, "width": 5, "x_coords": np.array([1, 2]), "y_coords": np.array([0, 3])}}
    human = {"a.png": dict(own["a.png"], category_id=np.array([2, 7]))}
    for d, q in (("0_query", own), ("1_query", human)):
        (tmp_path / d).mkdir()
        pickle.dump(q, open(tmp_path / d / "queries.pkl", "wb"))
    files = sorted(gather_previous_query_files(str(tmp_path)))
    assert [_has_human_labels(f) for f in files] == [False, True]
    merged = merge_previous_query_files([files[1]], ignore_index=19, verbose=False)
    assert merged["a.png"].shape == (4, 5) and merged["a.png"][0, 1] == 2 and merged["a.png"][3, 2] == 7
    assert (merged["a.png"] != 19).sum() == 2


# ---- wire format pinned to the reference: tests/golden/wire_golden.pkl from make_golden_wire.py ----
def _wire_golden():
    import pickle
    return pickle.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "wire_golden.pkl"), "rb"))


def _same_tree(a, b):
    assert type(a) is type(b), (type(a), type(b))
    if isinstance(a, dict):
        assert list(a.keys()) == list(b.keys())          # insertion order is part of the pickle image
        for k in a:
            _same_tree(a[k], b[k])
    elif isinstance(a, (list, tuple)):
        assert len(a) == len(b)
        for x, y in zip(a, b):
            _same_tree(x, y)
    elif isinstance(a, np.ndarray):
        assert a.dtype == b.dtype and a.shape == b.shape and np.array_equal(a, b)
    else:
        assert a == b


def test_wire_format_encode_is_byte_compatible_with_the_reference():
    """queries.pkl written through our encode_query is the reference's file byte for byte (query.py:72-88, model.py / main_al
    dump it with pickle) - so via/ tooling and train.py keep reading it."""
    import pickle
    from pixelpick_b200.query import QuerySelector
    g = _wire_golden()
    enc = {}
    for p, m in zip(g["paths"], g["masks"]):
        enc.update(QuerySelector.encode_query(p, m.shape, m))
    _same_tree(enc, g["encoded"])
    assert pickle.dumps(enc, protocol=4) == g["encoded_bytes"]


def test_wire_format_decode_matches_the_reference():
    from pixelpick_b200.query import QuerySelector
    g = _wire_golden()
    _same_tree(QuerySelector.decode_queries(g["encoded"]), g["decoded_list"])
    _same_tree(QuerySelector.decode_queries(g["encoded"], return_as_dict=True), g["decoded_dict"])
    _same_tree(QuerySelector.decode_queries({g["paths"][0]: g["encoded"][g["paths"][0]]}), g["decoded_one"])
    _same_tree(QuerySelector.decode_queries(g["human"], ignore_index=255, return_as_dict=True), g["human_255"])
    _same_tree(QuerySelector.decode_queries(g["human"], ignore_index=19), g["human_19"])


def test_merge_previous_query_files_matches_the_reference(tmp_path):
    import pickle
    from pixelpick_b200.query import gather_previous_query_files, merge_previous_query_files
    g = _wire_golden()
    files = []
    for r, d in enumerate(g["rounds"]):
        os.makedirs(tmp_path / f"{r}_query")
        f = str(tmp_path / f"{r}_query" / "queries.pkl")
        pickle.dump(d, open(f, "wb"))
        files.append(f)
    assert sorted(gather_previous_query_files(str(tmp_path))) == sorted(files)
    _same_tree(merge_previous_query_files(files, ignore_index=255, verbose=False), g["merged"])


def test_query_stats_matches_the_reference(tmp_path, capsys):
    """QueryStats (query.py:250-308): the reference recomputes a full entropy map per image and masks it; ours is fed the picks
    and the entropy AT the picks (pp_acq_entropy_at on the device).  Same inputs -> the same query_stats.pkl."""
    import pickle
    g = _wire_golden()["stats"]
    qs = q.QueryStats(Namespace(dir_root=str(tmp_path), experim_name="t", n_classes=11))
    for qm, y, logits in zip(g["queries"], g["y"], g["logits"]):
        idx = np.flatnonzero(qm)                                    # row-major ascending == np.where order
        prob = torch.softmax(torch.from_numpy(logits), dim=1)
        ent = (-prob * torch.log(prob)).sum(dim=1)[0].reshape(-1)[torch.from_numpy(idx)]
        qs.update_selected(idx, qm.shape[1], y, ent.numpy())
    qs.save(0)
    capsys.readouterr()
    got = pickle.load(open(tmp_path / "checkpoints" / "t" / "0_query" / "query_stats.pkl", "rb"))
    want = g["saved"]
    assert list(got.keys()) == list(want.keys())
    assert got["label_distribution"] == want["label_distribution"]
    assert got["avg_n_unique_labels"] == want["avg_n_unique_labels"]
    assert got["avg_spatial_coverage"] == pytest.approx(want["avg_spatial_coverage"], rel=1e-12)
    assert got["avg_entropy"] == pytest.approx(want["avg_entropy"], rel=1e-6)
    assert np.allclose(qs.list_entropy, g["list_entropy"], rtol=1e-6, atol=0)


def test_train_mirror_host_plumbing_with_standin_kernels(tmp_path, monkeypatch, capsys):
    """train.py:14-176 host logic of pixelpick_b200.train (epoch loop, scheduler stepping, logs, evaluation interval, best-model
    file) on the CPU, with the model and the sparse-CE kernel replaced by TEST-ONLY torch stand-ins (the product has no CPU path)."""
    import torch.nn.functional as F
    from pixelpick_b200 import train as T
    from pixelpick_b200.args import Arguments
    from pixelpick_b200.loss import labelled_pixel_list

    class Tiny(torch.nn.Module):
        def __init__(self, n_classes):
            super().__init__()
            self.backbone = torch.nn.Conv2d(3, 8, 3, stride=4, padding=1)
            self.aspp, self.low_level_conv = torch.nn.Conv2d(8, 8, 1), torch.nn.Conv2d(8, 8, 1)
            self.seg_head = torch.nn.Conv2d(8, n_classes, 1)

        def forward_lowres(self, x):
            return self.seg_head(self.low_level_conv(self.aspp(torch.relu(self.backbone(x)))))

    def standin_ce(lowres, y, queries, ignore_index, size=None, return_pred=False, px=None, n_valid=None):
        px = labelled_pixel_list(y, queries, ignore_index)
        up = F.interpolate(lowres, size=tuple(y.shape[-2:]), mode="bilinear", align_corners=True)
        at = up.permute(0, 2, 3, 1).reshape(up.shape[0], -1, up.shape[1])[px[0].long(), px[1].long()]
        loss = F.cross_entropy(at, px[2].long())
        return (loss, at.argmax(1).to(torch.int32), px) if return_pred else loss

    evals = []

    def standin_evaluate(model, dataloader, experim_name, epoch=None, dir_ckpt=None, **kw):
        evals.append((epoch, len(dataloader), kw["stride_total"]))
        return 0.1 * epoch

    monkeypatch.setattr(T, "get_model", lambda args: Tiny(args.n_classes))
    monkeypatch.setattr(T, "sparse_cross_entropy", standin_ce)
    monkeypatch.setattr(T, "evaluate", standin_evaluate)
    args = Arguments().parse_args(argv=["--dataset_name", "cs", "--dir_root", str(tmp_path), "--n_workers", "0", "--synthetic", "6", "32", "64",
                                        "--n_epochs", "4"])
    loader = T.get_dataloader(args, args.batch_size, 0, True)
    ck = str(tmp_path / "ck")
    model = T.train(args, loader, eval_interval=2, dir_ckpt=ck, device=torch.device("cpu"))
    capsys.readouterr()
    assert isinstance(model, Tiny)
    assert [e[0] for e in evals] == [2, 4] and all(e[2] == 8 for e in evals)      # every eval_interval epochs, on the val loader
    assert os.path.exists(os.path.join(ck, "best_model.pt"))
    assert sorted(torch.load(os.path.join(ck, "best_model.pt"))["model"]) == sorted(model.state_dict())
    for e in range(1, 5):                                                        # train.py:96-101: one log per epoch directory
        rows = open(os.path.join(ck, f"e{e:02d}", "log_train.txt")).read().strip().splitlines()
        assert rows[-1].split(",")[0] == str(e) and len(rows[-1].split(",")) == 4
    assert open(os.path.join(ck, "e01", "log_train.txt")).read().startswith("epoch,miou,pixel_acc,loss")


def test_labelled_pixel_lists_host_and_device_forms_agree():
    """loss.labelled_pixel_list (torch, used by the eager step) and labelled_pixel_list_host (NumPy, feeds the captured graph's
    fixed-capacity staging buffers) build the same (image, flat index, label) list: y != ignore_index and queries != 0
    (model.py:108-110 + F.cross_entropy's ignore_index), row-major; the host form pads to `capacity` and refuses to overflow."""
    from pixelpick_b200.loss import labelled_pixel_list, labelled_pixel_list_host
    rs = np.random.RandomState(4)
    y = torch.from_numpy(rs.randint(0, 12, size=(3, 9, 14)).astype(np.int64))       # 11 == ignore_index
    qm = torch.from_numpy((rs.rand(3, 9, 14) < 0.2).astype(np.uint8))
    for queries in (qm, None):
        a = labelled_pixel_list(y, queries, 11)
        b = labelled_pixel_list_host(y, queries, 11)
        n = int(b[3])
        assert n == a[0].numel() and all(t.dtype == torch.int32 for t in a + b[:3])
        for u, v in zip(a, b[:3]):
            assert torch.equal(u, v)
        want = ((y != 11) & (queries.bool() if queries is not None else True)).reshape(3, -1)
        assert n == int(want.sum())
        flat = a[0].long() * want.shape[1] + a[1].long()
        assert torch.equal(flat, torch.sort(flat).values) and bool(want.reshape(-1)[flat].all())
        assert torch.equal(a[2].long(), y.reshape(-1)[flat])
        padded = labelled_pixel_list_host(y, queries, 11, capacity=n + 7)
        assert all(t.numel() == n + 7 for t in padded[:3]) and int(padded[3]) == n
        assert all(torch.equal(t[:n], u) and not t[n:].any() for t, u in zip(padded[:3], a))
        with pytest.raises(_lib.PixelPickError):
            labelled_pixel_list_host(y, queries, 11, capacity=n - 1)


def test_session_host_buffer_contract_is_checked_before_the_library_copies():
    """pp_acq_session_* memcpy fixed-size blocks from / to the caller's host buffers; the Python layer rejects any buffer whose
    size, dtype or placement does not match before the call (no GPU needed for the check itself)."""
    chk = _lib.session_check_host_buffers
    n, C, H, W, k, ns = 3, 19, 8, 16, 6, 4
    logits = torch.zeros((n, C, H, W))
    m8, mb = torch.zeros((n, H, W), dtype=torch.uint8), torch.zeros((n, H * W), dtype=torch.bool)
    pos = torch.tensor([[0, 1, 2, 5]] * n, dtype=torch.int32)
    sel, topk = torch.zeros((n, ns), dtype=torch.int32), torch.zeros((n, k), dtype=torch.int32)
    chk("t", n, C, H, W, k, ns, logits, (m8, mb), pos, sel, topk)                  # the accepted forms (flat mask views included)
    chk("t", n, C, H, W, k, ns, logits, (None, None), None, sel)
    bad = [dict(logits=logits.double()), dict(logits=logits[:, :18].contiguous()), dict(logits=logits.permute(0, 1, 3, 2)),
           dict(masks=(m8.int(), None)), dict(masks=(m8[:, :4].contiguous(), None)), dict(pos=pos.long()),
           dict(pos=torch.tensor([[0, 1, 2, 6]] * n, dtype=torch.int32)), dict(pos=-pos - 1), dict(sel=sel[:2]),
           dict(sel=sel.long()), dict(topk=topk[:, :5].contiguous())]
    for kw in bad:
        with pytest.raises(_lib.PixelPickError):
            chk("t", n, C, H, W, k, ns, **{**dict(logits=logits, masks=(m8, mb), pos=pos, sel=sel, topk=topk), **kw})
    with pytest.raises(_lib.PixelPickError):
        chk("t", 0, C, H, W, k, ns, logits[:0])


def test_score_output_and_workspace_sizes_are_checked():
    class WS:
        n_img, HW = 4, 32 * 48
    ok = torch.empty((4, 32, 48), dtype=torch.float32, device="meta")  # a meta tensor: sizes without memory or a device
    with pytest.raises(_lib.PixelPickError):   # not a CUDA tensor
        _lib._check_score_outputs("t", 4, 32, 48, ok, None)
    with pytest.raises(_lib.PixelPickError):   # workspace sized for fewer images
        _lib._check_score_outputs("t", 5, 32, 48, None, WS)
    with pytest.raises(_lib.PixelPickError):   # or for another image size
        _lib._check_score_outputs("t", 4, 32, 64, None, WS)
    _lib._check_score_outputs("t", 4, 32, 48, None, WS)
    _lib._check_score_outputs("t", 3, 32, 48, None, WS)


def test_host_uncertainty_sampler_has_no_cpu_path_except_random(golden):
    """UncertaintySampler (query.py:225-247) scores through the device kernel, so CPU probability maps are refused; only the
    `random` strategy is host work, as in the reference (torch CPU generator, query.py:242-244)."""
    prob = torch.softmax(torch.from_numpy(golden["logits_c19"]), dim=1)
    for strat in ("entropy", "least_confidence", "margin_sampling"):
        with pytest.raises(_lib.PixelPickError):
            q.UncertaintySampler(strat)(prob)
    torch.manual_seed(3)
    r = q.UncertaintySampler("random")(prob)
    torch.manual_seed(3)
    assert r.shape == prob.shape[:1] + prob.shape[2:] and torch.equal(r, torch.rand(r.shape))
