"""CPU: the HOST logic of the drop-in QuerySelector (batching, RNG consumption, wire format, statistics, and the image
sharding over ranks) with the device kernels replaced by TEST-ONLY stand-ins built on the oracle.  The stand-ins live here,
not in the product: `pixelpick_b200` itself has no CPU path (tests/test_query_host.py checks that it refuses CPU tensors).

  * single process: `QuerySelector.__call__` == the reference's own output captured in tests/golden (call_*), and == the
    oracle for the top_n_percent = 0 / reverse_order / random variants, with identical NumPy RNG state afterwards;
  * gloo, world 2 and 3: image i is handled by rank i % world, every rank ends with the SAME dict, in dataloader order,
    equal to the single-process one; query_stats.pkl is written once (rank 0) and equals the single-process file.
"""
import os
import pickle
import socket
from argparse import Namespace

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import acq_oracle as orc

STRATS = ["entropy", "least_confidence", "margin_sampling"]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def make_args(strategy, n_classes, ignore_index, dir_root, top_n_percent=0.05, n_pixels_by_us=10, reverse_order=False,
              dataset_name="cs"):
    return Namespace(dataset_name=dataset_name, debug=False, dir_root=dir_root, experim_name="t", ignore_index=ignore_index,
                     mc_n_steps=20, n_classes=n_classes, n_pixels_by_us=n_pixels_by_us, network_name="deeplab",
                     query_strategy=strategy, reverse_order=reverse_order, stride_total=8, top_n_percent=top_n_percent,
                     use_mc_dropout=False, vote_type="soft")


class StubDataset:
    def __init__(self, logits, y, lab):
        self.logits, self.y = logits, y
        self.queries = [m.copy() for m in lab]
        self.labelled = None

    def label_queries(self, dict_queries, nth_query=None):
        self.labelled = (dict_queries, nth_query)


class StubLoader:
    def __init__(self, ds):
        self.dataset = ds

    def __iter__(self):
        for i in range(self.dataset.logits.shape[0]):
            d = {"x": self.dataset.logits[i:i + 1], "p_img": [f"img_{i:04d}.png"]}
            if self.dataset.y is not None:
                d["y"] = torch.from_numpy(self.dataset.y[i:i + 1])
            yield d


class StubModel:
    def eval(self):
        return self

    def __call__(self, x):
        return {"pred": x}


def install_standins(topk=orc.topk_indices_torch):
    """Replace the device entry points `QuerySelector` calls by oracle-based stand-ins (this process only)."""
    from pixelpick_b200 import _lib
    from pixelpick_b200 import query as q

    class Workspace:
        def __init__(self, *a):
            pass

        def prepare(self):
            pass

    def score_batch(self, model, x, h, w, labelled, void, keep, ws):
        pred = model(x)["pred"][:, :, :h, :w]
        b = pred.shape[0]
        uc = orc.uncertainty(orc.probabilities(pred), self.query_strategy).clone()
        fill = orc.FILL[self.query_strategy]
        uc[labelled] = fill
        if void is not None:
            uc[void] = fill
        if keep is not None:
            uc[~keep.view(b, h, w)] = fill
        return uc.reshape(b, h * w), ("full", pred)

    def select_pick(score, k, largest, pos, n=None, ws=None, hist0_valid=False):
        idx = torch.stack([torch.from_numpy(np.ascontiguousarray(topk(s, k, largest))).long() for s in score])
        return idx.gather(1, pos.long()) if pos is not None else idx[:, :n]

    def entropy_at(pred, sel):
        prob = F.softmax(pred, dim=1)
        ent = (-prob * torch.log(prob)).sum(dim=1)
        return ent.reshape(ent.shape[0], -1).gather(1, sel)

    _lib.TopKWorkspace = Workspace
    _lib.acq_select_pick = select_pick
    _lib.acq_entropy_at = entropy_at
    q.QuerySelector._score_batch = score_batch
    return q


def make_selector(q, args, loader, batch_imgs):
    qs = q.QuerySelector(args, loader, device=torch.device("cuda:0"), batch_imgs=batch_imgs)  # constructing touches no device
    qs.device = torch.device("cpu")                                                            # the stand-ins run on the host
    return qs


@pytest.fixture()
def standins(monkeypatch):
    from pixelpick_b200 import _lib
    from pixelpick_b200 import query as q
    for obj, name in ((_lib, "TopKWorkspace"), (_lib, "acq_select_pick"), (_lib, "acq_entropy_at"), (q.QuerySelector, "_score_batch")):
        monkeypatch.setattr(obj, name, getattr(obj, name))  # restored after the test
    return install_standins


@pytest.mark.parametrize("strat", STRATS)
@pytest.mark.parametrize("batch_imgs", [1, 2, 32])
def test_call_matches_reference_golden(golden, tmp_path, standins, capsys, strat, batch_imgs):
    q = standins()
    logits = torch.from_numpy(golden["call_logits"])
    ds = StubDataset(logits, golden["call_y"], golden["call_lab"])
    qs = make_selector(q, make_args(strat, 19, 19, str(tmp_path)), StubLoader(ds), batch_imgs)
    np.random.seed(0)
    d = qs(0, StubModel())
    capsys.readouterr()
    assert list(d) == [f"img_{i:04d}.png" for i in range(3)]  # dataloader order
    for i, (p, info) in enumerate(d.items()):
        assert info["height"] == 32 and info["width"] == 48
        assert np.array_equal(np.stack([info["x_coords"], info["y_coords"]]), golden[f"call_{strat}_{i}_xy"])
    assert ds.labelled[0] is d and ds.labelled[1] == 0
    stats = pickle.load(open(tmp_path / "checkpoints" / "t" / "0_query" / "query_stats.pkl", "rb"))
    assert sum(stats["label_distribution"].values()) == 30


@pytest.mark.parametrize("strat", STRATS + ["random"])
@pytest.mark.parametrize("kw", [dict(top_n_percent=0.0), dict(top_n_percent=0.05, reverse_order=True), dict(top_n_percent=0.05)])
def test_call_variants_match_oracle_and_consume_the_same_random_numbers(tmp_path, standins, capsys, strat, kw):
    q = standins()
    g = torch.Generator().manual_seed(17)
    logits = (torch.randn((5, 11, 24, 40), generator=g) * 3).float()
    rs = np.random.RandomState(17)
    y = rs.randint(0, 12, size=(5, 24, 40)).astype(np.int64)
    lab = rs.rand(5, 24, 40) < 0.01
    ds = StubDataset(logits, y, lab)
    qs = make_selector(q, make_args(strat, 11, 11, str(tmp_path), **kw), StubLoader(ds), 3)
    np.random.seed(5)
    torch.manual_seed(5)
    got = qs(1, StubModel())
    capsys.readouterr()
    state_np, state_t = np.random.get_state()[1].copy(), torch.get_rng_state().clone()
    np.random.seed(5)
    torch.manual_seed(5)
    if strat == "random":  # the oracle scores from logits; the random strategy draws its map instead (query.py:241-243)
        want = {}
        for i in range(5):
            uc = orc.apply_masks(torch.rand((1, 24, 40))[0], "random", lab[i], y[i] == 11)
            qm = orc.select_queries(uc, "random", 10, kw["top_n_percent"], kw.get("reverse_order", False))
            ys, xs = np.where(qm)
            want[f"img_{i:04d}.png"] = {"x_coords": xs, "y_coords": ys}
    else:
        want = orc.query_images([logits[i:i + 1] for i in range(5)], strat, lab, y == 11, [f"img_{i:04d}.png" for i in range(5)],
                                10, kw["top_n_percent"], kw.get("reverse_order", False))
    assert np.array_equal(state_np, np.random.get_state()[1]) and torch.equal(state_t, torch.get_rng_state())
    for p in want:
        assert np.array_equal(got[p]["x_coords"], want[p]["x_coords"]), p
        assert np.array_equal(got[p]["y_coords"], want[p]["y_coords"]), p


def test_statistics_accumulate_over_rounds_like_the_reference(golden, tmp_path, standins, capsys):
    """QueryStats is created once per selector and never reset (query.py:27,250-258): round 1's file covers rounds 0 and 1."""
    q = standins()
    logits = torch.from_numpy(golden["call_logits"])
    ds = StubDataset(logits, golden["call_y"], golden["call_lab"])
    qs = make_selector(q, make_args("entropy", 19, 19, str(tmp_path)), StubLoader(ds), 2)
    np.random.seed(0)
    qs(0, StubModel())
    qs(1, StubModel())
    capsys.readouterr()
    s1 = pickle.load(open(tmp_path / "checkpoints" / "t" / "1_query" / "query_stats.pkl", "rb"))
    assert sum(s1["label_distribution"].values()) == 60 and len(qs.query_stats.list_n_unique_labels) == 6


# ---- sharded over ranks (gloo) -------------------------------------------------------------------------------------
def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class MapDataset(torch.utils.data.Dataset):
    """map-style dataset with the reference interface (base_dataset.py:151-189 items, .queries, .label_queries) behind a real
    torch DataLoader: lets QuerySelector re-target the loader at this rank's images only; records what it served."""

    def __init__(self, logits, y, lab):
        self.logits, self.y = logits, y
        self.queries = [m.copy() for m in lab]
        self.labelled = None
        self.served = []

    def __len__(self):
        return self.logits.shape[0]

    def __getitem__(self, i):
        self.served.append(i)
        return {"x": self.logits[i], "y": torch.from_numpy(self.y[i]), "p_img": f"img_{i:04d}.png"}

    def label_queries(self, dict_queries, nth_query=None):
        self.labelled = (dict_queries, nth_query)


def _worker(rank, world, port, root, strat, kw, n_rounds, real_loader=False):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    q = install_standins()
    golden = np.load(os.path.join(ROOT, "tests", "golden", "query_golden.npz"))
    g = torch.Generator().manual_seed(23)
    logits = (torch.randn((7, 19, 32, 48), generator=g) * 3).float()
    logits[:3] = torch.from_numpy(golden["call_logits"])
    rs = np.random.RandomState(23)
    y = rs.randint(0, 20, size=(7, 32, 48)).astype(np.int64)
    y[:3] = golden["call_y"]
    lab = rs.rand(7, 32, 48) < 0.01
    lab[:3] = golden["call_lab"]
    ds = MapDataset(logits, y, lab) if real_loader else StubDataset(logits, y, lab)
    loader = torch.utils.data.DataLoader(ds, batch_size=1, shuffle=False, num_workers=0) if real_loader else StubLoader(ds)
    qs = make_selector(q, make_args(strat, 19, 19, os.path.join(root, f"w{world}"), **kw), loader, 2)
    np.random.seed(0)
    torch.manual_seed(0)
    out = []
    for r in range(n_rounds):
        if real_loader:
            ds.served = []
        d = qs(r, StubModel())
        out.append({"dict": d, "label_queries_nth": ds.labelled[1], "same_object": ds.labelled[0] is d,
                    "np_state": np.random.get_state()[1].copy(), "served": list(getattr(ds, "served", []))})
    pickle.dump(out, open(os.path.join(root, f"w{world}_r{rank}.pkl"), "wb"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("strat,kw", [("margin_sampling", dict(top_n_percent=0.05)), ("entropy", dict(top_n_percent=0.05, reverse_order=True)),
                                      ("random", dict(top_n_percent=0.0))])
def test_sharded_query_equals_single_process(tmp_path, strat, kw):
    import torch.multiprocessing as mp
    root = str(tmp_path)
    for world in (1, 2, 3):
        mp.spawn(_worker, args=(world, _free_port(), root, strat, kw, 2), nprocs=world, join=True)
    single = pickle.load(open(os.path.join(root, "w1_r0.pkl"), "rb"))
    names = [f"img_{i:04d}.png" for i in range(7)]
    if strat == "margin_sampling":  # the first three images are the reference's golden call
        golden = np.load(os.path.join(ROOT, "tests", "golden", "query_golden.npz"))
        for i in range(3):
            info = single[0]["dict"][names[i]]
            assert np.array_equal(np.stack([info["x_coords"], info["y_coords"]]), golden[f"call_{strat}_{i}_xy"])
    for world in (2, 3):
        for rank in range(world):
            got = pickle.load(open(os.path.join(root, f"w{world}_r{rank}.pkl"), "rb"))
            for r in range(2):
                assert list(got[r]["dict"]) == names                               # dataloader order on every rank
                for p in names:
                    for k in ("height", "width"):
                        assert got[r]["dict"][p][k] == single[r]["dict"][p][k]
                    for k in ("x_coords", "y_coords"):
                        assert np.array_equal(got[r]["dict"][p][k], single[r]["dict"][p][k]), (world, rank, r, p)
                assert np.array_equal(got[r]["np_state"], single[r]["np_state"])      # same RNG consumption on every rank
                assert got[r]["same_object"]
                assert got[r]["label_queries_nth"] == (r if rank == 0 else None)     # only rank 0 writes queries.pkl
        for r in range(2):
            f1 = os.path.join(root, "w1", "checkpoints", "t", f"{r}_query", "query_stats.pkl")
            fw = os.path.join(root, f"w{world}", "checkpoints", "t", f"{r}_query", "query_stats.pkl")
            a, b = pickle.load(open(f1, "rb")), pickle.load(open(fw, "rb"))
            assert a["label_distribution"] == b["label_distribution"]
            for k in ("avg_entropy", "avg_n_unique_labels", "avg_spatial_coverage"):
                # same lists in the same order: exact (the random strategy has no logits to take an entropy from: NaN)
                assert a[k] == b[k] or (np.isnan(a[k]) and np.isnan(b[k])), (world, r, k)


@pytest.mark.parametrize("strat,kw", [("margin_sampling", dict(top_n_percent=0.05)), ("entropy", dict(top_n_percent=0.0))])
def test_sharded_query_over_a_real_dataloader_reads_only_its_own_images(tmp_path, strat, kw):
    """multi-GPU fast path: with a map-style DataLoader every rank loads ONLY image i = rank (mod world), scores them, and
    the random ranks are drawn after one all-gather of the image sizes - same picks, same statistics file and the same NumPy
    stream state as the single-process walk over the whole loader."""
    import torch.multiprocessing as mp
    root = str(tmp_path)
    mp.spawn(_worker, args=(1, _free_port(), root, strat, kw, 2, False), nprocs=1, join=True)
    single = pickle.load(open(os.path.join(root, "w1_r0.pkl"), "rb"))
    names = [f"img_{i:04d}.png" for i in range(7)]
    for world in (2, 3):
        mp.spawn(_worker, args=(world, _free_port(), root, strat, kw, 2, True), nprocs=world, join=True)
        for rank in range(world):
            got = pickle.load(open(os.path.join(root, f"w{world}_r{rank}.pkl"), "rb"))
            for r in range(2):
                assert got[r]["served"] == list(range(rank, 7, world))               # 1/world of the dataset, nothing else
                assert list(got[r]["dict"]) == names
                for p in names:
                    for k in ("x_coords", "y_coords"):
                        assert np.array_equal(got[r]["dict"][p][k], single[r]["dict"][p][k]), (world, rank, r, p)
                assert np.array_equal(got[r]["np_state"], single[r]["np_state"])
        for r in range(2):
            f1 = os.path.join(root, "w1", "checkpoints", "t", f"{r}_query", "query_stats.pkl")
            fw = os.path.join(root, f"w{world}", "checkpoints", "t", f"{r}_query", "query_stats.pkl")
            a, b = pickle.load(open(f1, "rb")), pickle.load(open(fw, "rb"))
            assert a == b or all(a[k] == b[k] for k in a)


@pytest.mark.parametrize("strat", ["margin_sampling", "entropy"])
def test_call_with_human_labels_matches_reference_golden(golden, tmp_path, standins, capsys, strat):
    """query.py:144-222 with human_labels=True: the exclusion masks come from `dataset.list_labelled_queries` (label maps,
    ignore_index where unlabelled); no statistics file, no label_queries call; picks equal the reference's (wire_golden.pkl)."""
    q = standins()
    want = pickle.load(open(os.path.join(ROOT, "tests", "golden", "wire_golden.pkl"), "rb"))["human_call"][strat]
    logits, y, lab = torch.from_numpy(golden["call_logits"]), golden["call_y"], golden["call_lab"]

    class DS:
        list_labelled_queries = [np.where(lab[i], y[i], 19).astype(np.int64) for i in range(3)]
        called = False

        def label_queries(self, *a, **k):
            DS.called = True

    class Loader:
        dataset = DS()

        def __iter__(self):
            for i in range(3):
                yield {"x": logits[i:i + 1], "y": torch.from_numpy(y[i:i + 1]), "p_img": [f"img_{i:04d}.png"]}

    qs = make_selector(q, make_args(strat, 19, 19, str(tmp_path)), Loader(), 2)
    np.random.seed(0)
    d = qs(1, StubModel(), human_labels=True)
    capsys.readouterr()
    assert not DS.called and not (tmp_path / "checkpoints").exists()
    for i in range(3):
        info = d[f"img_{i:04d}.png"]
        assert np.array_equal(np.stack([info["x_coords"], info["y_coords"]]), want[i])
