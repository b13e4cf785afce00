"""CPU: host logic around the train step (args, optimizer groups, LR schedules, running metrics) against goldens produced by
the UNMODIFIED reference - tests/golden/host_golden.pkl, written by tests/golden/make_golden_host.py."""
import os
import pickle
import warnings
from argparse import Namespace

import numpy as np
import pytest
import torch

from pixelpick_b200.args import Arguments
from pixelpick_b200.utils import AverageMeter, RunningScore, get_lr_scheduler, get_optimizer, optimizer_kind

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def g():
    return pickle.load(open(os.path.join(HERE, "golden", "host_golden.pkl"), "rb"))


ARG_CASES = {
    "cs": ["--dataset_name", "cs"],
    "cv": ["--dataset_name", "cv"],
    "voc": ["--dataset_name", "voc"],
    "cs_entropy_rev": ["--dataset_name", "cs", "-qs", "entropy", "--reverse_order", "--seed", "3", "--suffix", "x"],
    "cv_fully_sup": ["--dataset_name", "cv", "--n_pixels_by_us", "0", "--debug"],
    "cv_top0_mc": ["--dataset_name", "cv", "--top_n_percent", "0", "--use_mc_dropout", "--vote_type", "hard"],
}
OURS_ONLY = {"synthetic", "cuda_graph", "gpu_augment"}  # switches this framework adds (documented in README.md)


@pytest.mark.parametrize("case", sorted(ARG_CASES))
def test_arguments_match_the_reference(g, case, tmp_path, monkeypatch, capsys):
    """args.py:10-205: every field the reference's parse_args produces, same name and value (incl. experim_name and the
    checkpoint directory layout); the reference also exports CUDA_VISIBLE_DEVICES from --gpu_ids, which a one-process-per-GPU
    launcher must not do, so that side effect is not mirrored."""
    monkeypatch.chdir(tmp_path)
    ours = vars(Arguments().parse_args(argv=["--dir_root", "root"] + ARG_CASES[case]))
    capsys.readouterr()
    want = g["args"][case]
    assert set(want) - set(ours) == set()
    assert set(ours) - set(want) == OURS_ONLY
    for k, v in want.items():
        assert ours[k] == v, (k, ours[k], v)
    assert os.path.exists(tmp_path / "root" / "checkpoints" / want["experim_name"] / "args.txt")


def _stub_model():
    m = torch.nn.Module()
    m.backbone = torch.nn.Conv2d(3, 4, 1)
    m.aspp = torch.nn.Conv2d(4, 4, 1)
    m.low_level_conv = torch.nn.Conv2d(4, 2, 1)
    m.seg_head = torch.nn.Conv2d(6, 3, 1)
    return m


@pytest.mark.parametrize("case", ["cs_Adam", "cv_Adam", "cv_SGD", "voc_SGD"])
def test_optimizer_groups_match_the_reference(g, case):
    """utils/utils.py:112-306: class and per-group lr / weight decay / momentum / betas / eps - including that VOC's declared
    weight decay (1e-4) and Adam's declared eps (1e-7) never reach the optimizer."""
    ds, kind = case.split("_")
    ns = Namespace(**g["args"][ds])
    ns.optimizer_type = kind
    want = g["optim"][case]
    opt = get_optimizer(ns, _stub_model())
    assert type(opt).__name__ == want["cls"] == optimizer_kind(ns)
    assert [len(pg["params"]) for pg in opt.param_groups] == want["n_params"]
    for got, ref in zip(opt.param_groups, want["groups"]):
        for k in ("lr", "weight_decay", "momentum", "dampening", "nesterov", "betas", "eps", "amsgrad", "maximize"):
            if k in ref:
                assert got[k] == ref[k], (k, got[k], ref[k])


def test_dataset_overrides_the_declared_optimizer_type(g):
    """`cs` builds Adam and `voc` builds SGD whatever optimizer_type says (utils.py:114,208)."""
    ns = Namespace(**g["args"]["cs"])
    ns.optimizer_type = "SGD"
    assert type(get_optimizer(ns, _stub_model())).__name__ == "Adam"
    ns = Namespace(**g["args"]["voc"])
    ns.optimizer_type, ns.lr_scheduler_type = "Adam", "MultiStepLR"
    opt = get_optimizer(ns, _stub_model())
    assert type(opt).__name__ == "SGD"
    assert type(get_lr_scheduler(ns, opt, iters_per_epoch=5)).__name__ == "Poly"  # utils.py:323-325


@pytest.mark.parametrize("kind,key", [("Poly", "poly"), ("MultiStepLR", "multistep")])
def test_lr_schedule_matches_the_reference(g, kind, key):
    """The learning rates of every iteration when the scheduler is driven as model.py:138-145 drives it."""
    m = _stub_model()
    opt = torch.optim.SGD([{"params": m.backbone.parameters(), "lr": 1e-3}, {"params": m.aspp.parameters(), "lr": 1e-2}])
    sched = get_lr_scheduler(Namespace(dataset_name="cs", lr_scheduler_type=kind, n_epochs=3), opt, iters_per_epoch=5)
    assert type(sched).__name__ == g[key]["cls"]
    lrs = []
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for epoch in range(1, 4):
            for _ in range(5):
                lrs.append([pg["lr"] for pg in opt.param_groups])
                opt.step()
                if kind == "Poly":
                    sched.step(epoch=epoch - 1)
            if kind == "MultiStepLR":
                sched.step(epoch=epoch - 1)
    lrs.append([pg["lr"] for pg in opt.param_groups])
    assert lrs == g[key]["lrs"]  # same float operations in the same order: exact


def test_running_score_matches_the_reference(g):
    s = g["score"]
    rs_full, rs_pairs, rs_conf = RunningScore(11), RunningScore(11), RunningScore(11)
    for lt, lp in s["batches"]:
        rs_full.update(lt, lp)
        keep = lt < 11                                   # what the train loop hands over: labelled pixels only
        rs_pairs.update_pairs(lt[keep], lp[keep])
        conf = np.zeros((11, 11))
        np.add.at(conf, (lt[keep], lp[keep]), 1)         # what the device accumulator hands over
        rs_conf.update_confusion(conf)
    for rs in (rs_full, rs_pairs, rs_conf):
        assert np.array_equal(rs.confusion_matrix, s["confusion"])
        scores, cls_iu = rs.get_scores()
        assert list(scores) == list(s["scores"])
        for k in scores:
            assert scores[k] == s["scores"][k], k
        assert all(cls_iu[c] == s["cls_iu"][c] for c in range(11))
    rs_full.reset()
    assert rs_full.confusion_matrix.sum() == 0


def test_average_meter_matches_the_reference(g):
    m = AverageMeter()
    for (v, n), want in zip([(0.5, 1), (1.25, 4), (3.0, 2)], g["meter"]):
        m.update(v, n)
        for k, w in want.items():
            assert float(getattr(m, k)) == w, k


def test_write_log_bytes_match_the_reference(g, tmp_path):
    from pixelpick_b200.utils import write_log
    fp = str(tmp_path / "log.txt")
    write_log(fp, header=["epoch", "mIoU", "pixel_acc", "loss"])
    write_log(fp, list_entities=[1, np.float64(0.25), 0.5, 1.75])
    write_log(fp, list_entities=[2, float("nan"), np.float32(0.5), "x"])
    assert open(fp, "rb").read() == g["log"]["rows"]
    write_log(fp, list_entities=[7, 8], header=["a", "b"])
    assert open(fp, "rb").read() == g["log"]["header_and_row"]


def test_evaluate_matches_the_reference(g, tmp_path, capsys):
    """eval.py:15-94 on the same stub model / loader the reference was run on: identical mIoU and log_val.txt bytes, although the
    images are micro-batched here (the stub has no fused path, so this is the plain `model(x)["pred"]` route, on the CPU)."""
    from pixelpick_b200.eval import evaluate

    class Stub(torch.nn.Module):
        def forward(self, x):
            return {"pred": torch.stack([x[:, 0], -x[:, 0], x[:, 1], x[:, 2], x.sum(1) * 0.3], dim=1)}

    class DS:
        n_classes, dataset_name = 5, "cs"

    class Loader(list):
        dataset = DS()

    gen = torch.Generator().manual_seed(0)
    sizes = [(20, 28)] * 5 + [(17, 23)] * 2 + [(20, 28)]
    items = [{"x": torch.randn((1, 3) + s, generator=gen), "y": torch.randint(0, 6, (1,) + s, generator=gen)} for s in sizes]
    miou = evaluate(Stub(), Loader(items), "stub", epoch=3, dir_ckpt=str(tmp_path), device=torch.device("cpu"), batch_imgs=4)
    capsys.readouterr()
    assert float(miou) == g["evaluate"]["miou"]
    assert open(tmp_path / "e03" / "val" / "log_val.txt", "rb").read() == g["evaluate"]["log"]


def test_train_epoch_trajectory_matches_the_reference(g, monkeypatch, capsys):
    """train.py:14-103 for three epochs: the reference ran dense `F.cross_entropy(model(x)["pred"], y*, ignore_index)`; ours runs
    the sparse labelled-pixel loss (a TEST-ONLY torch stand-in for pp_sparse_ce here) on `forward_lowres`.  Same data, same
    initial weights, our get_optimizer / Poly: the running loss after each epoch, the final parameters and the final learning
    rates agree (float32 round-off of a different but equivalent evaluation order)."""
    import torch.nn.functional as F
    from pixelpick_b200 import train as T
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

    gen = torch.Generator().manual_seed(42)
    batches = []
    for _ in range(3):
        q = torch.rand((4, 32, 64), generator=gen) < 0.01
        batches.append({"x": torch.randn((4, 3, 32, 64), generator=gen), "y": torch.randint(0, 20, (4, 32, 64), generator=gen),
                        "queries": q.to(torch.uint8)})

    class DS:
        ignore_index, n_classes = 19, 19

    class Loader(list):
        dataset = DS()

    monkeypatch.setattr(T, "sparse_cross_entropy", standin_ce)
    torch.manual_seed(0)
    model = Tiny(19)
    ns = Namespace(**g["args"]["cs"])
    ns.n_epochs = 3
    loader = Loader(batches)
    opt = get_optimizer(ns, model)
    sched = get_lr_scheduler(ns, optimizer=opt, iters_per_epoch=len(loader))
    tracker, avg = AverageMeter(), []
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for e in range(1, 4):
            model, opt, sched = T.train_epoch(e, loader, model, opt, sched, tracker, "golden", device=torch.device("cpu"))
            avg.append(float(tracker.avg))
    capsys.readouterr()
    want = g["train"]
    assert np.allclose(avg, want["avg_loss"], rtol=1e-5, atol=0)
    got = torch.cat([p.detach().flatten() for p in model.parameters()]).numpy()
    assert np.allclose(got, want["params"], rtol=1e-3, atol=2e-5)
    assert [pg["lr"] for pg in opt.param_groups] == want["lrs"]


@pytest.mark.parametrize("ds", ["cs", "cv"])
def test_model_train_epoch_log_matches_the_reference(g, ds, tmp_path, monkeypatch, capsys):
    """model.py:93-159 for three epochs on a bare Model object (eager step, TEST-ONLY torch stand-in for pp_sparse_ce):
    log_train.txt - epoch, running mIoU / pixel accuracy of the labelled pixels, running loss, meters reset per epoch - and the
    final parameters / learning rates equal the reference's, for the per-iteration Poly schedule (cs) and the per-epoch
    MultiStepLR (cv)."""
    import torch.nn.functional as F
    from pixelpick_b200 import model as M
    from pixelpick_b200.loss import labelled_pixel_list
    from pixelpick_b200.utils import write_log

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

    monkeypatch.setattr(M, "sparse_cross_entropy", standin_ce)
    ns = Namespace(**g["args"][ds])
    ns.n_epochs = 3
    nc = ns.n_classes
    gen = torch.Generator().manual_seed(7)
    batches = []
    for _ in range(3):
        q = torch.rand((4, 32, 64), generator=gen) < 0.01
        batches.append({"x": torch.randn((4, 3, 32, 64), generator=gen), "y": torch.randint(0, nc + 1, (4, 32, 64), generator=gen),
                        "queries": q.to(torch.uint8)})

    class DS:
        n_pixels_total = 0

    class Loader(list):
        dataset = DS()

    torch.manual_seed(0)
    net = Tiny(nc)
    m = object.__new__(M.Model)
    m.n_pixels_by_us, m.nth_query, m.dir_checkpoints, m.experim_name = 10, 0, str(tmp_path), "golden"
    m.device, m.ignore_index, m.debug, m.lr_scheduler_type = torch.device("cpu"), ns.ignore_index, False, ns.lr_scheduler_type
    m.running_loss, m.running_score = AverageMeter(), RunningScore(nc)
    m._use_graph, m._graph, m._graph_labels, m._graph_shape = False, None, None, None
    m.dataloader = Loader(batches)
    m.log_train = str(tmp_path / "log_train.txt")
    write_log(m.log_train, header=["epoch", "mIoU", "pixel_acc", "loss"])
    opt = get_optimizer(ns, net)
    sched = get_lr_scheduler(ns, optimizer=opt, iters_per_epoch=3)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for e in range(1, 4):
            net, opt, sched = m._train_epoch(e, net, opt, sched)
    capsys.readouterr()
    want = g["model_epoch"][ds]
    got_rows = open(m.log_train).read().splitlines()
    want_rows = want["log"].decode().splitlines()
    assert got_rows[0] == want_rows[0] and len(got_rows) == len(want_rows) == 4
    for a, b in zip(got_rows[1:], want_rows[1:]):
        a, b = a.split(","), b.split(",")
        assert a[0] == b[0]
        assert float(a[1]) == pytest.approx(float(b[1]), rel=1e-12) and float(a[2]) == pytest.approx(float(b[2]), rel=1e-12)
        assert float(a[3]) == pytest.approx(float(b[3]), rel=1e-5)
    got = torch.cat([p.detach().flatten() for p in net.parameters()]).numpy()
    assert np.allclose(got, want["params"], rtol=1e-3, atol=2e-5)
    assert [pg["lr"] for pg in opt.param_groups] == want["lrs"]


def test_model_val_log_and_best_checkpoint_match_the_reference(g, tmp_path, capsys):
    """model.py:177-239 on a bare Model object over four epochs with stub models of varying quality: log_val.txt bytes and the
    epochs at which best_miou_model.pt is rewritten (only on a strict improvement) equal the reference's."""
    from pixelpick_b200 import model as M
    from pixelpick_b200.utils import write_log

    class ValStub(torch.nn.Module):
        def __init__(self, noise):
            super().__init__()
            self.noise = torch.nn.Parameter(torch.tensor(float(noise)), requires_grad=False)

        def forward(self, x):
            return {"pred": x + self.noise * torch.roll(x, 1, dims=1) * 2.0}

    gen = torch.Generator().manual_seed(3)
    items = []
    for _ in range(4):
        y = torch.randint(0, 6, (1, 12, 16), generator=gen)
        x = torch.nn.functional.one_hot(y.clamp(max=4), 5).permute(0, 3, 1, 2).float() + 0.1 * torch.randn((1, 5, 12, 16), generator=gen)
        items.append({"x": x, "y": y})

    class Loader(list):
        dataset = None

    m = object.__new__(M.Model)
    m.n_pixels_by_us, m.nth_query, m.dir_checkpoints, m.experim_name = 10, 2, str(tmp_path), "golden"
    m.device, m.dataset_name, m.stride_total, m.debug, m.best_miou, m.n_classes = torch.device("cpu"), "cs", 8, False, -1.0, 5
    m.running_loss, m.running_score = AverageMeter(), RunningScore(5)
    m.dataloader_val = Loader(items)
    os.makedirs(tmp_path / "2_query")
    m.log_val = str(tmp_path / "2_query" / "log_val.txt")
    write_log(m.log_val, header=["epoch", "mIoU", "pixel_acc"])
    ck, saved = tmp_path / "2_query" / "best_miou_model.pt", []
    for e, noise in ((1, 0.5), (2, 1.0), (3, 0.47), (4, 0.49)):
        if ck.exists():
            ck.unlink()
        m._val(e, ValStub(noise))
        if ck.exists():
            saved.append(e)
    capsys.readouterr()
    want = g["model_val"]
    assert open(m.log_val, "rb").read() == want["log"]
    assert saved == want["saved_at"] and float(m.best_miou) == want["best"]
