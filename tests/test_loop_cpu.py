"""CPU: the HOST logic of the active-learning loop (pixelpick_b200.model.Model, mirror of model.py:17-175) - round structure,
files written, label merging, scheduler stepping - with the network and the device kernels replaced by TEST-ONLY torch
stand-ins.  The product itself has no CPU path (Model refuses to start without a CUDA device; that refusal is tested too)."""
import os
import pickle

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from pixelpick_b200.args import Arguments
from test_query_selector_cpu import install_standins  # tests/ is on sys.path (conftest.py lives there)


class Tiny(torch.nn.Module):
    def __init__(self, n_classes):
        super().__init__()
        self.backbone = torch.nn.Conv2d(3, 8, 3, stride=4, padding=1)
        self.aspp, self.low_level_conv = torch.nn.Conv2d(8, 8, 1), torch.nn.Conv2d(8, 8, 1)
        self.seg_head = torch.nn.Conv2d(8, n_classes, 1)

    def forward_lowres(self, x):
        return self.seg_head(self.low_level_conv(self.aspp(torch.relu(self.backbone(x)))))

    def forward(self, x):
        return {"pred": F.interpolate(self.forward_lowres(x), size=x.shape[2:], mode="bilinear", align_corners=True)}


def _standin_ce(lowres, y, queries, ignore_index, size=None, return_pred=False, px=None, n_valid=None):
    from pixelpick_b200.loss import labelled_pixel_list
    px = labelled_pixel_list(y, queries, ignore_index)
    up = F.interpolate(lowres, size=tuple(y.shape[-2:]), mode="bilinear", align_corners=True)
    at = up.permute(0, 2, 3, 1).reshape(up.shape[0], -1, up.shape[1])[px[0].long(), px[1].long()]
    loss = F.cross_entropy(at, px[2].long())
    return (loss, at.argmax(1).to(torch.int32), px) if return_pred else loss


def _standin_eval_confusion(lowres, size, labels, conf):
    pred = F.interpolate(lowres, size=tuple(size), mode="bilinear", align_corners=True).argmax(1)
    n = conf.shape[0]
    keep = (labels >= 0) & (labels < n)
    conf += torch.bincount(n * labels[keep] + pred[keep], minlength=n * n).reshape(n, n)


def test_model_refuses_to_start_without_a_cuda_device(tmp_path, capsys):
    from pixelpick_b200.model import Model
    args = Arguments().parse_args(argv=["--dataset_name", "cv", "--dir_root", str(tmp_path), "--synthetic", "4", "32", "64"])
    capsys.readouterr()
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            Model(args)


@pytest.mark.parametrize("dataset,extra", [("cv", []), ("cs", ["-qs", "entropy"]), ("cv", ["--n_pixels_by_us", "0"])])
def test_active_learning_rounds_on_standin_kernels(tmp_path, monkeypatch, capsys, dataset, extra):
    from pixelpick_b200 import _lib
    from pixelpick_b200 import model as M
    from pixelpick_b200 import query as Q
    for obj, name in ((_lib, "TopKWorkspace"), (_lib, "acq_select_pick"), (_lib, "acq_entropy_at"), (Q.QuerySelector, "_score_batch")):
        monkeypatch.setattr(obj, name, getattr(obj, name))  # restored after the test
    install_standins()
    monkeypatch.setattr(_lib, "eval_confusion_upsampled", _standin_eval_confusion)
    monkeypatch.setattr(M, "sparse_cross_entropy", _standin_ce)
    monkeypatch.setattr(M, "get_model", lambda args: Tiny(args.n_classes))
    argv = ["--dataset_name", dataset, "--dir_root", str(tmp_path), "--n_workers", "0", "--synthetic", "6", "32", "64",
            "--n_epochs", "2", "--max_budget", "30", "--no_cuda_graph"] + extra
    args = Arguments().parse_args(argv=argv)
    with monkeypatch.context() as ctor:                                # only to get past the constructor's refusal
        ctor.setattr(torch.cuda, "is_available", lambda: True)
        ctor.setattr(torch.cuda, "current_device", lambda: 0)
        m = M.Model(args)
    m.device = m.query_selector.device = torch.device("cpu")           # the stand-ins run on the host
    fully_sup = args.n_pixels_by_us == 0
    n0 = m.dataloader.dataset.n_pixels_total
    m()
    capsys.readouterr()
    ck = os.path.join(str(tmp_path), "checkpoints", args.experim_name)
    if fully_sup:                                                      # model.py:55-64
        d = os.path.join(ck, "fully_sup")
        assert open(os.path.join(d, "log_train.txt")).read().splitlines()[0] == "epoch,mIoU,pixel_acc,loss"
        assert len(open(os.path.join(d, "log_val.txt")).read().splitlines()) == 1 + 2
        assert os.path.exists(os.path.join(d, "best_miou_model.pt"))
        return
    n_stages = 30 // 10                                                # model.py:67-69
    for r in range(n_stages):
        d = os.path.join(ck, f"{r}_query")
        rows = open(os.path.join(d, "log_train.txt")).read().splitlines()
        assert rows[0] == "epoch,mIoU,pixel_acc,loss" and [x.split(",")[0] for x in rows[1:]] == ["1", "2"]
        assert len(open(os.path.join(d, "log_val.txt")).read().splitlines()) == 3
        assert sorted(torch.load(os.path.join(d, "best_miou_model.pt"))["model"]) == sorted(Tiny(args.n_classes).state_dict())
        q = pickle.load(open(os.path.join(d, "queries.pkl"), "rb"))   # written by the selector through the QUERY dataset
        assert len(q) == 6 and all(len(v["x_coords"]) == 10 for v in q.values())
        assert os.path.exists(os.path.join(d, "query_stats.pkl"))
    # model.py:84: the round's picks are merged into the TRAIN dataset and persisted once more under {nth_query + 1}_query
    assert os.path.exists(os.path.join(ck, f"{n_stages}_query", "queries.pkl"))
    assert m.dataloader.dataset.n_pixels_total == n0 + 6 * 10 * n_stages
    assert m.dataloader_query.dataset.n_pixels_total == n0 + 6 * 10 * n_stages
    for a, b in zip(m.dataloader.dataset.queries, m.dataloader_query.dataset.queries):
        assert np.array_equal(a, b)                                    # both views of the labelled set stay in step


# ---- the same loop, one process per "GPU" (gloo, world 2): sharded train loader + gradient all-reduce + sharded query ----
def _loop_worker(rank, world, port, root):
    import torch.distributed as dist
    from unittest import mock
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), PP_CUDNN_BENCHMARK="0")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pixelpick_b200 import _lib
    from pixelpick_b200 import model as M
    install_standins()
    _lib.eval_confusion_upsampled = _standin_eval_confusion
    M.sparse_cross_entropy = _standin_ce
    M.get_model = lambda args: Tiny(args.n_classes)
    argv = ["--dataset_name", "cv", "--dir_root", root, "--n_workers", "0", "--synthetic", "8", "32", "64", "--n_epochs", "2",
            "--max_budget", "20", "--no_cuda_graph"]
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        args = Arguments().parse_args(argv=argv)
        with mock.patch.object(torch.cuda, "is_available", lambda: True), mock.patch.object(torch.cuda, "current_device", lambda: 0):
            m = M.Model(args)
        m.device = m.query_selector.device = torch.device("cpu")
        seen = []
        orig = m.train_step

        def spy(model, optimizer, dict_data, reducer=None):
            seen.append(sorted(dict_data["p_img"]))
            out = orig(model, optimizer, dict_data, reducer)
            spy.model = model
            return out

        m.train_step = spy
        m()
    params = torch.cat([p.detach().flatten() for p in spy.model.parameters()])
    pickle.dump({"params": params, "seen": seen, "queries": m.dataloader.dataset.queries, "experim": args.experim_name},
                open(os.path.join(root, f"loop_r{rank}.pkl"), "wb"))
    dist.barrier()
    dist.destroy_process_group()


def test_active_learning_rounds_sharded_over_two_ranks(tmp_path):
    import socket
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    root = str(tmp_path)
    mp.spawn(_loop_worker, args=(2, port, root), nprocs=2, join=True)
    r0, r1 = (pickle.load(open(os.path.join(root, f"loop_r{r}.pkl"), "rb")) for r in (0, 1))
    assert torch.equal(r0["params"], r1["params"])                        # broadcast init + all-reduced gradients: replicas agree
    assert len(r0["seen"]) == len(r1["seen"]) > 0
    for a, b in zip(r0["seen"], r1["seen"]):
        assert not set(a) & set(b)                                         # the ranks train on disjoint images every step
    for a, b in zip(r0["queries"], r1["queries"]):
        assert np.array_equal(a, b)                                        # and end with the same labelled set
    ck = os.path.join(root, "checkpoints", r0["experim"])
    for r in range(2):
        q = pickle.load(open(os.path.join(ck, f"{r}_query", "queries.pkl"), "rb"))
        assert len(q) == 8 and all(len(v["x_coords"]) == 10 for v in q.values())
        assert os.path.exists(os.path.join(ck, f"{r}_query", "best_miou_model.pt"))
