"""GPU: the active-learning round loop (model.py mirror) end to end on a synthetic dataset, and its CLI surface."""
import os
import pickle

import numpy as np
import pytest
import torch

from pixelpick_b200.args import Arguments
from pixelpick_b200.model import Model

pytestmark = pytest.mark.gpu


def test_two_rounds_on_synthetic_data(tmp_path):
    args = Arguments().parse_args(argv=["--dataset_name", "cs", "--dir_root", str(tmp_path), "--n_workers", "0",
                                        "--max_budget", "20", "--n_epochs", "1", "--synthetic", "8", "64", "128"])
    assert args.batch_size == 4 and args.n_classes == 19 and args.lr_scheduler_type == "Poly"
    torch.backends.cudnn.benchmark = False  # the reference turns it on (args.py:197); autotuning ~50 conv shapes x 3 batch sizes costs minutes here
    m = Model(args)
    before = m.dataloader_query.dataset.n_pixels_total
    m()
    ck = tmp_path / "checkpoints" / args.experim_name
    for r in (0, 1):
        assert (ck / f"{r}_query" / "log_train.txt").exists() and (ck / f"{r}_query" / "log_val.txt").exists()
        assert (ck / f"{r}_query" / "best_miou_model.pt").exists()
        assert (ck / f"{r}_query" / "query_stats.pkl").exists()
    q = pickle.load(open(ck / "1_query" / "queries.pkl", "rb"))
    assert len(q) == 8 and all(len(v["x_coords"]) == 10 for v in q.values())
    # the query dataset's own labels grew by 10 px / image / round and never re-pick a labelled pixel
    assert m.dataloader_query.dataset.n_pixels_total == before + 2 * 8 * 10
    sd = torch.load(ck / "1_query" / "best_miou_model.pt")["model"]
    assert "seg_head.classifier.weight" in sd and "backbone.features.0.0.weight" in sd
    rows = open(ck / "1_query" / "log_train.txt").read().strip().splitlines()
    assert rows[0].startswith("epoch") and len(rows) == 2 and np.isfinite(float(rows[1].split(",")[3]))


def test_loss_decreases_on_a_fixed_batch(tmp_path):
    args = Arguments().parse_args(argv=["--dataset_name", "cs", "--dir_root", str(tmp_path), "--n_workers", "0",
                                        "--synthetic", "4", "64", "128"])
    from pixelpick_b200.utils import get_model, get_optimizer
    m = Model(args)
    model = get_model(args).to(m.device)
    opt = get_optimizer(args, model)
    batch = next(iter(m.dataloader))
    model.train()
    losses = [m.train_step(model, opt, batch)[0].item() for _ in range(30)]
    assert np.isfinite(losses).all() and np.mean(losses[-5:]) < 0.7 * np.mean(losses[:5]), losses
