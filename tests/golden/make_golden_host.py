"""Generate golden vectors for the HOST logic around the train step by running the UNMODIFIED reference.

    python tests/golden/make_golden_host.py       # needs /root/reference (build container only)

The reference (NoelShin/PixelPick @ 43c2981) is imported from /root/reference; nothing is copied.  Output
`tests/golden/host_golden.pkl`:

  args        vars(Arguments().parse_args()) for cs / cv / voc and a few flag combinations (args.py:10-205).  The shipped
              parser never registers `--p_dataset_config` although parse_args reads it (args.py:79); the reference's own
              entry scripts add it (query.py:365), and so does this script.
  optim       get_optimizer (utils/utils.py:112-306) on a stub DeepLab-shaped module for each dataset / optimizer type:
              class name and per-group hyper-parameters
  poly        the learning rates Poly (utils/lr_scheduler.py:4-21) hands out when driven the way model.py:138-139 drives
              it (`step(epoch=epoch-1)` once per iteration), 3 epochs x 5 iterations, two parameter groups
  multistep   the same for the MultiStepLR branch (model.py:144-145, utils.py:309-335)
  score       RunningScore.update / get_scores (utils/metrics.py:162-207) on seeded label maps with void pixels
  meter       AverageMeter (utils/metrics.py) running values
  evaluate    eval.evaluate (eval.py:15-94) with a stub model over a stub 5-class validation loader: returned mIoU and the
              bytes of e03/val/log_val.txt
  train       train.train_epoch (train.py:14-103) driven for 3 epochs on the CPU over a fixed 3-batch loader with a tiny
              DeepLab-shaped network, the reference's get_optimizer (cs: Adam) and Poly schedule: running loss after every
              epoch, final parameters, final learning rates.  dir_ckpt=None (with a directory the reference raises
              NameError: train.py never imports os)
  model_epoch Model._train_epoch (model.py:93-159) on a bare Model object (no datasets / visualiser: `object.__new__`, only the
              attributes the method reads) for cs (Adam + Poly, stepped per iteration) and cv (Adam + MultiStepLR, stepped per
              epoch): the bytes of log_train.txt after 3 epochs and the final parameters
  model_val   Model._val (model.py:177-239) on a bare Model object, four epochs with stub models of varying quality: the bytes
              of log_val.txt and the epochs at which best_miou_model.pt was rewritten
  log         the bytes write_log (utils/utils.py:66-72) leaves in a file after header / rows / header+row calls
"""
import contextlib
import io
import os
import pickle
import sys
import tempfile
import warnings
from argparse import Namespace

import numpy as np
import torch

REF = "/root/reference"
sys.dont_write_bytecode = True
sys.path.insert(0, REF)
import args as refargs  # noqa: E402
import utils.metrics as refmetrics  # noqa: E402
import utils.utils as refutils  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "host_golden.pkl")

ARG_CASES = {
    "cs": ["--dataset_name", "cs"],
    "cv": ["--dataset_name", "cv"],
    "voc": ["--dataset_name", "voc"],
    "cs_entropy_rev": ["--dataset_name", "cs", "-qs", "entropy", "--reverse_order", "--seed", "3", "--suffix", "x"],
    "cv_fully_sup": ["--dataset_name", "cv", "--n_pixels_by_us", "0", "--debug"],
    "cv_top0_mc": ["--dataset_name", "cv", "--top_n_percent", "0", "--use_mc_dropout", "--vote_type", "hard"],
}


def stub_model():
    m = torch.nn.Module()
    m.backbone = torch.nn.Conv2d(3, 4, 1)
    m.aspp = torch.nn.Conv2d(4, 4, 1)
    m.low_level_conv = torch.nn.Conv2d(4, 2, 1)
    m.seg_head = torch.nn.Conv2d(6, 3, 1)
    return m


def groups_of(opt):
    return [{k: v for k, v in g.items() if k != "params" and isinstance(v, (int, float, bool, tuple, type(None)))}
            for g in opt.param_groups]


def drive(sched, opt, kind, n_epochs=3, iters=5):
    out = []
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for epoch in range(1, n_epochs + 1):
            for _ in range(iters):
                out.append([g["lr"] for g in opt.param_groups])
                opt.step()
                if kind == "Poly":
                    sched.step(epoch=epoch - 1)
            if kind == "MultiStepLR":
                sched.step(epoch=epoch - 1)
    out.append([g["lr"] for g in opt.param_groups])
    return out


def main():
    g = {}
    cwd, argv0 = os.getcwd(), sys.argv
    g["args"] = {}
    for name, argv in ARG_CASES.items():
        with tempfile.TemporaryDirectory() as tmp:
            os.chdir(tmp)
            a = refargs.Arguments()
            a.parser.add_argument("--p_dataset_config", "-pdc", type=str, default=None)
            sys.argv = ["x", "--dir_root", "root"] + argv
            with contextlib.redirect_stdout(io.StringIO()):
                ns = a.parse_args()
            os.chdir(cwd)
        g["args"][name] = dict(vars(ns))
    sys.argv = argv0
    os.environ.pop("CUDA_VISIBLE_DEVICES", None)
    torch.backends.cudnn.benchmark = False

    g["optim"] = {}
    for ds, types in {"cs": ["Adam"], "cv": ["Adam", "SGD"], "voc": ["SGD"]}.items():
        for t in types:
            ns = Namespace(**g["args"][ds])
            ns.optimizer_type = t
            opt = refutils.get_optimizer(ns, stub_model())
            g["optim"][f"{ds}_{t}"] = {"cls": type(opt).__name__, "groups": groups_of(opt),
                                       "n_params": [len(pg["params"]) for pg in opt.param_groups]}

    for kind in ("Poly", "MultiStepLR"):
        m = stub_model()
        opt = torch.optim.SGD([{"params": m.backbone.parameters(), "lr": 1e-3}, {"params": m.aspp.parameters(), "lr": 1e-2}])
        ns = Namespace(dataset_name="cs", lr_scheduler_type=kind, n_epochs=3)
        sched = refutils.get_lr_scheduler(ns, opt, iters_per_epoch=5)
        g["poly" if kind == "Poly" else "multistep"] = {"cls": type(sched).__name__, "lrs": drive(sched, opt, kind)}

    rs = np.random.RandomState(5)
    score = refmetrics.RunningScore(11)
    batches = []
    for _ in range(3):
        lt = rs.randint(0, 12, size=(2, 9, 13)).astype(np.int64)   # 11 == void
        lp = rs.randint(0, 11, size=(2, 9, 13)).astype(np.int64)
        score.update(lt, lp)
        batches.append((lt, lp))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        scores, cls_iu = score.get_scores()
    g["score"] = {"batches": batches, "confusion": score.confusion_matrix.copy(), "scores": scores, "cls_iu": cls_iu}

    meter = refmetrics.AverageMeter()
    vals = []
    for v, n in [(0.5, 1), (1.25, 4), (3.0, 2)]:
        meter.update(v, n)
        vals.append({k: float(getattr(meter, k)) for k in ("val", "avg", "sum", "count") if hasattr(meter, k)})
    g["meter"] = vals

    with tempfile.TemporaryDirectory() as tmp:
        fp = os.path.join(tmp, "log.txt")
        refutils.write_log(fp, header=["epoch", "mIoU", "pixel_acc", "loss"])
        refutils.write_log(fp, list_entities=[1, np.float64(0.25), 0.5, 1.75])
        refutils.write_log(fp, list_entities=[2, float("nan"), np.float32(0.5), "x"])
        a = open(fp, "rb").read()
        refutils.write_log(fp, list_entities=[7, 8], header=["a", "b"])
        g["log"] = {"rows": a, "header_and_row": open(fp, "rb").read()}

    import eval as refeval  # the reference module eval.py

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
    with tempfile.TemporaryDirectory() as tmp, contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
        miou = refeval.evaluate(Stub(), Loader(items), "stub", epoch=3, dir_ckpt=tmp, device=torch.device("cpu"))
        g["evaluate"] = {"miou": float(miou), "log": open(os.path.join(tmp, "e03", "val", "log_val.txt"), "rb").read()}

    import train as reftrain  # the reference module train.py
    import torch.nn.functional as F

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

    def train_batches():
        gen = torch.Generator().manual_seed(42)
        out = []
        for _ in range(3):
            q = torch.rand((4, 32, 64), generator=gen) < 0.01
            out.append({"x": torch.randn((4, 3, 32, 64), generator=gen), "y": torch.randint(0, 20, (4, 32, 64), generator=gen),
                        "queries": q.to(torch.uint8)})
        return out

    class TrainDS:
        ignore_index, n_classes = 19, 19

    class TrainLoader(list):
        dataset = TrainDS()

    torch.manual_seed(0)
    model = Tiny(19)
    ns = Namespace(**g["args"]["cs"])
    ns.n_epochs = 3
    loader = TrainLoader(train_batches())
    opt = refutils.get_optimizer(ns, model)
    sched = refutils.get_lr_scheduler(ns, optimizer=opt, iters_per_epoch=len(loader))
    tracker = refmetrics.AverageMeter()
    avg = []
    with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for e in range(1, 4):
            fresh = TrainLoader([{k: v.clone() for k, v in b.items()} for b in loader])  # train_epoch overwrites y in place
            model, opt, sched = reftrain.train_epoch(e, fresh, model, opt, sched, tracker, "golden", device=torch.device("cpu"))
            avg.append(float(tracker.avg))
    g["train"] = {"avg_loss": avg, "params": torch.cat([p.detach().flatten() for p in model.parameters()]).numpy(),
                  "lrs": [pg["lr"] for pg in opt.param_groups]}

    import model as refmodel  # the reference module model.py
    g["model_epoch"] = {}
    for ds in ("cs", "cv"):
        ns = Namespace(**g["args"][ds])
        ns.n_epochs = 3
        nc = ns.n_classes
        gen = torch.Generator().manual_seed(7)
        batches = []
        for _ in range(3):
            q = torch.rand((4, 32, 64), generator=gen) < 0.01
            batches.append({"x": torch.randn((4, 3, 32, 64), generator=gen), "y": torch.randint(0, nc + 1, (4, 32, 64), generator=gen),
                            "queries": q.to(torch.uint8)})

        class EpochDS:
            n_pixels_total = 0

        class EpochLoader(list):
            dataset = EpochDS()

        torch.manual_seed(0)
        net = Tiny(nc)
        with tempfile.TemporaryDirectory() as tmp:
            m = object.__new__(refmodel.Model)
            m.n_pixels_by_us, m.nth_query, m.dir_checkpoints, m.experim_name = 10, 0, tmp, "golden"
            m.device, m.ignore_index, m.debug, m.lr_scheduler_type = torch.device("cpu"), ns.ignore_index, False, ns.lr_scheduler_type
            m.running_loss, m.running_score = refmetrics.AverageMeter(), refmetrics.RunningScore(nc)
            m.vis = lambda dict_tensors, fp=None: None
            m.log_train = os.path.join(tmp, "log_train.txt")
            refutils.write_log(m.log_train, header=["epoch", "mIoU", "pixel_acc", "loss"])
            opt = refutils.get_optimizer(ns, net)
            sched = refutils.get_lr_scheduler(ns, optimizer=opt, iters_per_epoch=3)
            with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()), warnings.catch_warnings():
                warnings.simplefilter("ignore")
                for e in range(1, 4):
                    m.dataloader = EpochLoader([{k: v.clone() for k, v in b.items()} for b in batches])
                    net, opt, sched = m._train_epoch(e, net, opt, sched)
            g["model_epoch"][ds] = {"log": open(m.log_train, "rb").read(),
                                    "params": torch.cat([p.detach().flatten() for p in net.parameters()]).numpy(),
                                    "lrs": [pg["lr"] for pg in opt.param_groups]}

    class ValStub(torch.nn.Module):
        def __init__(self, noise):
            super().__init__()
            self.noise = torch.nn.Parameter(torch.tensor(float(noise)), requires_grad=False)

        def forward(self, x):  # channel c of x carries a one-hot of the true class; `noise` degrades the prediction
            return {"pred": x + self.noise * torch.roll(x, 1, dims=1) * 2.0}

    gen = torch.Generator().manual_seed(3)
    val_items = []
    for _ in range(4):
        y = torch.randint(0, 6, (1, 12, 16), generator=gen)           # 5 == void
        x = torch.nn.functional.one_hot(y.clamp(max=4), 5).permute(0, 3, 1, 2).float() + 0.1 * torch.randn((1, 5, 12, 16), generator=gen)
        val_items.append({"x": x, "y": y})

    class ValLoader(list):
        dataset = None

    with tempfile.TemporaryDirectory() as tmp:
        m = object.__new__(refmodel.Model)
        m.n_pixels_by_us, m.nth_query, m.dir_checkpoints, m.experim_name = 10, 2, tmp, "golden"
        m.device, m.dataset_name, m.stride_total, m.debug, m.best_miou = torch.device("cpu"), "cs", 8, False, -1.0
        m.running_loss, m.running_score = refmetrics.AverageMeter(), refmetrics.RunningScore(5)
        m.vis = lambda dict_tensors, fp=None: None
        m.dataloader_val = ValLoader(val_items)
        os.makedirs(os.path.join(tmp, "2_query"))
        m.log_val = os.path.join(tmp, "2_query", "log_val.txt")
        refutils.write_log(m.log_val, header=["epoch", "mIoU", "pixel_acc"])
        saved = []
        ck = os.path.join(tmp, "2_query", "best_miou_model.pt")
        with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
            for e, noise in ((1, 0.5), (2, 1.0), (3, 0.47), (4, 0.49)):
                before = os.path.getmtime(ck) if os.path.exists(ck) else None
                if os.path.exists(ck):
                    os.remove(ck)
                m._val(e, ValStub(noise))
                if os.path.exists(ck):
                    saved.append(e)
        g["model_val"] = {"log": open(m.log_val, "rb").read(), "saved_at": saved, "best": float(m.best_miou)}

    pickle.dump(g, open(OUT, "wb"), protocol=4)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
