"""Mirror of the reference `eval.py` (evaluate(), eval.py:15-94, and its CLI, eval.py:97-134) on the B200 hot path.

The reference forwards one validation image at a time, materialises the full-resolution logits, takes argmax, copies the
label and prediction maps to the host and bins them with NumPy (RunningScore.update).  Here same-sized images are
micro-batched, the encoder runs on the fused inference kernels, and ONE kernel (`pp_eval_confusion_upsampled`) evaluates
the x4 bilinear upsample, the argmax and the confusion-matrix update on the device; the matrix is read once at the end.
Not mirrored: the PNG visualiser hook (host-side drawing; a `visualizer` callable is accepted and ignored with a note).
"""
import os
from copy import deepcopy
from math import ceil
from typing import Optional

import numpy as np
import torch
import torch.nn.functional as F

from . import _lib
from .utils import RunningScore, write_log


@torch.no_grad()
def confusion_over_loader(model, dataloader, n_classes: int, device, dataset_name: str = "", stride_total: int = 8,
                          debug: bool = False, batch_imgs: int = 16) -> np.ndarray:
    """Confusion matrix [n_classes, n_classes] of `model` over a (batch_size 1) validation dataloader — identical counts to
    RunningScore.update over the full maps (utils/metrics.py:162-177)."""
    model.eval()
    conf = torch.zeros((n_classes, n_classes), dtype=torch.int64, device=device)
    host = RunningScore(n_classes)  # only used by the fallback path
    fused_ok = hasattr(model, "forward_lowres") and n_classes in (11, 19, 21) and dataset_name != "voc"
    pending = []

    def flush():
        xs = torch.cat([b[0] for b in pending], dim=0).to(device, non_blocking=True)
        ys = torch.cat([b[1] for b in pending], dim=0)
        h, w = ys.shape[1:]
        if fused_ok and tuple(xs.shape[2:]) == (h, w):
            lowres = model.forward_lowres(xs)
            _lib.eval_confusion_upsampled(lowres, (h, w), ys.to(device, non_blocking=True).contiguous(), conf)
        else:  # VOC: reflect-pad to a stride multiple, crop the prediction back (eval.py:49-55)
            if dataset_name == "voc":
                pad_h = ceil(h / stride_total) * stride_total - xs.shape[2]
                pad_w = ceil(w / stride_total) * stride_total - xs.shape[3]
                xs = F.pad(xs, pad=(0, pad_w, 0, pad_h), mode="reflect")
            pred = model(xs)["pred"][:, :, :h, :w].argmax(dim=1)
            host.update(ys.numpy(), pred.cpu().numpy())
        pending.clear()

    for dict_data in dataloader:
        x, y = dict_data["x"], dict_data["y"]
        if pending and (pending[0][0].shape[1:] != x.shape[1:] or pending[0][1].shape[1:] != y.shape[1:]
                        or sum(b[0].shape[0] for b in pending) >= batch_imgs):
            flush()
        pending.append((x, y))
        if debug:
            break
    if pending:
        flush()
    return conf.cpu().numpy().astype(np.float64) + host.confusion_matrix


@torch.no_grad()
def evaluate(model, dataloader, experim_name: str, epoch: Optional[int] = None, dir_ckpt: Optional[str] = None,
             visualizer: Optional[callable] = None, visualize_interval: Optional[int] = 100, stride_total: int = 8,
             device: torch.device = torch.device("cuda:0"), debug: bool = False, batch_imgs: int = 16):
    """eval.py:15-94 — returns the mean IoU over the dataloader's validation set; writes `log_val.txt` under dir_ckpt."""
    if dir_ckpt is not None:
        dir_ckpt = f"{dir_ckpt}/e{epoch:02d}/val" if epoch is not None else f"{dir_ckpt}/val"
        os.makedirs(dir_ckpt, exist_ok=True)
    if visualizer is not None:
        print("pixelpick_b200.eval: the PNG visualiser hook is not mirrored (host-side drawing); ignoring it")
    ds = dataloader.dataset
    tracker = RunningScore(ds.n_classes)
    tracker.update_confusion(confusion_over_loader(model, dataloader, ds.n_classes, device,
                                                   getattr(ds, "dataset_name", ""), stride_total, debug, batch_imgs))
    scores = tracker.get_scores()[0]
    miou, pixel_acc = scores["Mean IoU"], scores["Pixel Acc"]
    if dir_ckpt is not None:
        write_log(f"{dir_ckpt}/log_val.txt", header=["epoch", "miou", "pixel_acc"])
        write_log(f"{dir_ckpt}/log_val.txt", list_entities=[epoch, miou, pixel_acc])
    print(f"\n{'=' * 100}\nExperim name: {experim_name}\nEpoch {epoch} | miou: {miou:.3f} | pixel_acc.: {pixel_acc:.3f}"
          f"\n{'=' * 100}\n")
    return miou


def main(argv=None):
    """eval.py:97-134: `python -m pixelpick_b200.eval --dataset_name cs --p_state_dict best_miou_model.pt`."""
    from .args import Arguments
    from .utils import get_dataloader, get_model
    parser = Arguments()
    parser.parser.add_argument("--p_state_dict", type=str, default="", help="path to a state_dict file")
    args = parser.parse_args(argv=argv, verbose=True) if argv is not None else parser.parse_args(verbose=True)
    dataloader = get_dataloader(deepcopy(args), val=True, query=False, shuffle=False, batch_size=1, n_workers=args.n_workers)
    device = torch.device("cuda", torch.cuda.current_device())
    model = get_model(args).to(device)
    if args.p_state_dict:
        model.load_state_dict(torch.load(args.p_state_dict, map_location=device)["model"])
    return evaluate(model=model, dataloader=dataloader, dir_ckpt=getattr(args, "dir_checkpoints", None),
                    experim_name=args.experim_name, stride_total=args.stride_total, device=device, debug=args.debug)


if __name__ == "__main__":
    main()
