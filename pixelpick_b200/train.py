"""Mirror of the reference `train.py` (train_epoch / train, train.py:14-176, and its CLI, train.py:179-253): the
single-round variant of the loop in model.py — optionally on HUMAN labels (`dict_data["labelled_queries"]`, a dense map that
is ignore_index everywhere except the annotated pixels, train.py:44-45) instead of masked ground truth.

The step is the same as pixelpick_b200.Model.train_step: forward_lowres (fused head) -> fused upsample + sparse CE at the
labelled pixels -> custom backward -> optimiser; running metrics come from the labelled pixels (identical confusion matrix).
Not mirrored: the PNG visualiser hook (accepted, ignored)."""
import os
from copy import deepcopy
from typing import Optional

import torch

from .eval import evaluate
from .loss import sparse_cross_entropy
from .query import gather_previous_query_files, merge_previous_query_files
from .utils import AverageMeter, RunningScore, get_dataloader, get_lr_scheduler, get_model, get_optimizer, write_log


def train_epoch(epoch, dataloader, model, optimizer, lr_scheduler, loss_tracker, experim_name: str,
                dir_ckpt: Optional[str] = None, visualizer: Optional[callable] = None,
                visualize_interval: Optional[int] = 100, human_labels: bool = False,
                device: torch.device = torch.device("cuda:0"), debug: bool = False):
    """train.py:14-103."""
    if dir_ckpt is not None:
        dir_ckpt = f"{dir_ckpt}/e{epoch:02d}"
        os.makedirs(dir_ckpt, exist_ok=True)
    ignore_index: int = dataloader.dataset.ignore_index
    miou_tracker = RunningScore(dataloader.dataset.n_classes)
    model.train()
    miou = pixel_acc = float("nan")
    for dict_data in dataloader:
        x = dict_data["x"].to(device, non_blocking=True)
        if human_labels:  # train.py:44-45: the annotated pixels carry their label, everything else is ignore_index
            y, mask = dict_data["labelled_queries"].to(device, torch.int64), None
        else:  # train.py:47-50
            y, mask = dict_data["y"].to(device), dict_data["queries"].to(device, torch.bool)
        lowres = model.forward_lowres(x)
        loss, pred_at, (_, _, px_label) = sparse_cross_entropy(lowres, y, mask, ignore_index, return_pred=True)
        optimizer.zero_grad(set_to_none=True)
        loss.backward()
        optimizer.step()
        miou_tracker.update_pairs(px_label.cpu().numpy(), pred_at.cpu().numpy())
        loss_tracker.update(loss.detach().item())
        scores = miou_tracker.get_scores()[0]
        miou, pixel_acc = scores["Mean IoU"], scores["Pixel Acc"]
        lr_scheduler.step(epoch=epoch - 1)  # train.py:80: every iteration, whatever the scheduler type
        if debug:
            break
    print(f"({experim_name}) Epoch {epoch} | mIoU.: {miou:.3f} | pixel acc.: {pixel_acc:.3f} | avg loss: {loss_tracker.avg:.3f}")
    if dir_ckpt is not None:
        if epoch == 1:
            write_log(f"{dir_ckpt}/log_train.txt", header=["epoch", "miou", "pixel_acc", "loss"])
        write_log(f"{dir_ckpt}/log_train.txt", list_entities=[epoch, miou, pixel_acc, loss_tracker.avg])
    return model, optimizer, lr_scheduler


def train(args, dataloader, eval_interval: int = 0, dir_ckpt: Optional[str] = None, visualizer: Optional[callable] = None,
          visualize_interval: int = 100, human_labels: bool = False, device: torch.device = torch.device("cuda:0")):
    """train.py:106-176: fresh model, n_epochs of train_epoch, optional evaluation every eval_interval epochs."""
    debug, experim_name, n_epochs = args.debug, args.experim_name, args.n_epochs
    print(f"\n({experim_name}) training...\n")
    model = get_model(args).to(device)
    optimizer = get_optimizer(args, model)
    lr_scheduler = get_lr_scheduler(args, optimizer=optimizer, iters_per_epoch=len(dataloader))
    loss_tracker, best_miou = AverageMeter(), -1.0
    epoch_kw = dict(dir_ckpt=dir_ckpt, visualizer=visualizer, visualize_interval=visualize_interval, human_labels=human_labels,
                    device=device, debug=debug)
    for e in range(1, 1 + n_epochs):
        model, optimizer, lr_scheduler = train_epoch(e, dataloader, model, optimizer, lr_scheduler, loss_tracker,
                                                     experim_name, **epoch_kw)
        if eval_interval > 0 and e % eval_interval == 0:
            val_loader = get_dataloader(deepcopy(args), 1, args.n_workers, False, val=True, query=False)
            current_miou = evaluate(model, val_loader, experim_name, e, dir_ckpt=dir_ckpt, stride_total=args.stride_total,
                                    device=device, debug=debug)
            if current_miou > best_miou and dir_ckpt is not None:  # train.py:171-172 (best_miou is never updated there)
                torch.save({"model": model.state_dict()}, f"{dir_ckpt}/best_model.pt")
        if debug:
            break
    return model


def _has_human_labels(p_queries: str) -> bool:
    """a queries.pkl annotated by a human carries `category_id` per image (via/convert_json_to_pkl.py:20-73)."""
    import pickle
    try:
        d = pickle.load(open(p_queries, "rb"))
        return len(d) > 0 and all("category_id" in info for info in d.values())
    except Exception:
        return False


def main(argv=None):
    """train.py:179-253.  With --dir_checkpoints pointing at earlier `*/queries.pkl` files (human annotations carrying
    `category_id`), the labels are merged (query.py:311-351) and training runs on them (human_labels=True)."""
    from torch.utils.data import DataLoader
    from .args import Arguments
    parser = Arguments()
    parser.parser.add_argument("--eval_interval", type=int, default=1,
                               help="how frequently the model is evaluated in epoch; 0 = never during training")
    args = parser.parse_args(argv=argv, verbose=True) if argv is not None else parser.parse_args(verbose=True)
    device = torch.device("cuda", torch.cuda.current_device())
    prev = [p for p in gather_previous_query_files(args.dir_checkpoints) if _has_human_labels(p)]
    if not prev:
        dataloader = get_dataloader(args=args, batch_size=args.batch_size, shuffle=True, n_workers=args.n_workers)
        nth_query, human = 0, False
    else:
        merged = merge_previous_query_files(prev, ignore_index=args.ignore_index)
        dataset = get_dataloader(args=args, batch_size=args.batch_size, shuffle=True, n_workers=args.n_workers,
                                 generate_init_queries=False).dataset
        if not hasattr(dataset, "update_labelled_queries"):
            raise NotImplementedError("human-label training needs a dataset with the reference's update_labelled_queries() "
                                      "(datasets/base_dataset.py); the synthetic stand-in has no image files to re-index")
        dataset.list_inputs = [p for p, _ in sorted(merged.items())]
        dataset.update_labelled_queries([q for _, q in sorted(merged.items())])
        dataloader = DataLoader(dataset, batch_size=args.batch_size, num_workers=args.n_workers, shuffle=True,
                                drop_last=len(dataset) % args.batch_size == 1)
        nth_query, human = len(prev) - 1, True
    dir_ckpt = args.dir_checkpoints
    args.dir_checkpoints = f"{dir_ckpt}/{nth_query}_query" if args.n_pixels_by_us > 0 else dir_ckpt
    os.makedirs(args.dir_checkpoints, exist_ok=True)
    return train(args, dataloader, eval_interval=args.eval_interval, dir_ckpt=args.dir_checkpoints, human_labels=human,
                 device=device)


if __name__ == "__main__":
    main()
