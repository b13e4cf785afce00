"""Command-line surface and per-dataset hyper-parameters of the reference (`args.py:10-205`), table-driven.

`Arguments().parse_args()` yields a Namespace with every field the reference produces, same names and values (pinned by
tests/test_host_golden.py against the reference's own output), plus what the shipped entry scripts need:
`--p_dataset_config` is registered (the reference reads it at args.py:79 without ever registering it),
`--network_name deeplab_rn50` selects the RN50-DeepLabv3+ composition, `--synthetic N H W` stands in for the dataset files,
`--n_epochs` overrides the default of 50 and `--no_cuda_graph` runs the train step eagerly.
Not mirrored: exporting CUDA_VISIBLE_DEVICES from `--gpu_ids` (args.py:193) - main_al does that for single-process runs only,
a one-process-per-GPU launcher must not."""
import os
import random
from argparse import ArgumentParser, Namespace
from pprint import pformat

import numpy as np
import torch

_STRATEGIES = ["least_confidence", "margin_sampling", "entropy", "random"]
_FLAG = dict(action="store_true", default=False)

# (flags, argparse keywords): args.py:14-57, in the reference's order
_SURFACE = [
    (("--debug", "-d"), _FLAG),
    (("--dir_root",), dict(type=str, default="..")),
    (("--dir_checkpoints",), dict(type=str, default="")),
    (("--gpu_ids",), dict(type=str, nargs="+", default="0")),
    (("--n_workers",), dict(type=int, default=4)),
    (("--network_name",), dict(type=str, default="deeplab", choices=["deeplab", "deeplab_rn50", "FPN"])),
    (("--seed", "-s"), dict(type=int, default=0)),
    (("--suffix",), dict(type=str, default="")),
    # active learning
    (("--n_pixels_by_us",), dict(type=int, default=10, help="pixels per image picked by uncertainty sampling in a round")),
    (("--top_n_percent",), dict(type=float, default=0.05)),
    (("--query_strategy", "-qs"), dict(type=str, default="margin_sampling", choices=_STRATEGIES)),
    (("--reverse_order",), _FLAG),
    # query by committee (MC dropout): accepted; the reference branch itself is dead code (query.py:186)
    (("--use_mc_dropout",), _FLAG),
    (("--mc_dropout_p",), dict(type=float, default=0.2)),
    (("--mc_n_steps",), dict(type=int, default=20)),
    (("--vote_type",), dict(type=str, default="soft", choices=["soft", "hard"])),
    (("--n_init_pixels",), dict(type=int, default=0)),
    (("--max_budget",), dict(type=int, default=100, help="pixel budget per image")),
    (("--nth_query",), dict(type=int, default=1)),
    # dataset
    (("--dataset_name",), dict(type=str, default="cv", choices=["cs", "cv", "voc"])),
    (("--dir_datasets",), dict(type=str, default="/scratch/shared/beegfs/gyungin/datasets")),
    (("--downsample",), dict(type=int, default=4, help="Cityscapes training-set downsampling")),
    (("--use_aug",), dict(type=bool, default=True)),
    (("--use_augmented_dataset",), _FLAG),
    (("--gpu_augment",), _FLAG),  # ours: the dataset delivers raw uint8 samples, augmentation + normalise run on the device
    # encoder
    (("--n_layers",), dict(type=int, default=50, choices=[18, 34, 50, 101])),
    (("--use_dilated_resnet",), dict(type=bool, default=True)),
    (("--weight_type",), dict(type=str, default="supervised", choices=["random", "supervised", "moco_v2"])),
    (("--width_multiplier",), dict(type=float, default=1.0)),
    # additions of this framework
    (("--p_dataset_config", "-pdc"), dict(type=str, default=None)),
    (("--synthetic",), dict(type=int, nargs=3, default=None, metavar=("N", "H", "W"))),
    (("--n_epochs",), dict(type=int, default=None, help="override the per-dataset default (50)")),
    (("--no_cuda_graph",), dict(dest="cuda_graph", action="store_false", default=True,
                                help="run the train step eagerly instead of replaying one captured CUDA graph")),
]

_ADAM = {"lr": 5e-4, "betas": (0.9, 0.999), "weight_decay": 2e-4, "eps": 1e-7}
_VOC_ROOT = "/scratch/shared/beegfs/gyungin/datasets/VOC2012"
# args.py:88-150.  The directory defaults are the reference's (its author's machines); a real run overrides them.
_DATASETS = {
    "cs": dict(batch_size=4, dir_dataset="/scratch/shared/beegfs/gyungin/datasets/cityscapes", ignore_index=19, n_classes=19,
               mean=[0.28689554, 0.32513303, 0.28389177], std=[0.18696375, 0.19017339, 0.18720214],
               optimizer_type="Adam", lr_scheduler_type="Poly", optimizer_params=_ADAM),
    "cv": dict(batch_size=4, dir_dataset="/Users/noel/Desktop/pixelpick/pixelpick_via_launch/camvid", downsample=1,
               ignore_index=11, n_classes=11, mean=[0.41189489566336, 0.4251328133025, 0.4326707089857],
               std=[0.27413549931506, 0.28506257482912, 0.28284674400252],
               optimizer_type="Adam", lr_scheduler_type="MultiStepLR", optimizer_params=_ADAM),
    "voc": dict(batch_size=10, dir_dataset=_VOC_ROOT, dir_augmented_dataset=f"{_VOC_ROOT}/VOCdevkit/VOC2012/train_aug",
                ignore_index=255, n_classes=21, mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225], size_base=400, size_crop=320,
                optimizer_type="SGD", lr_scheduler_type="Poly", optimizer_params={"lr": 1e-2, "weight_decay": 1e-4, "momentum": 0.9}),
}


def experiment_name(a: Namespace) -> str:
    """args.py:152-181: dataset[_d<downsample>]_network[_layers_weights]_strategy[_vote]_n[_p<top>][_reverse]_seed[_suffix][_debug]."""
    parts = [a.dataset_name]
    if a.dataset_name == "cs":
        parts.append(f"d{a.downsample}")
    parts.append(a.network_name)
    if a.network_name == "FPN":
        parts += [f"{a.n_layers}", f"{a.weight_type}"]
    if a.n_pixels_by_us > 0:
        parts.append(a.query_strategy)
        if a.use_mc_dropout:
            parts.append(a.vote_type)
        parts.append(str(a.n_pixels_by_us))
        if a.top_n_percent > 0.0:
            parts.append(f"p{a.top_n_percent}")
        if a.reverse_order:
            parts.append("reverse")
    else:
        parts.append("fully_sup")
    parts.append(str(a.seed))
    parts += [a.suffix] if a.suffix != "" else []
    parts += ["debug"] if a.debug else []
    return "_".join(parts)


class Arguments:
    def __init__(self):
        self.parser = ArgumentParser("PixelPick")
        for flags, kw in _SURFACE:
            self.parser.add_argument(*flags, **kw)

    @staticmethod
    def dataset_defaults(args: Namespace) -> Namespace:
        if args.dataset_name not in _DATASETS:
            raise ValueError(f"Unsupported dataset name: {args.dataset_name}")
        n_epochs = getattr(args, "n_epochs", None)
        for k, v in _DATASETS[args.dataset_name].items():
            setattr(args, k, dict(v) if isinstance(v, dict) else list(v) if isinstance(v, list) else v)
        args.n_epochs = 50 if n_epochs is None else n_epochs
        return args

    def parse_args(self, verbose: bool = False, argv=None):
        args = self.parser.parse_args(argv)
        on = args.use_aug  # consumed by the dataset readers (out of scope here); kept for a plugged-in dataset
        args.augmentations = {"geometric": dict.fromkeys(("random_scale", "random_hflip", "crop"), on),
                              "photometric": dict.fromkeys(("random_color_jitter", "random_grayscale", "random_gaussian_blur"), on)}
        args.stride_total = 8 if args.use_dilated_resnet else 32
        if args.p_dataset_config is None:
            args = self.dataset_defaults(args)
        else:  # a YAML file replaces the built-in table (args.py:79-86)
            import yaml
            assert os.path.exists(args.p_dataset_config), FileNotFoundError(args.p_dataset_config)
            args = Namespace(**{**vars(args), **yaml.safe_load(open(args.p_dataset_config, "r"))})
        args.experim_name = experiment_name(args)
        if args.dir_checkpoints == "":
            args.dir_checkpoints = f"{args.dir_root}/checkpoints/{args.experim_name}"
        os.makedirs(args.dir_checkpoints, exist_ok=True)
        with open(f"{args.dir_checkpoints}/args.txt", "w") as f:
            f.write(pformat(vars(args)))
        print(f"\nmodel name: {args.experim_name}\n")
        random.seed(args.seed), np.random.seed(args.seed), torch.manual_seed(args.seed)
        # args.py:197 turns the cuDNN autotuner on; PP_CUDNN_BENCHMARK=0 (set by tests/conftest.py) skips the minute it
        # costs to tune ~50 conv shapes in short runs
        torch.backends.cudnn.benchmark = os.environ.get("PP_CUDNN_BENCHMARK", "1") != "0"
        if verbose:
            for k, v in sorted(vars(args).items()):
                print(k, v)
        return args
