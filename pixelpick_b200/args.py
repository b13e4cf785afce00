"""Mirror of the reference `args.py` (argparse surface + per-dataset hyper-parameters, args.py:10-205) with the two
fixes the shipped entry script needs: `--p_dataset_config` is registered (default None), and
`--network_name deeplab_rn50` selects the RN50-DeepLabv3+ composition.  `--synthetic N H W` replaces the datasets."""
import os
import random
from argparse import ArgumentParser, Namespace
from pprint import pformat

import numpy as np
import torch


class Arguments:
    def __init__(self):
        p = ArgumentParser("PixelPick")
        p.add_argument("--debug", "-d", action="store_true", default=False)
        p.add_argument("--dir_root", type=str, default="..")
        p.add_argument("--dir_checkpoints", type=str, default="")
        p.add_argument("--gpu_ids", type=str, nargs="+", default="0")
        p.add_argument("--n_workers", type=int, default=4)
        p.add_argument("--network_name", type=str, default="deeplab", choices=["deeplab", "deeplab_rn50", "FPN"])
        p.add_argument("--seed", "-s", type=int, default=0)
        p.add_argument("--suffix", type=str, default="")
        p.add_argument("--n_pixels_by_us", type=int, default=10)
        p.add_argument("--top_n_percent", type=float, default=0.05)
        p.add_argument("--query_strategy", "-qs", type=str, default="margin_sampling",
                       choices=["least_confidence", "margin_sampling", "entropy", "random"])
        p.add_argument("--reverse_order", action="store_true", default=False)
        p.add_argument("--use_mc_dropout", action="store_true", default=False)
        p.add_argument("--mc_dropout_p", type=float, default=0.2)
        p.add_argument("--mc_n_steps", type=int, default=20)
        p.add_argument("--vote_type", type=str, default="soft", choices=["soft", "hard"])
        p.add_argument("--n_init_pixels", type=int, default=0)
        p.add_argument("--max_budget", type=int, default=100)
        p.add_argument("--nth_query", type=int, default=1)
        p.add_argument("--dataset_name", type=str, default="cv", choices=["cs", "cv", "voc"])
        p.add_argument("--dir_datasets", type=str, default="/scratch/shared/beegfs/gyungin/datasets")
        p.add_argument("--downsample", type=int, default=4)
        p.add_argument("--use_aug", type=bool, default=True)
        p.add_argument("--use_augmented_dataset", action="store_true", default=False)
        p.add_argument("--n_layers", type=int, default=50, choices=[18, 34, 50, 101])
        p.add_argument("--use_dilated_resnet", type=bool, default=True)
        p.add_argument("--weight_type", type=str, default="supervised", choices=["random", "supervised", "moco_v2"])
        p.add_argument("--width_multiplier", type=float, default=1.0)
        p.add_argument("--p_dataset_config", "-pdc", type=str, default=None)  # read at args.py:79, never registered there
        p.add_argument("--synthetic", type=int, nargs=3, default=None, metavar=("N", "H", "W"))
        p.add_argument("--n_epochs", type=int, default=None, help="override the per-dataset default (50)")
        p.add_argument("--no_cuda_graph", dest="cuda_graph", action="store_false", default=True,
                       help="run the train step eagerly instead of replaying one captured CUDA graph")
        self.parser = p

    @staticmethod
    def dataset_defaults(args: Namespace) -> Namespace:
        """args.py:88-150."""
        adam = {"lr": 5e-4, "betas": (0.9, 0.999), "weight_decay": 2e-4, "eps": 1e-7}
        n_epochs = getattr(args, "n_epochs", None)
        if args.dataset_name == "cs":
            args.batch_size, args.ignore_index, args.n_classes = 4, 19, 19
            args.dir_dataset = "/scratch/shared/beegfs/gyungin/datasets/cityscapes"
            args.mean, args.std = [0.28689554, 0.32513303, 0.28389177], [0.18696375, 0.19017339, 0.18720214]
            args.optimizer_type, args.lr_scheduler_type, args.optimizer_params = "Adam", "Poly", adam
        elif args.dataset_name == "cv":
            args.batch_size, args.ignore_index, args.n_classes, args.downsample = 4, 11, 11, 1
            args.dir_dataset = "/Users/noel/Desktop/pixelpick/pixelpick_via_launch/camvid"
            args.mean = [0.41189489566336, 0.4251328133025, 0.4326707089857]
            args.std = [0.27413549931506, 0.28506257482912, 0.28284674400252]
            args.optimizer_type, args.lr_scheduler_type, args.optimizer_params = "Adam", "MultiStepLR", adam
        elif args.dataset_name == "voc":
            args.batch_size, args.ignore_index, args.n_classes = 10, 255, 21
            args.dir_dataset = "/scratch/shared/beegfs/gyungin/datasets/VOC2012"
            args.dir_augmented_dataset = f"{args.dir_dataset}/VOCdevkit/VOC2012/train_aug"
            args.mean, args.std = [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]
            args.size_base, args.size_crop = 400, 320
            args.optimizer_type, args.lr_scheduler_type = "SGD", "Poly"
            args.optimizer_params = {"lr": 1e-2, "weight_decay": 1e-4, "momentum": 0.9}
        else:
            raise ValueError(f"Unsupported dataset name: {args.dataset_name}")
        args.n_epochs = 50 if n_epochs is None else n_epochs
        return args

    def parse_args(self, verbose: bool = False, argv=None):
        args = self.parser.parse_args(argv)
        aug = args.use_aug  # args.py:63-76: consumed by the dataset readers (out of scope here), kept for a plugged-in dataset
        args.augmentations = {"geometric": {"random_scale": aug, "random_hflip": aug, "crop": aug},
                              "photometric": {"random_color_jitter": aug, "random_grayscale": aug, "random_gaussian_blur": aug}}
        args.stride_total = 8 if args.use_dilated_resnet else 32
        if args.p_dataset_config is not None:
            import yaml
            assert os.path.exists(args.p_dataset_config), FileNotFoundError(args.p_dataset_config)
            d = vars(args)
            d.update(yaml.safe_load(open(args.p_dataset_config, "r")))
            args = Namespace(**d)
        else:
            args = self.dataset_defaults(args)
        kw = [args.dataset_name] + ([f"d{args.downsample}"] if args.dataset_name == "cs" else []) + [args.network_name]
        if args.n_pixels_by_us > 0:
            kw.append(args.query_strategy)
            if args.use_mc_dropout:
                kw.append(args.vote_type)
            kw.append(f"{args.n_pixels_by_us}")
            if args.top_n_percent > 0.0:
                kw.append(f"p{args.top_n_percent}")
            if args.reverse_order:
                kw.append("reverse")
        else:
            kw.append("fully_sup")
        kw.append(str(args.seed))
        if args.suffix != "":
            kw.append(args.suffix)
        if args.debug:
            kw.append("debug")
        args.experim_name = "_".join(kw)
        if args.dir_checkpoints == "":
            args.dir_checkpoints = f"{args.dir_root}/checkpoints/{args.experim_name}"
        os.makedirs(args.dir_checkpoints, exist_ok=True)
        with open(f"{args.dir_checkpoints}/args.txt", "w") as f:
            f.write(pformat(vars(args)))
        print(f"\nmodel name: {args.experim_name}\n")
        for fn in (random.seed, np.random.seed, torch.manual_seed):
            fn(args.seed)
        # args.py:197 turns the autotuner on; PP_CUDNN_BENCHMARK=0 (set by tests/conftest.py) skips the minute it costs to
        # tune ~50 conv shapes in short runs
        torch.backends.cudnn.benchmark = os.environ.get("PP_CUDNN_BENCHMARK", "1") != "0"
        if verbose:
            for k, v in sorted(vars(args).items()):
                print(k, v)
        return args
