"""Entry point mirroring the reference `main_al.py` (seeds, cudnn.benchmark, Model(args)()); run as
`python -m pixelpick_b200.main_al --dataset_name cv --n_pixels_by_us 10 -qs margin_sampling --synthetic 32 256 512`."""
import os
import random

import numpy as np
import torch


def main(args):
    os.environ["CUDA_VISIBLE_DEVICES"] = ",".join(args.gpu_ids[0])
    for fn in (random.seed, np.random.seed, torch.manual_seed):
        fn(args.seed)
    torch.backends.cudnn.benchmark = True
    from .model import Model
    Model(args)()


if __name__ == "__main__":
    from .args import Arguments
    main(Arguments().parse_args())
