"""Active-learning entry point (reference: `main_al.py`, which seeds the RNGs and runs `Model(args)()`).

    python -m pixelpick_b200.main_al --dataset_name cv --n_pixels_by_us 10 -qs margin_sampling --synthetic 32 256 512
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 -m pixelpick_b200.main_al ...   # one process per GPU

Single process: `--gpu_ids` selects the device through CUDA_VISIBLE_DEVICES as in the reference (main_al.py:9).  Under
torchrun every rank takes the GPU of its LOCAL_RANK and joins one NCCL group; the train loader is then sharded over ranks
(utils.make_loader), gradients are all-reduced once per step (dist.GradAllReducer) and the query round shards the images
(query.QuerySelector).  All ranks use the same seed: model initialisation, query draws and label sets stay identical."""
import os
import random

import numpy as np
import torch


def join_process_group():
    """-> (rank, world).  No-op outside torchrun."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world == 1:
        return 0, 1
    import torch.distributed as dist
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if not dist.is_initialized():
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if os.environ.get("OMP_NUM_THREADS") == "1":  # torchrun's default: one host thread per rank starves the loader / staging copies
        torch.set_num_threads(max(1, (os.cpu_count() or world) // world))
    return dist.get_rank(), world


def main(args):
    if int(os.environ.get("WORLD_SIZE", "1")) == 1:
        os.environ["CUDA_VISIBLE_DEVICES"] = ",".join(args.gpu_ids[0])
    rank, world = join_process_group()
    random.seed(args.seed), np.random.seed(args.seed), torch.manual_seed(args.seed)
    torch.backends.cudnn.benchmark = os.environ.get("PP_CUDNN_BENCHMARK", "1") != "0"
    from .model import Model
    try:
        Model(args)()
    finally:
        if world > 1:
            torch.distributed.destroy_process_group()


if __name__ == "__main__":
    from .args import Arguments
    main(Arguments().parse_args())
