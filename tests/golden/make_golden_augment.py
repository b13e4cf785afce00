"""Golden vectors of the reference's joint geometric augmentation (datasets/base_dataset.py:48-127) - build container only:

    python tests/golden/make_golden_augment.py     # needs /root/reference  ->  tests/golden/augment_golden.npz

BaseDataset._geometric_augmentations (random scale 0.5-2.0 with PIL BILINEAR for the image / PIL NEAREST for the label map /
torch nearest for the query masks, pad to the crop size, random crop, random horizontal flip) is run UNMODIFIED on small
synthetic images with Python's `random` seeded, for several seeds and both an up- and a down-scaling draw.  Stored per case:
the inputs, the draws it consumed (scale, crop offsets, flip - re-derived by replaying `random` with the same seed), and
the outputs (image uint8, label map, query mask, human-label map)."""
import os
import random
import sys

import numpy as np
import torch
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
sys.dont_write_bytecode = True
sys.path.insert(0, "/root/reference")
import importlib.util  # noqa: E402

# the reference's `datasets` directory has no __init__.py and loses against an installed `datasets` package: load the file
_spec = importlib.util.spec_from_file_location("ref_base_dataset", "/root/reference/datasets/base_dataset.py")
_mod = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_mod)
BaseDataset = _mod.BaseDataset

OUT = os.path.join(HERE, "augment_golden.npz")
CASES = [dict(h=64, w=96, crop=(48, 80), seed=s) for s in range(6)] + [dict(h=40, w=56, crop=(64, 64), seed=11),
                                                                      dict(h=96, w=128, crop=(96, 128), seed=12)]


def main():
    out = {"n_cases": np.array(len(CASES))}
    for ci, c in enumerate(CASES):
        rs = np.random.RandomState(100 + ci)
        h, w = c["h"], c["w"]
        x = rs.randint(0, 256, size=(h, w, 3)).astype(np.uint8)
        # smooth the image a little so that resampling differences of 1 LSB are visible but not everything is noise
        x = ((x.astype(np.float32) + np.roll(x, 1, 0) + np.roll(x, 1, 1)) / 3).astype(np.uint8)
        y = rs.randint(0, 19, size=(h, w)).astype(np.uint8)
        q = (rs.rand(h, w) < 0.05).astype(np.uint8) * 255
        lq = np.where(q > 0, y, 19).astype(np.uint8)
        ds = BaseDataset()
        ds.geometric_augmentations = {"random_scale": True, "crop": True, "random_hflip": True}
        ds.crop_size, ds.mean_val, ds.ignore_index = c["crop"], (73, 83, 72), 19
        random.seed(c["seed"])
        xo, yo, qo, lqo = ds._geometric_augmentations(Image.fromarray(x), Image.fromarray(y), torch.from_numpy(q),
                                                      torch.from_numpy(lq))
        # the draws, replayed: uniform(0.5, 2.0), randint x2, random()
        random.seed(c["seed"])
        scale = random.uniform(0.5, 2.0)
        w_rs, h_rs = int(w * scale), int(h * scale)
        hp, wp = max(h_rs, c["crop"][0]), max(w_rs, c["crop"][1])
        sh, sw = random.randint(0, hp - c["crop"][0]), random.randint(0, wp - c["crop"][1])
        flip = random.random() > 0.5
        out[f"c{ci}_cfg"] = np.array([h, w, c["crop"][0], c["crop"][1], c["seed"], h_rs, w_rs, sh, sw, int(flip)])
        out[f"c{ci}_scale"] = np.array(scale)
        out[f"c{ci}_x"], out[f"c{ci}_y"], out[f"c{ci}_q"], out[f"c{ci}_lq"] = x, y, q, lq
        out[f"c{ci}_xo"] = np.asarray(xo)
        out[f"c{ci}_yo"] = np.asarray(yo)
        out[f"c{ci}_qo"] = qo.numpy()
        out[f"c{ci}_lqo"] = np.asarray(lqo)
        print(ci, c, "scale", round(scale, 4), "resized", (h_rs, w_rs), "crop at", (sh, sw), "flip", flip,
              out[f"c{ci}_xo"].shape, out[f"c{ci}_qo"].shape, out[f"c{ci}_qo"].max())
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
