"""Golden vectors of the reference's photometric augmentation (datasets/base_dataset.py:129-141, GaussianBlur :192-210) - build
container only:

    python tests/golden/make_golden_photometric.py     # needs /root/reference, torchvision, cv2  ->  photometric_golden.npz

`BaseDataset._photometric_augmentations` (RandomApply([ColorJitter(0.8, 0.8, 0.8, 0.2)], 0.8) -> RandomGrayscale(0.2) -> cv2
GaussianBlur with p = 0.5) is run UNMODIFIED on small synthetic RGB images with torch's and NumPy's global streams seeded.
Stored per case: the seed, the image size and the output; the input is regenerated from the seed (`make_input`).  Library
versions that produced the file are stored too (Pillow / torchvision / OpenCV define the arithmetic the oracle restates)."""
import importlib.util
import os
import sys

import numpy as np
import torch
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
sys.dont_write_bytecode = True
sys.path.insert(0, "/root/reference")
_spec = importlib.util.spec_from_file_location("ref_base_dataset", "/root/reference/datasets/base_dataset.py")
_mod = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_mod)
BaseDataset = _mod.BaseDataset

OUT = os.path.join(HERE, "photometric_golden.npz")
CASES = [dict(h=40, w=56, seed=s) for s in range(20)] + [dict(h=96, w=128, seed=100 + s) for s in range(6)]


def make_input(h, w, seed):
    rs = np.random.RandomState(1000 + seed)
    x = rs.randint(0, 256, size=(h, w, 3)).astype(np.uint8)
    x = ((x.astype(np.float32) + np.roll(x, 1, 0) + np.roll(x, 1, 1)) / 3).astype(np.uint8)  # some spatial structure
    x[0, :, :] = rs.randint(0, 256, size=(w, 3))  # and a row of saturated noise (hue wrap-around, clipping)
    return x


def main():
    import cv2
    import PIL
    import torchvision
    out = {"n_cases": np.array(len(CASES)),
           "versions": np.array([f"Pillow {PIL.__version__}", f"torchvision {torchvision.__version__}", f"opencv {cv2.__version__}"])}
    kinds = {"jitter": 0, "gray": 0, "blur": 0}
    for ci, c in enumerate(CASES):
        x = make_input(c["h"], c["w"], c["seed"])
        ds = BaseDataset()
        ds.photometric_augmentations = {"random_color_jitter": True, "random_grayscale": True, "random_gaussian_blur": True}
        torch.manual_seed(c["seed"])
        np.random.seed(c["seed"])
        y = np.asarray(ds._photometric_augmentations(Image.fromarray(x)))
        assert y.dtype == np.uint8 and y.shape == x.shape
        out[f"c{ci}_cfg"] = np.array([c["h"], c["w"], c["seed"]])
        out[f"c{ci}_out"] = y
        print(ci, c, "changed px", int((y != x).any(-1).sum()))
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
