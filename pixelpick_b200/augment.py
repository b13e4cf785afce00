"""Input pipeline on the GPU - host side of `pp_augment_geometric`.

The reference augments every training sample in DataLoader workers with PIL / torchvision (datasets/base_dataset.py:48-127):
random scale in [0.5, 2.0] (image BILINEAR, label map NEAREST, query masks nearest through torchvision's TENSOR path), pad to
the crop size, random crop, random horizontal flip, then to_tensor + normalize.  Here the raw uint8 batch goes to the device
once and ONE kernel produces the normalised crop, the label crop and the mask crops.

What stays on the host, per sample: the four random draws - Python's `random` in the reference's call order (`uniform`,
`randint`, `randint`, `random`), so a seeded run draws the same numbers - and the resampling tables, built in double precision
exactly as Pillow builds them (Resample.c:precompute_coeffs + normalize_coeffs_8bpc; Geometry.c's running-sum nearest
indices), vectorised with NumPy.  The kernel is then pure integer arithmetic and reproduces PIL's output bit for bit
(tests/test_augment_gpu.py against golden outputs of the reference method)."""
import math
import random as _random
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib

PRECISION_BITS = 32 - 8 - 2
_HDR = 20


def draw_geometric(h: int, w: int, crop: Tuple[int, int], rng=_random, random_scale=True, do_crop=True, random_hflip=True):
    """The draws of BaseDataset._geometric_augmentations for one sample, in its order -> (scale, start_h, start_w, flip)."""
    scale = rng.uniform(0.5, 2.0) if random_scale else 1.0
    w_rs, h_rs = int(w * scale), int(h * scale)
    sh = sw = 0
    if do_crop:
        hp, wp = max(h_rs, crop[0]), max(w_rs, crop[1])
        sh, sw = rng.randint(0, hp - crop[0]), rng.randint(0, wp - crop[1])
    flip = (rng.random() > 0.5) if random_hflip else False
    return scale, sh, sw, flip


def _bilinear_tables(in_size: int, out_size: int):
    """Resample.c:precompute_coeffs (bilinear, support 1) + normalize_coeffs_8bpc, vectorised over the output axis."""
    scale = filterscale = in_size / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = 1.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    ss = 1.0 / filterscale
    center = (np.arange(out_size, dtype=np.float64) + 0.5) * scale
    lo = np.maximum((center - support + 0.5).astype(np.int64), 0)  # (int) truncates toward zero, as astype does
    hi = np.minimum((center + support + 0.5).astype(np.int64), in_size)
    n = hi - lo
    xs = np.arange(ksize, dtype=np.float64)[None, :]
    t = np.abs((xs + lo[:, None] - center[:, None] + 0.5) * ss)
    w = np.where((t < 1.0) & (xs < n[:, None]), 1.0 - t, 0.0)
    ww = np.zeros(out_size, dtype=np.float64)
    for c in range(ksize):  # left-to-right sum, the order of the C loop
        ww = ww + np.where(c < n, w[:, c], 0.0)
    v = np.where(ww[:, None] != 0.0, w / np.where(ww[:, None] != 0.0, ww[:, None], 1.0), w)
    kk = (0.5 + v * (1 << PRECISION_BITS)).astype(np.int64)
    kk = np.where(xs < n[:, None], kk, 0).astype(np.int32)
    return lo.astype(np.int32), n.astype(np.int32), kk, ksize


def _pil_nearest(in_size: int, out_size: int):
    a = in_size / out_size
    xo = np.cumsum(np.concatenate(([a * 0.5], np.full(out_size - 1, a, dtype=np.float64))))  # sequential double sum
    return np.clip(xo.astype(np.int64), 0, in_size - 1).astype(np.int32)


def _torch_nearest(in_size: int, out_size: int):
    scale = np.float32(in_size) / np.float32(out_size)
    idx = np.floor(np.arange(out_size, dtype=np.float32) * scale).astype(np.int64)
    return np.minimum(idx, in_size - 1).astype(np.int32)


def build_tables(H: int, W: int, params: Sequence[Tuple[float, int, int, bool]]):
    """-> (header int32 [B, 20], tables int32 [total]) for pp_augment_geometric."""
    hdr = np.zeros((len(params), _HDR), dtype=np.int32)
    chunks: List[np.ndarray] = []
    off = 0

    def push(a):
        nonlocal off
        a = np.ascontiguousarray(a, dtype=np.int32).reshape(-1)
        chunks.append(a)
        o = off
        off += a.size
        return o

    for b, (scale, sh, sw, flip) in enumerate(params):
        w_rs, h_rs = int(W * scale), int(H * scale)
        xlo, xn, xk, ksx = _bilinear_tables(W, w_rs)
        ylo, yn, yk, ksy = _bilinear_tables(H, h_rs)
        hdr[b, :7] = (h_rs, w_rs, sh, sw, int(flip), ksx, ksy)
        hdr[b, 7] = push(_pil_nearest(W, w_rs))
        hdr[b, 8] = push(_pil_nearest(H, h_rs))
        hdr[b, 9] = push(_torch_nearest(W, w_rs))
        hdr[b, 10] = push(_torch_nearest(H, h_rs))
        hdr[b, 11], hdr[b, 12], hdr[b, 13] = push(xlo), push(xn), push(xk)
        hdr[b, 14], hdr[b, 15], hdr[b, 16] = push(ylo), push(yn), push(yk)
    return hdr, np.concatenate(chunks)


class GpuGeometricAugment:
    """`x, y, queries, labelled = aug(x_u8, y_u8, q_u8, lq_u8, params)` on CUDA tensors:
    x_u8 uint8 [B, H, W, 3] (HWC, as `np.asarray(Image)`), y_u8 / q_u8 / lq_u8 uint8 [B, H, W] or None, params = one
    `draw_geometric(...)` tuple per sample -> x float32 [B, 3, ch, cw] normalised, y / labelled uint8 [B, ch, cw], queries
    uint8 (0 / 1)."""

    def __init__(self, crop_size, mean, std, ignore_index, mean_val=None):
        self.crop = (int(crop_size[0]), int(crop_size[1]))
        self.mean = torch.tensor(list(mean), dtype=torch.float32)
        self.std = torch.tensor(list(std), dtype=torch.float32)
        mv = mean_val if mean_val is not None else tuple(int(m * 255) for m in mean)  # cityscapes.py:60 / camvid.py:52
        self.mean_val = torch.tensor(list(mv), dtype=torch.int32)
        self.ignore_index = int(ignore_index)

    def __call__(self, x_u8, y_u8=None, q_u8=None, lq_u8=None, params=None):
        _lib._need_cuda(x_u8, y_u8, q_u8, lq_u8)
        B, H, W, C3 = x_u8.shape
        assert C3 == 3 and x_u8.dtype == torch.uint8 and x_u8.is_contiguous() and len(params) == B
        for t in (y_u8, q_u8, lq_u8):
            assert t is None or (t.dtype == torch.uint8 and t.is_contiguous() and tuple(t.shape) == (B, H, W))
        hdr, tab = build_tables(H, W, params)
        dev = x_u8.device
        hdr_d = torch.from_numpy(hdr).to(dev, non_blocking=True)
        tab_d = torch.from_numpy(tab).to(dev, non_blocking=True)
        ch, cw = self.crop
        x_out = torch.empty((B, 3, ch, cw), dtype=torch.float32, device=dev)
        mk = lambda src: torch.empty((B, ch, cw), dtype=torch.uint8, device=dev) if src is not None else None
        y_out, q_out, lq_out = mk(y_u8), mk(q_u8), mk(lq_u8)
        _lib.check(_lib.lib().pp_augment_geometric(
            _lib._ptr(x_u8), _lib._ptr(y_u8), _lib._ptr(q_u8), _lib._ptr(lq_u8), B, H, W, _lib._ptr(hdr_d), _lib._ptr(tab_d), ch, cw,
            self.mean.data_ptr(), self.std.data_ptr(), self.mean_val.data_ptr(), self.ignore_index, _lib._ptr(x_out),
            _lib._ptr(y_out), _lib._ptr(q_out), _lib._ptr(lq_out), _lib._stream(x_u8)), "pp_augment_geometric")
        return x_out, y_out, q_out, lq_out
