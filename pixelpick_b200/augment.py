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

    def __call__(self, x_u8, y_u8=None, q_u8=None, lq_u8=None, params=None, as_uint8=False):
        """as_uint8: leave the image as uint8 [B, ch, cw, 3] (the PIL image before to_tensor) for the photometric step"""
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
        mk = lambda src: torch.empty((B, ch, cw), dtype=torch.uint8, device=dev) if src is not None else None
        y_out, q_out, lq_out = mk(y_u8), mk(q_u8), mk(lq_u8)
        if as_uint8:
            x_out = torch.empty((B, ch, cw, 3), dtype=torch.uint8, device=dev)
            _lib.check(_lib.lib().pp_augment_geometric_u8(
                _lib._ptr(x_u8), _lib._ptr(y_u8), _lib._ptr(q_u8), _lib._ptr(lq_u8), B, H, W, _lib._ptr(hdr_d), _lib._ptr(tab_d), ch, cw,
                self.mean_val.data_ptr(), self.ignore_index, _lib._ptr(x_out), _lib._ptr(y_out), _lib._ptr(q_out), _lib._ptr(lq_out),
                _lib._stream(x_u8)), "pp_augment_geometric_u8")
            return x_out, y_out, q_out, lq_out
        x_out = torch.empty((B, 3, ch, cw), dtype=torch.float32, device=dev)
        _lib.check(_lib.lib().pp_augment_geometric(
            _lib._ptr(x_u8), _lib._ptr(y_u8), _lib._ptr(q_u8), _lib._ptr(lq_u8), B, H, W, _lib._ptr(hdr_d), _lib._ptr(tab_d), ch, cw,
            self.mean.data_ptr(), self.std.data_ptr(), self.mean_val.data_ptr(), self.ignore_index, _lib._ptr(x_out),
            _lib._ptr(y_out), _lib._ptr(q_out), _lib._ptr(lq_out), _lib._stream(x_u8)), "pp_augment_geometric")
        return x_out, y_out, q_out, lq_out


# =====================================================================================================================
# photometric augmentation (datasets/base_dataset.py:129-141) - host side of `pp_augment_photometric`
# =====================================================================================================================
_PHDR = 16
_JITTER = (0.8, 0.8, 0.8, 0.2)  # ColorJitter(brightness, contrast, saturation, hue) of base_dataset.py:131


def draw_photometric(color_jitter=True, grayscale=True, blur=True):
    """The draws of BaseDataset._photometric_augmentations for one sample, from the same global streams in the same order as
    torchvision / the reference consume them: RandomApply (`torch.rand(1)`), ColorJitter.get_params (`torch.randperm(4)`, four
    `torch.empty(1).uniform_`), RandomGrayscale (`torch.rand(1)`), GaussianBlur (`np.random.random_sample()` once or twice).
    -> {"jitter": None | (order, brightness, contrast, saturation, hue), "gray": bool, "blur": None | sigma}"""
    jitter = None
    if color_jitter and not (0.8 < torch.rand(1)):
        b, c, s, h = _JITTER
        order = [int(i) for i in torch.randperm(4)]
        fac = [float(torch.empty(1).uniform_(lo, hi)) for lo, hi in ((max(0.0, 1 - b), 1 + b), (max(0.0, 1 - c), 1 + c),
                                                                     (max(0.0, 1 - s), 1 + s), (-h, h))]
        jitter = (order, fac[0], fac[1], fac[2], fac[3])
    gray = bool(torch.rand(1) < 0.2) if grayscale else False
    sigma = None
    if blur and np.random.random_sample() < 0.5:
        sigma = (2.0 - 0.1) * np.random.random_sample() + 0.1
    return {"jitter": jitter, "gray": gray, "blur": sigma}


def blur_ksize(h: int, w: int) -> int:
    return int((0.1 * min(w, h) // 2 * 2) + 1)  # base_dataset.py:138-140


def gaussian_taps_q8(ksize: int, sigma: float) -> np.ndarray:
    """The taps OpenCV's uint8 GaussianBlur uses: the double Gaussian scaled to 256, rounded half to even from the outside in
    with the rounding error carried along, the centre tap taking the rest (smooth.dispatch.cpp, the bit-exact path)."""
    x = np.arange(ksize, dtype=np.float64) - (ksize - 1) * 0.5
    k = np.exp(-0.5 * x * x / (sigma * sigma))
    k = k / k.sum()
    res = np.zeros(ksize, dtype=np.int64)
    err, s = 0.0, 0
    for i in range(ksize // 2):
        adj = k[i] * 256.0 + err
        v0 = int(np.rint(adj))
        err = adj - v0
        res[i] = res[ksize - 1 - i] = v0
        s += v0
    res[ksize // 2] = 256 - 2 * s
    return res.astype(np.int32)


def build_photometric_header(draws, H: int, W: int):
    """-> (header int32 [B, 16], taps int32 [B, ksize] or None, ksize) for pp_augment_photometric."""
    hdr = np.zeros((len(draws), _PHDR), dtype=np.int32)
    ks = blur_ksize(H, W)
    any_blur = any(d["blur"] is not None for d in draws) and ks / 2 < min(H, W)
    taps = np.zeros((len(draws), ks), dtype=np.int32) if any_blur else None
    for b, d in enumerate(draws):
        if d["jitter"] is not None:
            order, fb, fc, fs, hue = d["jitter"]
            hdr[b, 0] = 1
            hdr[b, 1:5] = order
            hdr[b, 5:8] = np.array([fb, fc, fs], dtype=np.float32).view(np.int32)  # Blend.c takes a C float
            hdr[b, 8] = int(np.int32(hue * 255).astype(np.uint8))  # torchvision's uint8 hue shift
        hdr[b, 9] = int(bool(d["gray"]))
        if d["blur"] is not None and any_blur:
            hdr[b, 10] = 1
            taps[b] = gaussian_taps_q8(ks, d["blur"])
    return hdr, taps, (ks if any_blur else 0)


class GpuPhotometricAugment:
    """`x = aug(x_u8, draws)`: x_u8 uint8 [B, H, W, 3] on CUDA (the crop the geometric step leaves), draws = one
    `draw_photometric()` per sample -> float32 [B, 3, H, W] normalised (and the uint8 image with return_u8=True)."""

    def __init__(self, mean, std):
        self.mean = torch.tensor(list(mean), dtype=torch.float32)
        self.std = torch.tensor(list(std), dtype=torch.float32)
        self._ws = None

    def __call__(self, x_u8, draws, return_u8=False):
        _lib._need_cuda(x_u8)
        B, H, W, C3 = x_u8.shape
        assert C3 == 3 and x_u8.dtype == torch.uint8 and x_u8.is_contiguous() and len(draws) == B
        hdr, taps, ks = build_photometric_header(draws, H, W)
        dev = x_u8.device
        hdr_d = torch.from_numpy(hdr).to(dev, non_blocking=True)
        taps_d = torch.from_numpy(taps).to(dev, non_blocking=True) if taps is not None else None
        need = _lib.C.c_size_t()
        _lib.check(_lib.lib().pp_augment_photometric_workspace_bytes(B, H, W, _lib.C.byref(need)), "pp_augment_photometric_workspace_bytes")
        if self._ws is None or self._ws.numel() < need.value or self._ws.device != dev:
            self._ws = torch.empty(need.value, dtype=torch.uint8, device=dev)
        out = torch.empty((B, 3, H, W), dtype=torch.float32, device=dev)
        out_u8 = torch.empty_like(x_u8) if return_u8 else None
        _lib.check(_lib.lib().pp_augment_photometric(
            _lib._ptr(x_u8), B, H, W, _lib._ptr(hdr_d), _lib._ptr(taps_d), ks, self.mean.data_ptr(), self.std.data_ptr(),
            _lib._ptr(self._ws), self._ws.numel(), _lib._ptr(out), _lib._ptr(out_u8), _lib._stream(x_u8)), "pp_augment_photometric")
        return (out, out_u8) if return_u8 else out


class GpuAugment:
    """The reference's whole training-sample pipeline (base_dataset.py:174-183) for a raw uint8 batch on the device:
    joint geometric augmentation -> photometric augmentation of the image -> to_tensor + normalize.
    `x, y, queries, labelled = aug(x_u8, y_u8, q_u8, lq_u8)` draws per sample from the streams the reference uses (Python's
    `random` for the geometry, torch / NumPy for the photometric part), or takes the draws as arguments."""

    def __init__(self, crop_size, mean, std, ignore_index, mean_val=None, photometric=True, geometric_flags=None,
                 photometric_flags=None):
        """geometric_flags / photometric_flags: `args.augmentations["geometric" | "photometric"]` (args.py:62-76); default all on"""
        self.geo = GpuGeometricAugment(crop_size, mean, std, ignore_index, mean_val)
        g, f = geometric_flags or {}, photometric_flags or {}
        self.geo_flags = dict(random_scale=g.get("random_scale", True), do_crop=g.get("crop", True), random_hflip=g.get("random_hflip", True))
        self.photo_flags = dict(color_jitter=f.get("random_color_jitter", True), grayscale=f.get("random_grayscale", True),
                                blur=f.get("random_gaussian_blur", True))
        self.photo = GpuPhotometricAugment(mean, std) if photometric and any(self.photo_flags.values()) else None

    def __call__(self, x_u8, y_u8=None, q_u8=None, lq_u8=None, geo_params=None, photo_draws=None):
        B, H, W, _ = x_u8.shape
        if geo_params is None or (self.photo is not None and photo_draws is None):
            geo_params, photo_draws = [], []
            for _ in range(B):  # per sample: geometry first, then the photometric draws (base_dataset.py:175-178)
                geo_params.append(draw_geometric(H, W, self.geo.crop, **self.geo_flags))
                if self.photo is not None:
                    photo_draws.append(draw_photometric(**self.photo_flags))
        if self.photo is None:
            return self.geo(x_u8, y_u8, q_u8, lq_u8, geo_params)
        x_crop, y, q, lq = self.geo(x_u8, y_u8, q_u8, lq_u8, geo_params, as_uint8=True)
        return self.photo(x_crop, photo_draws), y, q, lq
