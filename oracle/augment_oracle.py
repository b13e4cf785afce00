"""ORACLE (test infrastructure, NOT product code) - NumPy restatement of the reference's joint geometric augmentation
(datasets/base_dataset.py:48-127: random scale -> pad -> random crop -> horizontal flip, applied jointly to the image, the
label map, the query mask and the human-label map) including the resampling arithmetic of the libraries it calls:

  * image   : PIL `Image.resize(BILINEAR)` (Pillow src/libImaging/Resample.c): separable triangle filter whose support grows
              with the down-scaling factor, coefficients normalised in double and rounded to 22-bit fixed point, a horizontal
              pass into a uint8 intermediate, then a vertical pass;
  * label   : PIL NEAREST = source index (int) of a running sum a/2 + a + a + ... with a = in / out (Geometry.c affine scaling,
              pixel centres);
  * queries : torch `F.interpolate(mode="nearest")` = min(floor(dst * float32(in / out)), in - 1) (no pixel centres) - the
              reference resizes the mask TENSORS with torchvision's tensor path, so masks and labels use different conventions.

Pinned by tests/golden/make_golden_augment.py -> augment_golden.npz (outputs of the unmodified reference method) in
tests/test_augment_oracle_golden.py.  Only tests/ and bench.py's CPU baseline may import this."""
import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2


def pil_bilinear_coeffs(in_size, out_size):
    """(xmin[out], n[out], k[out][ksize] int32 fixed point) of Resample.c:precompute_coeffs + normalize_coeffs_8bpc."""
    scale = filterscale = in_size / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = 1.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    xmin = np.zeros(out_size, dtype=np.int32)
    cnt = np.zeros(out_size, dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        lo = int(center - support + 0.5)
        lo = max(lo, 0)
        hi = int(center + support + 0.5)
        hi = min(hi, in_size)
        n = hi - lo
        w = []
        for x in range(n):
            t = (x + lo - center + 0.5) * ss
            t = -t if t < 0.0 else t
            w.append(1.0 - t if t < 1.0 else 0.0)
        ww = sum(w)  # left-to-right double sum, as the C loop
        acc = 0.0
        for v in w:
            acc += v
        ww = acc
        for x in range(n):
            v = w[x] / ww if ww != 0.0 else w[x]
            kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        xmin[xx], cnt[xx] = lo, n
    return xmin, cnt, kk


def _clip8(v):
    return np.clip(v >> PRECISION_BITS, 0, 255).astype(np.uint8)


def pil_resize_bilinear(img, out_hw):
    """img uint8 [H, W, C] -> uint8 [h, w, C] exactly as PIL (horizontal pass first, uint8 intermediate)."""
    H, W, _ = img.shape
    h, w = out_hw
    cur = img
    if w != W:
        xmin, cnt, kk = pil_bilinear_coeffs(W, w)
        out = np.empty((H, w, img.shape[2]), dtype=np.uint8)
        for xx in range(w):
            acc = np.full((H, img.shape[2]), 1 << (PRECISION_BITS - 1), dtype=np.int64)
            for x in range(cnt[xx]):
                acc += cur[:, xmin[xx] + x, :].astype(np.int64) * int(kk[xx, x])
            out[:, xx, :] = _clip8(acc)
        cur = out
    if h != H:
        ymin, cnt, kk = pil_bilinear_coeffs(H, h)
        out = np.empty((h, cur.shape[1], img.shape[2]), dtype=np.uint8)
        for yy in range(h):
            acc = np.full((cur.shape[1], img.shape[2]), 1 << (PRECISION_BITS - 1), dtype=np.int64)
            for y in range(cnt[yy]):
                acc += cur[ymin[yy] + y, :, :].astype(np.int64) * int(kk[yy, y])
            out[yy] = _clip8(acc)
        cur = out
    return cur


def pil_nearest_index(in_size, out_size):
    """Pillow Geometry.c:ImagingScaleAffine: xo = a / 2, then xo += a per output pixel (a running double sum, not a * (x + 0.5)),
    source index = (int) xo."""
    a = in_size / out_size
    idx = np.empty(out_size, dtype=np.int64)
    xo = a * 0.5
    for i in range(out_size):
        idx[i] = int(xo)
        xo += a
    return idx.clip(0, in_size - 1)


def torch_nearest_index(in_size, out_size):
    scale = np.float32(in_size) / np.float32(out_size)
    idx = np.floor(np.arange(out_size, dtype=np.float32) * scale).astype(np.int64)
    return np.minimum(idx, in_size - 1)


def geometric(x, y, q, lq, scale, crop, start, flip, mean_val, ignore_index):
    """base_dataset.py:48-127 with the draws given: scale (uniform(0.5, 2.0)), start = (randint, randint), flip (random() > 0.5).
    x uint8 [H, W, 3], y uint8 [H, W], q uint8 [H, W] (0 / 255), lq uint8 [H, W] -> the four crops (q as 0 / 1)."""
    H, W = y.shape
    w_rs, h_rs = int(W * scale), int(H * scale)
    xr = pil_resize_bilinear(x, (h_rs, w_rs))
    iy, ix = pil_nearest_index(H, h_rs), pil_nearest_index(W, w_rs)
    yr = y[iy][:, ix]
    ty, tx = torch_nearest_index(H, h_rs), torch_nearest_index(W, w_rs)
    qr, lqr = q[ty][:, tx], lq[ty][:, tx]
    ph, pw = max(crop[0] - h_rs, 0), max(crop[1] - w_rs, 0)
    xp = np.empty((h_rs + ph, w_rs + pw, 3), dtype=np.uint8)
    xp[...] = np.asarray(mean_val, dtype=np.uint8)
    xp[:h_rs, :w_rs] = xr
    yp = np.full((h_rs + ph, w_rs + pw), ignore_index, dtype=np.uint8)
    yp[:h_rs, :w_rs] = yr
    qp = np.zeros((h_rs + ph, w_rs + pw), dtype=np.uint8)
    qp[:h_rs, :w_rs] = qr
    lqp = np.full((h_rs + ph, w_rs + pw), ignore_index, dtype=np.uint8)
    lqp[:h_rs, :w_rs] = lqr
    sh, sw = start
    sl = (slice(sh, sh + crop[0]), slice(sw, sw + crop[1]))
    outs = [xp[sl], yp[sl], qp[sl], lqp[sl]]
    if flip:
        outs = [o[:, ::-1] for o in outs]
    outs[2] = outs[2] // 255
    return [np.ascontiguousarray(o) for o in outs]
