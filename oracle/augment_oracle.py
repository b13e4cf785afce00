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


# =====================================================================================================================
# Photometric augmentation (datasets/base_dataset.py:129-141, GaussianBlur :192-210): NumPy restatement of what the
# libraries the reference calls compute on uint8 images - torchvision's PIL path (ImageEnhance -> Pillow's Blend.c, the
# "L" conversion, Convert.c's RGB <-> HSV) and OpenCV's bit-exact uint8 GaussianBlur.  Pinned three ways:
#   * tests/golden/make_golden_photometric.py -> photometric_golden.npz: outputs of the UNMODIFIED reference method
#     `BaseDataset._photometric_augmentations` under seeded torch / NumPy streams (tests/test_augment_oracle_golden.py);
#   * against the live Pillow over all 2^24 colours (RGB -> HSV and HSV -> RGB) and against live cv2 for random sigmas, when
#     those libraries are importable (same test file);
#   * the product's draw function must consume the random streams exactly as torchvision / the reference do.
# =====================================================================================================================
def blend_u8(deg, img, alpha):
    """PIL.Image.blend(deg, img, alpha) on uint8 arrays (Blend.c): fp32 in1 + alpha * (in2 - in1), truncated; clipped when
    alpha is outside [0, 1]; alpha == 0 / 1 copy an operand."""
    a = np.float32(alpha)
    if a == 0.0:
        return deg.copy()
    if a == 1.0:
        return img.copy()
    d, i = deg.astype(np.int32), img.astype(np.int32)
    t = d.astype(np.float32) + a * (i - d).astype(np.float32)
    if 0.0 <= a <= 1.0:
        return t.astype(np.int32).astype(np.uint8)
    return np.where(t <= 0, 0, np.where(t >= 255, 255, t.astype(np.int32))).astype(np.uint8)


def rgb_to_l(x):
    """PIL convert("L") (Convert.c, ITU-R 601-2 in 16-bit fixed point)."""
    r, g, b = [x[..., k].astype(np.int64) for k in range(3)]
    return ((r * 19595 + g * 38470 + b * 7471 + 0x8000) >> 16).astype(np.uint8)


def rgb2hsv_u8(x):
    """Convert.c:rgb2hsv_row: fp32 quotients; `2.0 + rc - bc` and `fmod(h / 6.0 + 1.0, 1.0)` in double, each rounded to fp32
    on assignment; then (int)(h * 255.0) in double."""
    r, g, b = [x[..., k].astype(np.int32) for k in range(3)]
    maxc = np.maximum(r, np.maximum(g, b))
    minc = np.minimum(r, np.minimum(g, b))
    gray = maxc == minc
    f32, f64 = np.float32, np.float64
    cr = (maxc - minc).astype(f32)
    crs = np.where(gray, f32(1), cr)
    s = cr / np.where(maxc == 0, 1, maxc).astype(f32)
    rc, gc, bc = [(maxc - c).astype(f32) / crs for c in (r, g, b)]
    h = np.where(r == maxc, bc - gc, np.where(g == maxc, (2.0 + rc.astype(f64) - bc.astype(f64)).astype(f32),
                                              (4.0 + gc.astype(f64) - rc.astype(f64)).astype(f32))).astype(f32)
    h = np.fmod(h.astype(f64) / 6.0 + 1.0, 1.0).astype(f32)
    uh = np.clip((h.astype(f64) * 255.0).astype(np.int32), 0, 255)
    us = np.clip((s.astype(f64) * 255.0).astype(np.int32), 0, 255)
    return np.stack([np.where(gray, 0, uh), np.where(gray, 0, us), maxc], -1).astype(np.uint8)


def hsv2rgb_u8(x):
    """Convert.c:hsv2rgb: sector / remainder from (float)h * 6.0 / 255.0 in double, f and s / 255 rounded to fp32, the three
    candidates as round((float)v * (1.0 - ...)) in double (half away from zero)."""
    h, s, v = [x[..., k].astype(np.int32) for k in range(3)]
    f64, f32 = np.float64, np.float32
    h6 = h.astype(f64) * 6.0 / 255.0
    i = np.floor(h6).astype(np.int32)
    f = (h6 - i).astype(f32).astype(f64)
    fs = (s.astype(f64) / 255.0).astype(f32).astype(f64)
    vd = v.astype(f64)
    rnd = lambda a: np.clip(np.floor(a + 0.5).astype(np.int32), 0, 255)
    p, q, t = rnd(vd * (1.0 - fs)), rnd(vd * (1.0 - fs * f)), rnd(vd * (1.0 - fs * (1.0 - f)))
    m = i % 6
    r = np.choose(m, [v, q, p, p, t, v])
    g = np.choose(m, [t, v, v, q, p, p])
    b = np.choose(m, [p, p, t, v, v, q])
    z = s == 0
    return np.stack([np.where(z, v, r), np.where(z, v, g), np.where(z, v, b)], -1).astype(np.uint8)


def adjust_brightness(x, f):
    return blend_u8(np.zeros_like(x), x, f)  # ImageEnhance.Brightness: degenerate = black


def adjust_contrast(x, f):
    l = rgb_to_l(x)
    mean = int(float(l.astype(np.int64).sum()) / l.size + 0.5)  # int(ImageStat.Stat(L).mean[0] + 0.5)
    return blend_u8(np.full_like(x, mean), x, f)


def adjust_saturation(x, f):
    return blend_u8(np.repeat(rgb_to_l(x)[..., None], 3, -1), x, f)  # ImageEnhance.Color: degenerate = L image


def hue_shift_u8(hue_factor):
    """torchvision _functional_pil.adjust_hue: np.int32(hue_factor * 255).astype(np.uint8)"""
    return int(np.int32(hue_factor * 255).astype(np.uint8))


def adjust_hue(x, hue_factor):
    hsv = rgb2hsv_u8(x)
    hsv[..., 0] = (hsv[..., 0].astype(np.int32) + hue_shift_u8(hue_factor)).astype(np.uint8)  # wraps, as the uint8 +=
    return hsv2rgb_u8(hsv)


def gaussian_taps_q8(ksize, sigma):
    """OpenCV's bit-exact Gaussian taps for uint8 images (smooth.dispatch.cpp: getGaussianKernelBitExact +
    getGaussianKernelFixedPoint_ED): the double kernel exp(-x^2 / (2 sigma^2)) / sum scaled by 256, rounded half to even from
    the outside in with the rounding error carried to the next tap, the centre tap taking what is left of 256."""
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
    return res


def gaussian_blur_u8(x, ksize, sigma):
    """cv2.GaussianBlur(x, (ksize, ksize), sigma) for uint8 HWC: separable, Q8.8 after the horizontal pass, Q16.16 after the
    vertical, (v + 2^15) >> 16, BORDER_REFLECT_101."""
    k = gaussian_taps_q8(ksize, sigma)
    r = ksize // 2
    H, W = x.shape[:2]
    p = np.pad(x.astype(np.int64), ((r, r), (r, r), (0, 0)), mode="reflect")
    h = np.zeros((H + 2 * r, W, x.shape[2]), dtype=np.int64)
    for t in range(ksize):
        h += k[t] * p[:, t:t + W]
    v = np.zeros(x.shape, dtype=np.int64)
    for t in range(ksize):
        v += k[t] * h[t:t + H]
    return ((v + (1 << 15)) >> 16).astype(np.uint8)


def blur_ksize(h, w):
    return int((0.1 * min(w, h) // 2 * 2) + 1)  # base_dataset.py:138-140


def photometric_oracle(x, draw):
    """BaseDataset._photometric_augmentations on a uint8 HWC image with the draws given:
    draw = {"jitter": None | (order[4], brightness, contrast, saturation, hue), "gray": bool, "blur": None | sigma}."""
    x = np.ascontiguousarray(x, dtype=np.uint8)
    if draw["jitter"] is not None:
        order, b, c, s, hue = draw["jitter"]
        for fn in order:
            x = (adjust_brightness(x, b) if fn == 0 else adjust_contrast(x, c) if fn == 1 else
                 adjust_saturation(x, s) if fn == 2 else adjust_hue(x, hue))
    if draw["gray"]:
        x = np.repeat(rgb_to_l(x)[..., None], 3, -1)
    if draw["blur"] is not None:
        x = gaussian_blur_u8(x, blur_ksize(x.shape[0], x.shape[1]), draw["blur"])
    return x
