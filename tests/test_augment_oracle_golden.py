"""CPU: the augmentation oracle (oracle/augment_oracle.py: PIL's fixed-point bilinear resample, PIL / torch nearest, pad, crop,
flip) against the outputs of the UNMODIFIED reference method BaseDataset._geometric_augmentations
(tests/golden/make_golden_augment.py -> augment_golden.npz): bit-exact for the image, the label map and both masks."""
import os

import numpy as np
import pytest

from oracle import augment_oracle as aug

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "augment_golden.npz"))


@pytest.mark.parametrize("ci", range(int(GOLD["n_cases"])))
def test_geometric_augmentation_equals_the_reference(ci):
    h, w, ch, cw, seed, h_rs, w_rs, sh, sw, flip = [int(v) for v in GOLD[f"c{ci}_cfg"]]
    scale = float(GOLD[f"c{ci}_scale"])
    assert (int(h * scale), int(w * scale)) == (h_rs, w_rs)
    xo, yo, qo, lqo = aug.geometric(GOLD[f"c{ci}_x"], GOLD[f"c{ci}_y"], GOLD[f"c{ci}_q"], GOLD[f"c{ci}_lq"], scale, (ch, cw),
                                    (sh, sw), bool(flip), (73, 83, 72), 19)
    assert np.array_equal(yo, GOLD[f"c{ci}_yo"])
    assert np.array_equal(qo, GOLD[f"c{ci}_qo"])
    assert np.array_equal(lqo, GOLD[f"c{ci}_lqo"])
    d = np.abs(xo.astype(int) - GOLD[f"c{ci}_xo"].astype(int))
    assert d.max() == 0, (d.max(), (d > 0).mean())


# ---- photometric augmentation (base_dataset.py:129-141) ----
PGOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "photometric_golden.npz"))


def _photo_input(h, w, seed):  # tests/golden/make_golden_photometric.py:make_input
    rs = np.random.RandomState(1000 + seed)
    x = rs.randint(0, 256, size=(h, w, 3)).astype(np.uint8)
    x = ((x.astype(np.float32) + np.roll(x, 1, 0) + np.roll(x, 1, 1)) / 3).astype(np.uint8)
    x[0, :, :] = rs.randint(0, 256, size=(w, 3))
    return x


def test_photometric_augmentation_equals_the_reference_under_the_same_seeds():
    """the product's draw function consumes torch's / NumPy's global streams exactly as torchvision and the reference's
    GaussianBlur do, and the oracle's arithmetic turns those draws into the reference method's output, bit for bit"""
    import torch
    from pixelpick_b200.augment import draw_photometric
    seen = {"jitter": 0, "gray": 0, "blur": 0, "orders": set()}
    for ci in range(int(PGOLD["n_cases"])):
        h, w, seed = [int(v) for v in PGOLD[f"c{ci}_cfg"]]
        torch.manual_seed(seed)
        np.random.seed(seed)
        draw = draw_photometric()
        got = aug.photometric_oracle(_photo_input(h, w, seed), draw)
        assert np.array_equal(got, PGOLD[f"c{ci}_out"]), (ci, draw)
        seen["jitter"] += draw["jitter"] is not None
        seen["gray"] += bool(draw["gray"])
        seen["blur"] += draw["blur"] is not None
        if draw["jitter"] is not None:
            seen["orders"].add(tuple(draw["jitter"][0]))
    n = int(PGOLD["n_cases"])
    assert 0 < seen["jitter"] < n and 0 < seen["gray"] < n and 0 < seen["blur"] < n and len(seen["orders"]) >= 8, seen


def test_hsv_conversions_equal_pillow_on_every_colour():
    Image = pytest.importorskip("PIL.Image")
    for r0 in range(0, 256, 32):
        rr, gg, bb = np.meshgrid(np.arange(r0, r0 + 32), np.arange(256), np.arange(256), indexing="ij")
        x = np.stack([rr, gg, bb], -1).astype(np.uint8).reshape(32 * 256, 256, 3)
        assert np.array_equal(aug.rgb2hsv_u8(x), np.asarray(Image.fromarray(x, "RGB").convert("HSV")))
        assert np.array_equal(aug.hsv2rgb_u8(x), np.asarray(Image.fromarray(x, "HSV").convert("RGB")))


def test_enhance_steps_equal_torchvision_and_blur_equals_opencv():
    TF = pytest.importorskip("torchvision.transforms.functional")
    cv2 = pytest.importorskip("cv2")
    from PIL import Image
    rs = np.random.RandomState(3)
    x = rs.randint(0, 256, size=(37, 53, 3)).astype(np.uint8)
    x[0, :, 0] = np.arange(53) * 4
    pil = Image.fromarray(x)
    for f in list(rs.uniform(0.2, 1.8, size=40)) + [0.0, 1.0, 0.5, 2.0]:
        f = float(np.float32(f))
        assert np.array_equal(aug.adjust_brightness(x, f), np.asarray(TF.adjust_brightness(pil, f)))
        assert np.array_equal(aug.adjust_contrast(x, f), np.asarray(TF.adjust_contrast(pil, f)))
        assert np.array_equal(aug.adjust_saturation(x, f), np.asarray(TF.adjust_saturation(pil, f)))
    for hue in list(rs.uniform(-0.2, 0.2, size=20)) + [-0.5, 0.5, 0.0]:
        assert np.array_equal(aug.adjust_hue(x, float(hue)), np.asarray(TF.adjust_hue(pil, float(hue))))
    for ks in (5, 9, 25):
        for sigma in rs.uniform(0.1, 2.0, size=12):
            assert np.array_equal(aug.gaussian_blur_u8(x, ks, float(sigma)), cv2.GaussianBlur(x, (ks, ks), float(sigma)))
