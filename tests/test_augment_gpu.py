"""GPU: pp_augment_geometric (pixelpick_b200/augment.py) against the golden outputs of the UNMODIFIED reference method
BaseDataset._geometric_augmentations (tests/golden/augment_golden.npz) with the reference's random draws injected, and a full
Cityscapes-shape batch against the oracle.  Label map and masks: bit-exact.  Image: the uint8 crop PIL produced, pushed through
to_tensor + normalize in float32 - equal to the kernel's output bit for bit."""
import os
import random

import numpy as np
import pytest
import torch

from oracle import augment_oracle as orc
from pixelpick_b200.augment import GpuGeometricAugment, draw_geometric

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "augment_golden.npz"))
MEAN, STD, MEAN_VAL = [0.28689554, 0.32513303, 0.28389177], [0.18696375, 0.19017339, 0.18720214], (73, 83, 72)
DEV = torch.device("cuda:0")


def _normalised(x_u8_hwc):
    t = torch.from_numpy(np.ascontiguousarray(x_u8_hwc)).permute(2, 0, 1).float().div(255)  # TF.to_tensor
    return t.sub(torch.tensor(MEAN)[:, None, None]).div(torch.tensor(STD)[:, None, None])   # TF.normalize


def test_draws_follow_the_reference_order():
    for ci in range(int(GOLD["n_cases"])):
        h, w, ch, cw, seed, h_rs, w_rs, sh, sw, flip = [int(v) for v in GOLD[f"c{ci}_cfg"]]
        random.seed(seed)
        scale, a, b, f = draw_geometric(h, w, (ch, cw))
        assert scale == float(GOLD[f"c{ci}_scale"]) and (a, b, int(f)) == (sh, sw, flip)


@pytest.mark.gpu
@pytest.mark.parametrize("ci", range(int(GOLD["n_cases"])))
def test_kernel_equals_the_reference_method(ci):
    h, w, ch, cw, seed, h_rs, w_rs, sh, sw, flip = [int(v) for v in GOLD[f"c{ci}_cfg"]]
    aug = GpuGeometricAugment((ch, cw), MEAN, STD, ignore_index=19, mean_val=MEAN_VAL)
    to = lambda k: torch.from_numpy(GOLD[f"c{ci}_{k}"]).unsqueeze(0).contiguous().to(DEV)
    x, y, q, lq = aug(to("x"), to("y"), to("q"), to("lq"), [(float(GOLD[f"c{ci}_scale"]), sh, sw, bool(flip))])
    torch.cuda.synchronize()
    assert np.array_equal(y[0].cpu().numpy(), GOLD[f"c{ci}_yo"])
    assert np.array_equal(q[0].cpu().numpy(), GOLD[f"c{ci}_qo"])
    assert np.array_equal(lq[0].cpu().numpy(), GOLD[f"c{ci}_lqo"])
    want = _normalised(GOLD[f"c{ci}_xo"])
    assert torch.equal(x[0].cpu(), want), float((x[0].cpu() - want).abs().max())


@pytest.mark.gpu
def test_cityscapes_batch_equals_the_oracle():
    B, H, W, crop = 4, 256, 512, (256, 512)
    rs = np.random.RandomState(5)
    x = rs.randint(0, 256, size=(B, H, W, 3)).astype(np.uint8)
    y = rs.randint(0, 20, size=(B, H, W)).astype(np.uint8)
    q = ((rs.rand(B, H, W) < 0.01) * 255).astype(np.uint8)
    random.seed(9)
    params = [draw_geometric(H, W, crop) for _ in range(B)]
    params[0] = (0.5, 0, 0, True)   # the strongest down-scale (5-tap filters, maximal padding) ...
    params[1] = (2.0, 200, 300, False)  # ... and the strongest up-scale
    aug = GpuGeometricAugment(crop, MEAN, STD, ignore_index=19, mean_val=MEAN_VAL)
    xo, yo, qo, _ = aug(torch.from_numpy(x).to(DEV), torch.from_numpy(y).to(DEV), torch.from_numpy(q).to(DEV), None, params)
    torch.cuda.synchronize()
    for b, (scale, sh, sw, flip) in enumerate(params):
        wx, wy, wq, _ = orc.geometric(x[b], y[b], q[b], q[b], scale, crop, (sh, sw), flip, MEAN_VAL, 19)
        assert np.array_equal(yo[b].cpu().numpy(), wy) and np.array_equal(qo[b].cpu().numpy(), wq)
        assert torch.equal(xo[b].cpu(), _normalised(wx)), b


# ---- photometric augmentation (pp_augment_photometric, base_dataset.py:129-141) ----
PGOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "photometric_golden.npz"))


def _photo_input(h, w, seed):  # tests/golden/make_golden_photometric.py:make_input
    rs = np.random.RandomState(1000 + seed)
    x = rs.randint(0, 256, size=(h, w, 3)).astype(np.uint8)
    x = ((x.astype(np.float32) + np.roll(x, 1, 0) + np.roll(x, 1, 1)) / 3).astype(np.uint8)
    x[0, :, :] = rs.randint(0, 256, size=(w, 3))
    return x


@pytest.mark.gpu
def test_photometric_kernel_equals_the_reference_method():
    """the reference's _photometric_augmentations output under seeded streams (goldens) == draws replayed by
    augment.draw_photometric + pp_augment_photometric, bit for bit; images of one size go through as ONE batch"""
    from pixelpick_b200.augment import GpuPhotometricAugment, draw_photometric
    aug = GpuPhotometricAugment(MEAN, STD)
    by_size = {}
    for ci in range(int(PGOLD["n_cases"])):
        h, w, seed = [int(v) for v in PGOLD[f"c{ci}_cfg"]]
        torch.manual_seed(seed)
        np.random.seed(seed)
        by_size.setdefault((h, w), []).append((ci, _photo_input(h, w, seed), draw_photometric()))
    for (h, w), items in by_size.items():
        x = torch.from_numpy(np.stack([it[1] for it in items])).to(DEV)
        out, out_u8 = aug(x, [it[2] for it in items], return_u8=True)
        torch.cuda.synchronize()
        for b, (ci, _, draw) in enumerate(items):
            want = PGOLD[f"c{ci}_out"]
            got = out_u8[b].cpu().numpy()
            assert np.array_equal(got, want), (ci, draw, int((got != want).sum()))
            assert torch.equal(out[b].cpu(), _normalised(want)), ci


@pytest.mark.gpu
def test_hue_step_on_every_colour():
    """all 2^24 colours through RGB -> HSV -> shift -> RGB on the device == the oracle (itself equal to Pillow on every colour)"""
    from pixelpick_b200.augment import GpuPhotometricAugment
    aug = GpuPhotometricAugment(MEAN, STD)
    rr, gg, bb = np.meshgrid(np.arange(256), np.arange(256), np.arange(256), indexing="ij")
    cube = np.stack([rr, gg, bb], -1).astype(np.uint8).reshape(1, 4096, 4096, 3)
    x = torch.from_numpy(cube).to(DEV)
    for hue in (0.0, 0.137, -0.2):
        # brightness / contrast / saturation with factor 1.0 copy the image (Image.blend's alpha == 1 shortcut)
        draw = {"jitter": ([0, 3, 1, 2], 1.0, 1.0, 1.0, hue), "gray": False, "blur": None}
        _, got = aug(x, [draw], return_u8=True)
        torch.cuda.synchronize()
        want = orc.adjust_hue(cube[0], hue)
        assert np.array_equal(got[0].cpu().numpy(), want), hue


@pytest.mark.gpu
def test_full_training_sample_pipeline_equals_the_oracle():
    """base_dataset.py:174-183 on a Cityscapes-shape batch: geometric (uint8 crop) -> photometric -> normalise"""
    from pixelpick_b200.augment import GpuAugment, draw_photometric
    B, H, W, crop = 6, 256, 512, (256, 512)
    rs = np.random.RandomState(7)
    x = rs.randint(0, 256, size=(B, H, W, 3)).astype(np.uint8)
    x = ((x.astype(np.float32) + np.roll(x, 1, 1) + np.roll(x, 1, 2)) / 3).astype(np.uint8)
    y = rs.randint(0, 20, size=(B, H, W)).astype(np.uint8)
    random.seed(3)
    torch.manual_seed(3)
    np.random.seed(3)
    geo = [draw_geometric(H, W, crop) for _ in range(B)]
    photo = [draw_photometric() for _ in range(B)]
    photo[0] = {"jitter": ([1, 3, 0, 2], 1.7, 0.3, 1.8, -0.19), "gray": False, "blur": 1.9}   # everything on
    photo[1] = {"jitter": None, "gray": True, "blur": 0.1}                                      # sigma 0.1: an identity blur
    photo[2] = {"jitter": ([2, 1, 3, 0], 0.2, 1.8, 0.2, 0.2), "gray": True, "blur": None}
    aug = GpuAugment(crop, MEAN, STD, ignore_index=19, mean_val=MEAN_VAL)
    xo, yo, _, _ = aug(torch.from_numpy(x).to(DEV), torch.from_numpy(y).to(DEV), None, None, geo, photo)
    torch.cuda.synchronize()
    for b in range(B):
        scale, sh, sw, flip = geo[b]
        wx, wy, _, _ = orc.geometric(x[b], y[b], y[b], y[b], scale, crop, (sh, sw), flip, MEAN_VAL, 19)
        wx = orc.photometric_oracle(wx, photo[b])
        assert np.array_equal(yo[b].cpu().numpy(), wy)
        assert torch.equal(xo[b].cpu(), _normalised(wx)), (b, photo[b])
