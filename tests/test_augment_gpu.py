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
