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
