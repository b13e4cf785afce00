"""CPU: the oracle's TRAINING path (oracle.deeplab_oracle.train_steps: forward in train mode, sparse CE, the `cs` Adam
groups, BatchNorm running statistics) against the trajectory the UNMODIFIED reference modules produced
(tests/golden/make_golden_train.py -> train_golden.npz), for MobileNetV2-DeepLab and the RN50-DeepLabv3+ composition at
the Cityscapes crop (256x512).  This pins what the `-m gpu` parity tests compare the CUDA path with."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import deeplab_oracle as orc

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
GOLD = np.load(os.path.join(HERE, "golden", "train_golden.npz"))
C, H, W = 19, 256, 512
CFG = {"mobilenet": dict(B=4, seed=5), "resnet": dict(B=4, seed=6)}
# 1 % of the pixels labelled (1310 / image) instead of the benchmark's 10: the loss and the gradients then average over
# ~5000 pixels, so the comparison measures the arithmetic instead of the noise of a 40-pixel mean on a random-init network
N_LAB = 1310


def golden_batch(B, seed):
    """the generator's inputs, restated (importing make_golden_train would import the reference)."""
    rs = np.random.RandomState(seed)
    x = torch.from_numpy(rs.standard_normal((B, 3, H, W)).astype(np.float32))
    y = rs.randint(0, C, size=(B, H, W)).astype(np.int64)
    y[rs.rand(B, H, W) < 0.01] = C
    q = np.zeros((B, H * W), dtype=bool)
    for i in range(B):
        q[i, rs.choice(H * W, N_LAB, replace=False)] = True
    return x, torch.from_numpy(y), torch.from_numpy(q.reshape(B, H, W))


def init_state(backbone):
    from argparse import Namespace
    from pixelpick_b200.deeplab import DeepLab  # parameter names / shapes only (constructed on the CPU, never run)
    m = DeepLab(Namespace(use_mc_dropout=False, mc_dropout_p=0.2, n_classes=C), backbone=backbone)
    return orc.reference_init_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()}, seed=CFG[backbone]["seed"])


@pytest.mark.parametrize("backbone", ["mobilenet", "resnet"])
def test_oracle_training_trajectory_equals_the_reference(backbone):
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    cfg = CFG[backbone]
    k_steps = int(GOLD["k_steps"])
    sd0 = init_state(backbone)
    batches = [golden_batch(cfg["B"], 100 * cfg["seed"] + k) for k in range(k_steps)]
    sd, losses, grads0 = orc.train_steps(sd0, batches, backbone, C)
    # losses: fp32 CPU kernels on possibly another host (oneDNN picks its kernels per ISA): 2e-4 relative over 3 steps
    assert np.allclose(losses, GOLD[f"{backbone}_losses"], rtol=2e-4, atol=1e-5), (losses, GOLD[f"{backbone}_losses"])
    names = [str(n) for n in GOLD[f"{backbone}_grad_names"]]
    # the reference registers MobileNetV2.features twice (low_level_features / high_level_features): same tensors
    canon = {n: n for n in names}
    for n in names:
        for alias in ("backbone.low_level_features.", "backbone.high_level_features."):
            if n.startswith(alias):
                idx, rest = n[len(alias):].split(".", 1)
                canon[n] = f"backbone.features.{int(idx)}.{rest}"
    got = np.array([grads0[canon[n]].norm().item() for n in names])
    want = GOLD[f"{backbone}_grad_norms"]
    assert np.allclose(got, want, rtol=2e-3, atol=1e-7), np.abs(got / np.maximum(want, 1e-12) - 1).max()
    for key in GOLD.files:
        if key.startswith(f"{backbone}_grad::"):
            n = key.split("::", 1)[1]
            g, r = grads0[n].numpy(), GOLD[key]
            assert np.allclose(g, r, rtol=1e-3, atol=2e-4 * np.abs(r).max()), (n, np.abs(g - r).max(), np.abs(r).max())
    # eval mode after the K steps (running statistics + updated weights): what the query round sees
    x, _, _ = golden_batch(2, 100 * cfg["seed"] + 50)
    with torch.no_grad():
        out = orc.deeplab_forward(sd, x, backbone=backbone)
    ref_lr = GOLD[f"{backbone}_eval_lowres0"]
    scale = np.abs(ref_lr).max()
    assert np.abs(out["lowres"][0].numpy() - ref_lr).max() < 2e-3 * scale
    agree = (out["pred"].argmax(1).numpy() == GOLD[f"{backbone}_eval_argmax"]).mean()
    assert agree > 0.999, agree
    top2 = F.softmax(out["pred"], dim=1).topk(2, dim=1).values
    margin = (top2[:, 0] - top2[:, 1]).abs().double().mean().item()
    assert abs(margin - float(GOLD[f"{backbone}_eval_margin_mean"])) < 1e-3 * float(GOLD[f"{backbone}_eval_margin_mean"]) + 1e-6
    norm = sum(float(v.double().pow(2).sum()) for k, v in sd.items()
               if v.dtype.is_floating_point and not k.startswith(("backbone.low_level_features.", "backbone.high_level_features."))) ** 0.5
    # (the reference's state_dict counts MobileNetV2's aliased tensors twice; compare the RN50 norm only)
    if backbone == "resnet":
        assert abs(norm - float(GOLD["resnet_final_param_norm"])) < 1e-5 * norm
