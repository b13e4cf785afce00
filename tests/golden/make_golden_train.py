"""Golden TRAINING trajectories of the T path from the UNMODIFIED reference modules (build container only):

    python tests/golden/make_golden_train.py      # needs /root/reference  ->  tests/golden/train_golden.npz

For both networks - DeepLab(MobileNetV2) (networks/deeplab.py) and the RN50-DeepLabv3+ composition of reference modules
(ResNetBackbone('resnet50_dilated8') -> ASPP('resnet', 8) -> low_level_conv -> SegmentHead, make_golden_model.py) - at the
Cityscapes crop the benchmark runs (256x512, 19 classes; 1 % of the pixels labelled, batch 4):

  * weights  = oracle.deeplab_oracle.reference_init_state_dict(shapes, seed): the reference's own initial DISTRIBUTIONS
               drawn from NumPy's host-independent RandomState (torch's seeded CPU normal stream differs between hosts),
               i.e. the well-conditioned network a real run starts from.  Not stored: both sides rebuild them.
  * K = 3 optimisation steps of model.py:103-122 with the optimiser the reference builds for `cs`
               (utils/utils.py:112-141: Adam, encoder lr/10, torch default betas / eps), Dropout p forced to 0
               (GPU Philox != CPU Mersenne twister), BatchNorm in train mode.
  Stored: the K losses, every parameter's gradient norm at step 0 (+ a few full gradients), and - after the K steps, in eval
  mode - the 1/4-resolution logits of image 0, the full-resolution argmax map of a fresh input and its margin scores'
  checksum (what the query round consumes).
"""
import os
import sys
from argparse import Namespace

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden_model as mg  # noqa: E402  (imports the reference modules, stubs the pretrained-weight download)
from oracle.deeplab_oracle import reference_init_state_dict  # noqa: E402

sys.path.insert(0, "/root/reference")
from utils.utils import get_optimizer as ref_get_optimizer  # noqa: E402

OUT = os.path.join(HERE, "train_golden.npz")
K_STEPS = 3
C, H, W = 19, 256, 512
FULL_GRADS = ["seg_head.classifier.weight", "seg_head.classifier.bias", "seg_head.segment_head.5.weight", "aspp.bn1.bias",
              "low_level_conv.0.weight", "aspp.aspp3.bn.weight", "aspp.global_avg_pool.1.weight"]
CFG = {"mobilenet": dict(B=4, seed=5), "resnet": dict(B=4, seed=6)}
# 1 % of the pixels labelled (1310 / image) instead of the benchmark's 10: the loss and the gradients then average over
# ~5000 pixels, so the comparison measures the arithmetic instead of the noise of a 40-pixel mean on a random-init network
N_LAB = 1310


def batch(B, seed):
    """host-independent inputs (NumPy RandomState): x ~ N(0, 1), y ~ U{0..18} with 1 % void, N_LAB labelled px / image."""
    rs = np.random.RandomState(seed)
    x = torch.from_numpy(rs.standard_normal((B, 3, H, W)).astype(np.float32))
    y = rs.randint(0, C, size=(B, H, W)).astype(np.int64)
    y[rs.rand(B, H, W) < 0.01] = C
    q = np.zeros((B, H * W), dtype=bool)
    for i in range(B):
        q[i, rs.choice(H * W, N_LAB, replace=False)] = True
    return x, torch.from_numpy(y), torch.from_numpy(q.reshape(B, H, W))


def ref_model(backbone):
    m = mg.RefDeepLab(mg.ARGS) if backbone == "mobilenet" else mg.RefRN50DeepLab()
    for mod in m.modules():
        if isinstance(mod, nn.Dropout):
            mod.p = 0.0
    return m


def forward_pred(m, x, backbone):
    if backbone == "mobilenet":
        return m(x)["pred"]
    return m(x)[0]


def lowres_logits(m, x, backbone):
    """the 1/4-resolution head output (before deeplab.py:55's final upsample)."""
    if backbone == "resnet":
        return m(x)[1]
    feats = {}
    h = m.seg_head.register_forward_hook(lambda mod, inp, out: feats.__setitem__("pred", out["pred"]))
    try:
        m(x)
    finally:
        h.remove()
    return feats["pred"]


def main():
    out = {"k_steps": np.array(K_STEPS)}
    for backbone, cfg in CFG.items():
        m = ref_model(backbone)
        shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
        m.load_state_dict(reference_init_state_dict(shapes, seed=cfg["seed"]))
        opt_args = Namespace(dataset_name="cs", network_name="deeplab", optimizer_type="Adam",
                             optimizer_params={"lr": 5e-4, "betas": (0.9, 0.999), "weight_decay": 2e-4, "eps": 1e-7})
        opt = ref_get_optimizer(opt_args, m)
        m.train()
        losses = []
        for k in range(K_STEPS):
            x, y, q = batch(cfg["B"], 100 * cfg["seed"] + k)
            y = y.clone()
            y.flatten()[~q.flatten()] = C  # model.py:108-110
            loss = F.cross_entropy(forward_pred(m, x, backbone), y, ignore_index=C)  # model.py:116
            opt.zero_grad()
            loss.backward()
            if k == 0:
                named = dict(m.named_parameters())
                names = sorted(named)
                out[f"{backbone}_grad_names"] = np.array(names)
                out[f"{backbone}_grad_norms"] = np.array([named[n].grad.norm().item() for n in names])
                for n in FULL_GRADS:
                    out[f"{backbone}_grad::{n}"] = named[n].grad.numpy().copy()
            opt.step()
            losses.append(loss.item())
            print(backbone, "step", k, "loss", loss.item())
        out[f"{backbone}_losses"] = np.array(losses)
        m.eval()
        x, _, _ = batch(2, 100 * cfg["seed"] + 50)
        with torch.no_grad():
            lr = lowres_logits(m, x, backbone)
            pred = F.interpolate(lr, size=(H, W), mode="bilinear", align_corners=True)
        out[f"{backbone}_eval_lowres0"] = lr[0].numpy()
        out[f"{backbone}_eval_argmax"] = pred.argmax(1).numpy().astype(np.uint8)
        top2 = F.softmax(pred, dim=1).topk(2, dim=1).values
        out[f"{backbone}_eval_margin_mean"] = np.array((top2[:, 0] - top2[:, 1]).abs().double().mean().item())
        sd = m.state_dict()
        out[f"{backbone}_final_param_norm"] = np.array(sum(float(v.double().pow(2).sum()) for k_, v in sd.items()
                                                        if v.dtype.is_floating_point) ** 0.5)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
