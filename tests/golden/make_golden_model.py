"""Golden vectors for the T path from the UNMODIFIED reference modules (build container only):

    python tests/golden/make_golden_model.py      # needs /root/reference

Weights are NOT stored (23 MB): both sides rebuild them with oracle.deeplab_oracle.synthetic_state_dict
from the parameter names/shapes.  Stored: eval logits of DeepLab(MobileNetV2) on a 48x64 input, train-mode
(dropout p forced to 0 at run time, BatchNorm batch statistics) loss / per-parameter gradient norms / a few
full gradients / updated running stats on a 2x3x64x64 batch, and eval logits of the RN50-DeepLabv3+ composition
(ResNetBackbone('resnet50_dilated8') -> ASPP('resnet', 8) -> low_level_conv(256->48) -> SegmentHead)."""
import os
import sys
from argparse import Namespace

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.dont_write_bytecode = True
sys.path.insert(0, "/root/reference")
sys.path.insert(1, ROOT)
import networks.mobilenet_v2 as ref_mnv2  # noqa: E402

ref_mnv2.MobileNetV2._load_pretrained_model = lambda self: None  # no network: seeded synthetic weights instead
from networks.deeplab import DeepLab as RefDeepLab  # noqa: E402
from networks.aspp import ASPP as RefASPP  # noqa: E402
from networks.decoders import SegmentHead as RefSegmentHead  # noqa: E402
from networks.backbones.resnet_backbone import ResNetBackbone  # noqa: E402
from oracle.deeplab_oracle import synthetic_state_dict  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "model_golden.npz")
ARGS = Namespace(use_mc_dropout=False, mc_dropout_p=0.2, n_classes=19)


def inputs(seed, shape):
    return torch.randn(shape, generator=torch.Generator().manual_seed(seed))


class RefRN50DeepLab(nn.Module):
    """The composition of reference modules SURVEY.md fact 1 describes (no such class exists in the reference)."""

    def __init__(self):
        super().__init__()
        self.backbone = ResNetBackbone("resnet50_dilated8")
        self.aspp = RefASPP("resnet", 8, nn.BatchNorm2d)
        self.low_level_conv = nn.Sequential(nn.Conv2d(256, 48, 1, bias=False), nn.BatchNorm2d(48), nn.ReLU())
        self.seg_head = RefSegmentHead(ARGS)

    def forward(self, x):
        c2, _, _, c5 = self.backbone(x)
        y = self.aspp(c5)
        ll = self.low_level_conv(c2)
        y = F.interpolate(y, size=ll.shape[2:], mode="bilinear", align_corners=True)
        out = self.seg_head(torch.cat((y, ll), dim=1))
        return F.interpolate(out["pred"], size=x.shape[2:], mode="bilinear", align_corners=True), out["pred"]


def main():
    out = {}
    # ---------------- MobileNetV2-DeepLab, eval ----------------
    m = RefDeepLab(ARGS)
    shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    sd = synthetic_state_dict(shapes, seed=1)
    m.load_state_dict(sd)
    m.eval()
    x = inputs(10, (1, 3, 48, 64))
    with torch.no_grad():
        o = m(x)
    out["mnv2_eval_pred"] = o["pred"].numpy()
    # ---------------- MobileNetV2-DeepLab, one train step (dropout off) ----------------
    m.load_state_dict(sd)
    m.train()
    for mod in m.modules():
        if isinstance(mod, nn.Dropout):
            mod.p = 0.0
    x = inputs(11, (2, 3, 64, 64))
    rs = np.random.RandomState(11)
    y = torch.from_numpy(rs.randint(0, 19, size=(2, 64, 64)).astype(np.int64))
    q = torch.zeros((2, 64 * 64), dtype=torch.bool)
    for i in range(2):
        q[i, torch.from_numpy(rs.choice(64 * 64, 10, replace=False))] = True
    q = q.view(2, 64, 64)
    yy = y.clone()
    yy.flatten()[~q.flatten()] = 19  # model.py:108-110
    pred = m(x)["pred"]
    loss = F.cross_entropy(pred, yy, ignore_index=19)  # model.py:116
    loss.backward()
    out["mnv2_train_pred"] = pred.detach().numpy()
    out["mnv2_train_loss"] = np.array(loss.item())
    names = [n for n, _ in m.named_parameters()]
    out["mnv2_train_grad_names"] = np.array(names)
    out["mnv2_train_grad_norms"] = np.array([p.grad.norm().item() for _, p in m.named_parameters()])
    for n in ["seg_head.classifier.weight", "seg_head.classifier.bias", "seg_head.segment_head.5.weight",
              "aspp.bn1.bias", "low_level_conv.0.weight", "aspp.aspp3.bn.weight", "aspp.global_avg_pool.1.weight"]:
        out["mnv2_train_grad::" + n] = dict(m.named_parameters())[n].grad.numpy()
    g = dict(m.named_parameters())["seg_head.segment_head.4.weight"].grad
    out["mnv2_train_grad::seg_head.segment_head.4.weight[:8]"] = g[:8].numpy()
    g = dict(m.named_parameters())["aspp.aspp2.atrous_conv.weight"].grad
    out["mnv2_train_grad::aspp.aspp2.atrous_conv.weight[:4]"] = g[:4].numpy()
    out["mnv2_train_running_mean::aspp.bn1"] = m.aspp.bn1.running_mean.numpy()
    out["mnv2_train_running_var::seg_head.segment_head.1"] = m.seg_head.segment_head[1].running_var.numpy()
    # ---------------- RN50-DeepLabv3+ composition, eval ----------------
    r = RefRN50DeepLab()
    shapes = {k: tuple(v.shape) for k, v in r.state_dict().items()}
    r.load_state_dict(synthetic_state_dict(shapes, seed=2))
    r.eval()
    x = inputs(12, (1, 3, 64, 96))
    with torch.no_grad():
        pred, lowres = r(x)
    out["rn50_eval_pred"] = pred.numpy()
    out["rn50_n_params"] = np.array(sum(p.numel() for p in r.parameters()))
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
