"""GPU debug: head forward/backward in isolation vs the CPU oracle head (identical features, dense labels)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from argparse import Namespace
import numpy as np, torch, torch.nn as nn, torch.nn.functional as F
from oracle import deeplab_oracle as orc
from pixelpick_b200.deeplab import DeepLab, _HeadFn
from pixelpick_b200.loss import sparse_cross_entropy

DEV = torch.device("cuda:0")
A = Namespace(use_mc_dropout=False, mc_dropout_p=0.2, n_classes=19)
m = DeepLab(A)
sd = orc.synthetic_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()}, seed=1)
m.load_state_dict(sd); m.to(DEV).train()
for mod in m.modules():
    if isinstance(mod, nn.Dropout): mod.p = 0.0
B, H, W = 2, int(os.environ.get("HH", 128)), int(os.environ.get("WW", 256))
g = torch.Generator().manual_seed(0)
high = torch.randn((B, 320, H // 16, W // 16), generator=g).abs()
low = torch.randn((B, 24, H // 4, W // 4), generator=g)
y = torch.randint(0, 19, (B, H, W), generator=g)
nlab = int(sys.argv[1]) if len(sys.argv) > 1 else 0
q = torch.ones((B, H, W), dtype=torch.bool)
if nlab:
    q = torch.zeros((B, H * W), dtype=torch.bool)
    for i in range(B): q[i, torch.randperm(H * W, generator=g)[:nlab]] = True
    q = q.view(B, H, W)
# oracle
sdr = {k: (v.clone().requires_grad_(True) if v.dtype.is_floating_point and not k.endswith(("running_mean", "running_var")) else v) for k, v in sd.items()}
hr, lr_ = high.clone().requires_grad_(True), low.clone().requires_grad_(True)
c = orc._Ctx(sdr, True)
pred_lr, _ = orc.head(c, hr, lr_, [1, 6, 12, 18])
pred = F.interpolate(pred_lr, size=(H, W), mode="bilinear", align_corners=True)
ref_loss = orc.sparse_ce_loss(pred, y, q, 19); ref_loss.backward()
# ours
hg, lg = high.to(DEV).requires_grad_(True), low.to(DEV).requires_grad_(True)
pre = m._pooled_branch(hg)
m._rng_step = torch.zeros(1, dtype=torch.int64, device=DEV)
out = _HeadFn.apply(m, 1, hg, lg, pre, *m._head_params())
loss = sparse_cross_entropy(out, y.to(DEV), q.to(DEV), 19); loss.backward()
print("loss", loss.item(), ref_loss.item(), "lowres rel", ((out.detach().cpu() - pred_lr.detach()).abs().max() / pred_lr.abs().max()).item())
def cmp(name, a, b):
    a, b = a.float().cpu().flatten(), b.flatten()
    print(f"  {name:40s} cos {torch.dot(a, b).item() / (a.norm().item() * b.norm().item() + 1e-30):.4f} ratio {a.norm().item() / (b.norm().item() + 1e-30):.3f}")
cmp("d_high", hg.grad, hr.grad); cmp("d_low", lg.grad, lr_.grad)
P = dict(m.named_parameters())
for n in ["seg_head.classifier.weight", "seg_head.segment_head.5.weight", "seg_head.segment_head.4.weight", "seg_head.segment_head.1.bias",
          "seg_head.segment_head.0.weight", "low_level_conv.0.weight", "low_level_conv.1.weight", "aspp.bn1.weight", "aspp.bn1.bias",
          "aspp.conv1.weight", "aspp.aspp1.atrous_conv.weight", "aspp.aspp1.bn.weight", "aspp.aspp2.atrous_conv.weight",
          "aspp.aspp3.atrous_conv.weight", "aspp.aspp4.atrous_conv.weight", "aspp.aspp4.bn.bias", "aspp.global_avg_pool.1.weight",
          "aspp.global_avg_pool.2.weight"]:
    cmp(n, P[n].grad, sdr[n].grad)
w = P["aspp.conv1.weight"].grad.float().cpu(); r = sdr["aspp.conv1.weight"].grad
cmp("aspp.conv1.weight[:, :1024]", w[:, :1024], r[:, :1024]); cmp("aspp.conv1.weight[:, 1024:]", w[:, 1024:], r[:, 1024:])
