"""Launches the SegmentHead conv (fwd, dgrad) and wgrad kernels once each at the bench shape, for ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pixelpick_b200 import _lib
dev = torch.device("cuda:0")
B = 32
x = torch.randn((B, 64, 128, 320), device=dev).to(torch.bfloat16)
wt = torch.randn((256, 304, 3, 3), device=dev) * 0.02
w = _lib.pack_conv_weight(wt, 320, 256)
wd = _lib.pack_conv_weight(wt, 256, 320, True)
sc, sf = torch.ones(256, device=dev), torch.zeros(256, device=dev)
for _ in range(2):
    y = _lib.conv_igemm(x, w, 256, scale=sc, shift=sf, relu=True)
    dx = _lib.conv_igemm(y, wd, 320)
    dw = _lib.conv_wgrad(x, 304, y, 256, 9)
torch.cuda.synchronize()
print("ok", y.shape, dx.shape, dw.shape)
# ASPP (aspp.py:49-52) at the ResNet-50 OS8 shape: 1x1 and dilated 3x3 branch, forward + weight gradient, and the fused
# four-branch data gradient (pp_conv_igemm_multi)
xa = torch.randn((B, 32, 64, 2048), device=dev).to(torch.bfloat16)
wa1 = _lib.pack_conv_weight(torch.randn((256, 2048, 1, 1), device=dev) * 0.02, 2048, 256)
wa3 = _lib.pack_conv_weight(torch.randn((256, 2048, 3, 3), device=dev) * 0.02, 2048, 256)
cat = torch.empty((B, 32, 64, 1024), dtype=torch.bfloat16, device=dev)
for _ in range(2):
    _lib.conv_igemm(xa, wa1, 256, dil=1, scale=sc, shift=sf, relu=True, out=cat, c_off=0)
    _lib.conv_igemm(xa, wa3, 256, dil=12, scale=sc, shift=sf, relu=True, out=cat, c_off=256)
    dwa = _lib.conv_wgrad(xa, 2048, cat[..., 256:512], 256, 9, 12)
torch.cuda.synchronize()
print("aspp ok", cat.shape, dwa.shape)
