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
