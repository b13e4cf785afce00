"""GPU: time of the device input pipeline for one Cityscapes-shape batch (B = 32, 256x512): geometric only vs geometric +
photometric (pp_augment_geometric(_u8) + pp_augment_photometric), draws and tables included."""
import os, sys, time, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from pixelpick_b200.augment import GpuAugment, draw_geometric, draw_photometric

dev = torch.device("cuda:0")
B, H, W, crop = 32, 256, 512, (256, 512)
MEAN, STD = [0.28689554, 0.32513303, 0.28389177], [0.18696375, 0.19017339, 0.18720214]
x = torch.randint(0, 256, (B, H, W, 3), dtype=torch.uint8, device=dev)
y = torch.randint(0, 20, (B, H, W), dtype=torch.uint8, device=dev)
q = (torch.rand((B, H, W), device=dev) < 0.01).to(torch.uint8) * 255
for photometric in (False, True):
    aug = GpuAugment(crop, MEAN, STD, 19, photometric=photometric)
    random.seed(0); torch.manual_seed(0); np.random.seed(0)
    for _ in range(3):
        aug(x, y, q, q)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 10
    for _ in range(n):
        aug(x, y, q, q)
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / n * 1e3
    geo = [draw_geometric(H, W, crop) for _ in range(B)]
    ph = [draw_photometric() for _ in range(B)]
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        aug(x, y, q, q, geo, ph)
    b.record(); torch.cuda.synchronize()
    print(f"photometric={photometric}: {wall:.2f} ms / batch of {B} wall (draws + tables + kernels); device+host with fixed draws {a.elapsed_time(b) / n:.2f} ms")
