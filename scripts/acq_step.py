"""GPU (plain or under ncu): a few steps of the acquisition-kernel pipeline of bench.py (256 images, logits in HBM)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
K = int(sys.argv[2]) if len(sys.argv) > 2 else 3
r = bench.bench_acq_pipeline(B, K, 3, torch.device("cuda:0"), 1, 0, overlap=False)
print(r)
