"""pixelpick_b200 — B200-native hot paths of PixelPick (query/acquisition and the DeepLabv3+ train step)
behind the reference's own Python interfaces.  The compute lives in `csrc/` (hand-written sm_100a CUDA
behind the C-ABI of `include/pixelpick_b200.h`); the modules here mirror the reference's classes."""
__version__ = "0.1.0"
