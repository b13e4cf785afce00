"""Sparse labelled-pixel cross entropy — host side of `pp_sparse_ce` (reference: model.py:108-116).

The reference overwrites every unlabelled target with `ignore_index` and runs a dense
`F.cross_entropy(logits_fullres, y, ignore_index)`.  Here the labelled pixels are turned into a short
(img, flat index, label) list and ONE kernel evaluates the x4 bilinear upsample (align_corners=True,
deeplab.py:55), log-softmax, NLL and the gradient w.r.t. the 1/4-resolution logits at those pixels only.
"""
import torch

from . import _lib


def labelled_pixel_list(y, queries, ignore_index):
    """(px_img, px_idx, px_label) int32 device tensors of the pixels that contribute to the loss:
    queries != 0 (model.py:109) and y != ignore_index (F.cross_entropy's ignore_index)."""
    B = y.shape[0]
    yf = y.reshape(B, -1)
    keep = yf != ignore_index
    if queries is not None:
        keep = keep & queries.reshape(B, -1).to(torch.bool)
    nz = keep.nonzero(as_tuple=False)  # row-major: image, then flat index (one host sync for the size)
    px_img = nz[:, 0].to(torch.int32).contiguous()
    px_idx = nz[:, 1].to(torch.int32).contiguous()
    px_label = yf[keep].to(torch.int32).contiguous()
    return px_img, px_idx, px_label


class _SparseCE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits_lowres, size, px_img, px_idx, px_label, n_valid=None):
        loss, grad, pred = _lib.sparse_ce(logits_lowres, size, px_img, px_idx, px_label, want_grad=True, want_pred=True,
                                          n_valid=n_valid)
        ctx.save_for_backward(grad)
        ctx.in_dtype = logits_lowres.dtype
        ctx.mark_non_differentiable(pred)
        return loss.reshape(()), pred

    @staticmethod
    def backward(ctx, g_loss, _g_pred):
        (grad,) = ctx.saved_tensors
        return (grad * g_loss).to(ctx.in_dtype), None, None, None, None, None


class LabelCapacityError(_lib.PixelPickError):
    """more labelled pixels in a batch than the fixed-capacity list of a captured step can hold"""


def labelled_pixel_list_host(y, queries, ignore_index, capacity=None, n_classes=None):
    """Host-side (NumPy) version for CPU batches straight from the dataloader: no device sync, optional padding to a
    fixed `capacity` (CUDA graphs).  Returns (px_img, px_idx, px_label, n_valid) as int32 CPU tensors.
    Labels are validated like F.cross_entropy does (model.py:116): a target that is neither `ignore_index` nor in
    [0, n_classes) raises instead of silently training on a wrong class."""
    import numpy as np
    yn = y.numpy() if isinstance(y, torch.Tensor) else np.asarray(y)
    B = yn.shape[0]
    yf = yn.reshape(B, -1)
    keep = yf != ignore_index
    if queries is not None:
        qn = queries.numpy() if isinstance(queries, torch.Tensor) else np.asarray(queries)
        keep &= qn.reshape(B, -1).astype(bool)
    img, idx = np.nonzero(keep)
    lab = yf[img, idx]
    n = img.size
    if n and (lab.min() < 0 or (n_classes is not None and lab.max() >= n_classes)):
        bad = lab[(lab < 0) | (lab >= (n_classes if n_classes is not None else np.iinfo(np.int64).max))]
        raise IndexError(f"Target {int(bad[0])} is out of bounds.")  # the message of F.cross_entropy on the CPU
    cap = n if capacity is None else capacity
    if n > cap:
        raise LabelCapacityError(f"{n} labelled pixels exceed the captured capacity {cap}")
    out = [np.zeros(cap, dtype=np.int32) for _ in range(3)]
    out[0][:n], out[1][:n], out[2][:n] = img, idx, lab
    return tuple(torch.from_numpy(a) for a in out) + (torch.tensor([n], dtype=torch.int32),)


def sparse_cross_entropy(logits_lowres, y, queries, ignore_index, size=None, return_pred=False, px=None, n_valid=None):
    """== F.cross_entropy(F.interpolate(logits_lowres, size, 'bilinear', align_corners=True), y*, ignore_index)
    with y* = y where queries else ignore_index.  NaN when no pixel is labelled (as the reference).
    px = (px_img, px_idx, px_label) device lists may be given instead of (y, queries) (+ n_valid: device int32 [1])."""
    if px is None:
        size = tuple(y.shape[-2:]) if size is None else tuple(size)
        px = labelled_pixel_list(y, queries, ignore_index)
    loss, pred = _SparseCE.apply(logits_lowres, tuple(size), px[0], px[1], px[2], n_valid)
    if return_pred:
        return loss, pred, px
    return loss
