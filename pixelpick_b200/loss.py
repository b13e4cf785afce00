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
    def forward(ctx, logits_lowres, size, px_img, px_idx, px_label):
        loss, grad, pred = _lib.sparse_ce(logits_lowres, size, px_img, px_idx, px_label, want_grad=True, want_pred=True)
        ctx.save_for_backward(grad)
        ctx.in_dtype = logits_lowres.dtype
        ctx.mark_non_differentiable(pred)
        return loss.reshape(()), pred

    @staticmethod
    def backward(ctx, g_loss, _g_pred):
        (grad,) = ctx.saved_tensors
        return (grad * g_loss).to(ctx.in_dtype), None, None, None, None


def sparse_cross_entropy(logits_lowres, y, queries, ignore_index, size=None, return_pred=False):
    """== F.cross_entropy(F.interpolate(logits_lowres, size, 'bilinear', align_corners=True), y*, ignore_index)
    with y* = y where queries else ignore_index.  NaN when no pixel is labelled (as the reference)."""
    size = tuple(y.shape[-2:]) if size is None else tuple(size)
    px = labelled_pixel_list(y, queries, ignore_index)
    loss, pred = _SparseCE.apply(logits_lowres, size, *px)
    if return_pred:
        return loss, pred, px
    return loss
