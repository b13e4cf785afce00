"""ctypes binding of `libpixelpick_b200.so` (the C-ABI declared in include/pixelpick_b200.h).

There is NO fallback: if the library is missing or a call fails, an exception is raised.  PyTorch is
used only as the owner of device memory and streams — the wrappers below unwrap `data_ptr()`,
shapes/strides and `torch.cuda.current_stream().cuda_stream` and pass plain pointers.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libpixelpick_b200.so")

PP_F32, PP_BF16 = 0, 1
STRATEGIES = {"entropy": 0, "least_confidence": 1, "margin_sampling": 2}
# fill value for excluded pixels and top-k direction (query.py:50,57-61,198)
FILL = {"entropy": 0.0, "least_confidence": 0.0, "margin_sampling": 1.0, "random": 1.0}
UPSAMPLED_SCORE_CLASSES = (11, 19, 21)  # class counts pp_acq_score_upsampled / pp_eval_confusion_upsampled are instantiated for
LARGEST = {"entropy": True, "least_confidence": True, "margin_sampling": False, "random": False}

_lib = None


class PixelPickError(RuntimeError):
    pass


_vp, _i, _i64, _sz, _f = C.c_void_p, C.c_int, C.c_int64, C.c_size_t, C.c_float

_SIGNATURES = {
    "pp_version": ([], _i),
    "pp_last_error": ([], C.c_char_p),
    "pp_launch_count": ([], C.c_longlong),
    "pp_host_ord_key": ([_f, _i], C.c_uint),
    "pp_host_bucket0": ([_f, _i], C.c_uint),
    "pp_device_info": ([C.POINTER(_i), C.POINTER(_i), C.POINTER(_i), C.c_char_p, _i], _i),
    "pp_acq_score": ([_vp, _i, _i, _i, _i, _i, _i64, _i64, _i64, _vp, _vp, _vp, _i, _vp, _vp, _vp], _i),
    "pp_acq_score_select": ([_vp, _i, _i, _i, _i, _i, _i64, _i64, _i64, _vp, _vp, _vp, _i, _i, _vp, _vp, _sz, _vp], _i),
    "pp_acq_score_upsampled": ([_vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _i, _vp, _vp, _vp], _i),
    "pp_acq_topk_workspace_bytes": ([_i, _i, _i, C.POINTER(_sz)], _i),
    "pp_acq_topk_prepare": ([_vp, _sz, _i, _i, _i, _vp], _i),
    "pp_acq_topk_hist0": ([_vp], _vp),
    "pp_acq_topk": ([_vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _sz, _vp], _i),
    "pp_acq_select": ([_vp, _i, _i, _i, _i, _i, _vp, _sz, _vp], _i),
    "pp_acq_pick": ([_vp, _sz, _i, _i, _i, _vp, _i, _vp, _vp], _i),
    "pp_acq_gather": ([_vp, _i, _i, _vp, _i, _vp, _vp], _i),
    "pp_acq_entropy_at": ([_vp, _i, _i, _i, _i, _i, _i64, _i64, _i64, _vp, _i, _vp, _vp], _i),
    "pp_acq_entropy_at_upsampled": ([_vp, _i, _i, _i, _i, _i, _i, _vp, _i, _vp, _vp], _i),
    "pp_query_stats_at": ([_vp, _i, _i, _i, _i, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp], _i),
    "pp_acq_session_create": ([C.POINTER(_vp), _i, _i, _i, _i, _i, _i], _i),
    "pp_acq_session_destroy": ([_vp], _i),
    "pp_acq_session_run_host": ([_vp, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp], _i),
    "pp_acq_session_begin_host": ([_vp, _vp, _vp, _vp, _i, _i], _i),
    "pp_acq_session_finish_host": ([_vp, _vp, _vp], _i),
    "pp_sparse_ce": ([_vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _i, _vp, _f, _vp, _vp, _vp, _vp], _i),
    "pp_eval_confusion_upsampled": ([_vp, _i, _i, _i, _i, _i, _i, _vp, _i, _vp, _vp, _vp], _i),
    "pp_metrics_accumulate": ([_vp, _vp, _vp, _i, _i, _vp, _vp, _vp, _vp, _vp], _i),
    "pp_upsample_bilinear_ac": ([_vp, _i, _i, _i, _i, _vp, _i, _i, _vp], _i),
    "pp_upsample_bilinear_ac_bwd": ([_vp, _i, _i, _i, _i, _vp, _i, _i, _vp], _i),
    "pp_conv_wgrad": ([_vp, _i, _i, _vp, _i, _i, _i, _i, _i, _i, _i, _vp, _i, _i, _vp], _i),
    "pp_bn_stats": ([_vp, _i64, _i, _i, _i, _vp, _vp], _i),
    "pp_bn_finalize": ([_vp, _i, _i64, _f, _f, _vp, _vp, _vp, _vp, _vp, _i, _vp], _i),
    "pp_bn_apply": ([_vp, _i64, _i, _i, _i, _vp, _vp, _i, _f, C.c_uint64, C.c_uint64, _vp, _vp, _i, _i, _vp], _i),
    "pp_bn_bwd": ([_vp, _i, _i, _vp, _i, _i, _i64, _i, _vp, _vp, _vp, _vp, _i, _f, C.c_uint64, C.c_uint64, _vp, _vp,
                   _vp, _vp], _i),
    "pp_bn_apply_stats": ([_vp, _i64, _i, _i, _i, _vp, _vp, _vp, _f, _f, _vp, _vp, _vp, _i, _vp, _i, _vp, _i, _i, _vp], _i),
    "pp_bn_apply_res": ([_vp, _i64, _i, _i, _i, _vp, _vp, _i, _f, C.c_uint64, C.c_uint64, _vp, _vp, _i, _vp, _i, _i, _vp], _i),
    "pp_bn_bwd_res": ([_vp, _i, _i, _vp, _i, _i, _i64, _i, _vp, _vp, _vp, _vp, _i, _f, C.c_uint64, C.c_uint64, _vp, _vp, _i,
                       _vp, _vp, _vp, _i, _i, _vp], _i),
    "pp_bn_scratch_bytes": ([_i, C.POINTER(_sz)], _i),
    "pp_bn_fwd_fused": ([_vp, _i64, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _f, _f, _i, _f, C.c_uint64, C.c_uint64, _vp, _vp, _i,
                         _vp, _i, _i, _vp, _vp, _vp], _i),
    "pp_bn_bwd_fused": ([_vp, _i, _i, _vp, _i, _i, _i64, _i, _vp, _vp, _vp, _vp, _i, _f, C.c_uint64, C.c_uint64, _vp, _vp, _i,
                         _vp, _vp, _vp, _i, _i, _vp, _vp], _i),
    "pp_conv_igemm_multi": ([_vp, _i, _i, _i, _i, _i, _i, _vp, _i, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp, _i, _vp, _i, _vp, _i,
                             _i, _i, _i, _vp], _i),
    "pp_conv_igemm_stats": ([_vp, _i, _i, _i, _i, _i, _i, _vp, _i, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp, _i, _vp, _i, _vp, _i,
                             _i, _i, _i, _vp, _vp], _i),
    "pp_conv_set_epilogue": ([_i], _i),
    "pp_conv_wgrad_multi": ([_vp, _i, _i, _i, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _i, _i, _i, _vp], _i),
    "pp_maxpool3x3s2_fwd": ([_vp, _i, _i, _i, _i, _vp, _vp, _vp], _i),
    "pp_maxpool3x3s2_bwd": ([_vp, _vp, _i, _i, _i, _i, _vp, _vp], _i),
    "pp_adam_step_multi": ([_i, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp], _i),
    "pp_dwconv3x3_fwd": ([_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp], _i),
    "pp_dwconv3x3_fwd_bnact": ([_vp, _vp, _vp, _vp, _i, _vp, _i, _i, _i, _i, _i, _i, _vp], _i),
    "pp_dwconv3x3_dgrad": ([_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp], _i),
    "pp_dwconv3x3_wgrad": ([_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp], _i),
    "pp_upsample_nhwc_bf16": ([_vp, _i, _i, _i, _i, _i, _vp, _i, _i, _i, _i, _vp], _i),
    "pp_upsample_nhwc_bf16_bwd": ([_vp, _i, _i, _i, _i, _i, _i, _vp, _i, _i, _vp], _i),
    "pp_pack_conv_weight": ([_vp, _i, _i, _i, _i, _vp, _i, _i, _vp, _i, _i, _vp], _i),
    "pp_pack_conv_weights_batched": ([_vp, _i, _i, _vp], _i),
    "pp_pack_conv_weights_tiles": ([_i, _i, _i, _i, _i], _i),
    "pp_augment_geometric": ([_vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp, _i, _i, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp], _i),
    "pp_augment_geometric_u8": ([_vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp, _i, _i, _vp, _i, _vp, _vp, _vp, _vp, _vp], _i),
    "pp_augment_photometric_workspace_bytes": ([_i, _i, _i, C.POINTER(C.c_size_t)], _i),
    "pp_augment_photometric": ([_vp, _i, _i, _i, _vp, _vp, _i, _vp, _vp, _vp, C.c_size_t, _vp, _vp, _vp], _i),
    "pp_to_nhwc_bf16": ([_vp, _i, _i64, _i64, _i64, _i64, _i, _i, _i, _i, _vp, _i, _i, _vp], _i),
    "pp_conv_igemm": ([_vp, _i, _i, _i, _i, _i, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _i, _vp, _i, _i, _i, _i, _vp], _i),
}


def exported_symbols():
    """Names include/pixelpick_b200.h declares (kept in sync by tests/test_abi.py)."""
    return sorted(_SIGNATURES)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise PixelPickError(
                f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU / PyTorch fallback for the hot paths)")
        l = C.CDLL(LIB_PATH)
        for name, (argtypes, restype) in _SIGNATURES.items():
            fn = getattr(l, name)
            fn.argtypes = argtypes
            fn.restype = restype
        _lib = l
    return _lib


def check(rc, what):
    if rc != 0:
        raise PixelPickError(f"{what} failed ({rc}): {lib().pp_last_error().decode()}")


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream(t):
    return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _dtype_code(t):
    if t.dtype == torch.float32:
        return PP_F32
    if t.dtype == torch.bfloat16:
        return PP_BF16
    raise PixelPickError(f"unsupported logits dtype {t.dtype}")


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise PixelPickError("pixelpick_b200 kernels need CUDA tensors (no CPU fallback)")


def _mask(t, n, H, W):
    if t is None:
        return None
    if t.dtype == torch.bool:
        t = t.view(torch.uint8)
    if t.dtype != torch.uint8:
        raise PixelPickError("masks must be bool/uint8")
    t = t.reshape(n, H, W)
    return t if t.is_contiguous() else t.contiguous()


def _check_score_outputs(what, n, H, W, out, hist0_ws):
    """the kernel writes n*H*W floats into `out` and n level-0 histograms into the workspace: reject anything smaller."""
    if out is not None and (not out.is_cuda or out.dtype != torch.float32 or not out.is_contiguous() or out.numel() != n * H * W):
        raise PixelPickError(f"{what}: out must be a contiguous CUDA float32 tensor of {n}x{H}x{W} elements, got "
                             f"{out.dtype} {tuple(out.shape)}")
    if hist0_ws is not None and (hist0_ws.n_img < n or hist0_ws.HW != H * W):
        raise PixelPickError(f"{what}: workspace was sized for {hist0_ws.n_img} images of {hist0_ws.HW} pixels, "
                             f"the batch has {n} of {H * W}")


class TopKWorkspace:
    """Device workspace of the radix-select/sort, reusable across batches of the same shape."""

    def __init__(self, n_img, HW, k, device):
        sz = _sz()
        check(lib().pp_acq_topk_workspace_bytes(n_img, HW, k, C.byref(sz)), "pp_acq_topk_workspace_bytes")
        self.n_img, self.HW, self.k = n_img, HW, k
        self.nbytes = sz.value
        self.buf = torch.empty(self.nbytes, dtype=torch.uint8, device=device)
        assert self.buf.data_ptr() % 256 == 0
        # What is known about the zeroed region (in stream order): None = nothing; "all" = all zero (after prepare());
        # n = zero for batches of n images - a completed select zeroes what it consumed (include/pixelpick_b200.h:
        # pp_acq_topk_prepare) but leaves its per-image bucket state behind, and where that lies depends on the batch size.
        # Reset to None by everything that fills the level-0 histogram or starts a select; an exception keeps it None.
        self._clean = None

    def prepare(self, n_img=None, force=False):
        """zero the histogram / counters unless a completed select has left them zero already (always the full-capacity
        region: the regions of smaller batches are prefixes of it)"""
        if self._clean is not None and not force:
            return
        check(lib().pp_acq_topk_prepare(_ptr(self.buf), self.nbytes, self.n_img, self.HW, self.k, _stream(self.buf)),
              "pp_acq_topk_prepare")
        self._clean = "all"

    def _begin_fill(self, n):
        """before a kernel accumulates the level-0 histogram of a batch of n images here: it must start from zero"""
        if self._clean not in ("all", n):
            self.prepare(force=True)
        self._clean = None

    def _begin_select(self, n, hist0_valid):
        if not hist0_valid and self._clean not in ("all", n):
            self.prepare(force=True)  # the select builds the histogram itself, from zero
        self._clean = None

    def hist0_ptr(self):
        return C.c_void_p(lib().pp_acq_topk_hist0(_ptr(self.buf)))


def acq_score(logits, strategy, labelled=None, void_mask=None, keep=None, out=None, hist0_ws=None):
    """score map [n, H, W] float32 of `logits` [n, C, H, W] (any strides with stride_w == 1)."""
    _need_cuda(logits, labelled, void_mask, keep)
    n, Cc, H, W = logits.shape
    if logits.stride(3) != 1:
        logits = logits.contiguous()
    labelled, void_mask, keep = (_mask(m, n, H, W) for m in (labelled, void_mask, keep))
    _check_score_outputs("acq_score", n, H, W, out, hist0_ws)
    if out is None:
        out = torch.empty((n, H, W), dtype=torch.float32, device=logits.device)
    if hist0_ws is not None:
        hist0_ws._begin_fill(n)
    check(lib().pp_acq_score(_ptr(logits), _dtype_code(logits), n, Cc, H, W, logits.stride(0), logits.stride(1),
                             logits.stride(2), _ptr(labelled), _ptr(void_mask), _ptr(keep),
                             STRATEGIES[strategy], _ptr(out),
                             hist0_ws.hist0_ptr() if hist0_ws is not None else None, _stream(logits)),
          "pp_acq_score")
    return out


def acq_score_upsampled(logits_lowres, size, strategy, labelled=None, void_mask=None, keep=None, out=None,
                        hist0_ws=None):
    _need_cuda(logits_lowres, labelled, void_mask, keep)
    n, Cc, h, w = logits_lowres.shape
    H, W = size
    logits_lowres = logits_lowres.float().contiguous()
    labelled, void_mask, keep = (_mask(m, n, H, W) for m in (labelled, void_mask, keep))
    _check_score_outputs("acq_score_upsampled", n, H, W, out, hist0_ws)
    if out is None:
        out = torch.empty((n, H, W), dtype=torch.float32, device=logits_lowres.device)
    if hist0_ws is not None:
        hist0_ws._begin_fill(n)
    check(lib().pp_acq_score_upsampled(_ptr(logits_lowres), n, Cc, h, w, H, W, _ptr(labelled), _ptr(void_mask),
                                       _ptr(keep), STRATEGIES[strategy], _ptr(out),
                                       hist0_ws.hist0_ptr() if hist0_ws is not None else None,
                                       _stream(logits_lowres)),
          "pp_acq_score_upsampled")
    return out


def acq_topk(score_map, k, largest, ws=None, hist0_valid=False, return_values=False):
    """sorted top-k flat indices [n, k] int32 of score_map [n, H*W] (NaN largest, ties -> lower index)."""
    _need_cuda(score_map)
    n = score_map.shape[0]
    sm = score_map.reshape(n, -1)
    if not sm.is_contiguous() or sm.dtype != torch.float32:
        sm = sm.float().contiguous()
    HW = sm.shape[1]
    if ws is None:
        ws = TopKWorkspace(n, HW, k, sm.device)
        ws.prepare()
        hist0_valid = False
    idx = torch.empty((n, k), dtype=torch.int32, device=sm.device)
    val = torch.empty((n, k), dtype=torch.float32, device=sm.device) if return_values else None
    ws._begin_select(n, hist0_valid)
    check(lib().pp_acq_topk(_ptr(sm), n, HW, k, int(bool(largest)), int(bool(hist0_valid)), _ptr(idx), _ptr(val),
                            _ptr(ws.buf), ws.nbytes, _stream(sm)), "pp_acq_topk")
    ws._clean = n  # the select's kernels hand the workspace back zeroed
    return (idx, val) if return_values else idx


def acq_select_pick(score_map, k, largest, pos, n=None, ws=None, hist0_valid=False):
    """flat indices [n_img, n] of the elements at ranks `pos` ([n_img, n] int32, or None for ranks 0..n-1) of the sorted
    top-k list — without sorting it (pp_acq_select + pp_acq_pick).  Equals acq_gather(acq_topk(...), pos)."""
    _need_cuda(score_map)
    n_img = score_map.shape[0]
    sm = score_map.reshape(n_img, -1)
    if not sm.is_contiguous() or sm.dtype != torch.float32:
        sm = sm.float().contiguous()
    HW = sm.shape[1]
    if ws is None:
        ws = TopKWorkspace(n_img, HW, k, sm.device)
        ws.prepare()
        hist0_valid = False
    if pos is not None:
        pos = pos.to(device=sm.device, dtype=torch.int32).contiguous()
        n = pos.shape[1]
    out = torch.empty((n_img, n), dtype=torch.int32, device=sm.device)
    ws._begin_select(n_img, hist0_valid)
    check(lib().pp_acq_select(_ptr(sm), n_img, HW, k, int(bool(largest)), int(bool(hist0_valid)), _ptr(ws.buf), ws.nbytes,
                              _stream(sm)), "pp_acq_select")
    ws._clean = n_img  # the select's kernels hand the workspace back zeroed
    check(lib().pp_acq_pick(_ptr(ws.buf), ws.nbytes, n_img, HW, k, _ptr(pos), n, _ptr(out), _stream(sm)), "pp_acq_pick")
    return out


def acq_score_select_supported(logits, C, H, W):
    """shapes the fused scoring + select kernel covers (see pp_acq_score_select)"""
    if logits.dtype != torch.float32 or W % 4 or C not in UPSAMPLED_SCORE_CLASSES or logits.stride(3) != 1:
        return False
    if any(s % 4 for s in logits.stride()[:3]) or logits.data_ptr() % 16:
        return False
    HW = H * W
    return any(HW % (c * 4096) == 0 and HW // c <= 16384 for c in (1, 2, 4, 8, 16))


def acq_score_select_pick(logits, strategy, k, pos, labelled=None, void_mask=None, keep=None, n=None, ws=None, score_out=None,
                          mark=None):
    """acq_select_pick(acq_score(logits, ...)) in one pass over the logits: flat indices [n_img, n] at ranks `pos` of the sorted
    top-k list; the score map is only written when `score_out` is given.  `ws` must be prepared (TopKWorkspace.prepare)."""
    _need_cuda(logits, labelled, void_mask, keep)
    n_img, Cc, H, W = logits.shape
    HW = H * W
    labelled, void_mask, keep = (_mask(m, n_img, H, W) for m in (labelled, void_mask, keep))
    if ws is None:
        ws = TopKWorkspace(n_img, HW, k, logits.device)
        ws.prepare()
    _check_score_outputs("acq_score_select_pick", n_img, H, W, score_out, ws)
    ws._begin_select(n_img, False)
    check(lib().pp_acq_score_select(_ptr(logits), _dtype_code(logits), n_img, Cc, H, W, logits.stride(0), logits.stride(1),
                                    logits.stride(2), _ptr(labelled), _ptr(void_mask), _ptr(keep), STRATEGIES[strategy], k,
                                    _ptr(score_out), _ptr(ws.buf), ws.nbytes, _stream(logits)), "pp_acq_score_select")
    ws._clean = n_img  # its radix tail hands the workspace back zeroed
    if mark is not None:
        mark.record()  # bench.py: a CUDA event between the fused kernel (+ radix tail) and the pick
    if pos is not None:
        pos = pos.to(device=logits.device, dtype=torch.int32).contiguous()
        n = pos.shape[1]
    out = torch.empty((n_img, n), dtype=torch.int32, device=logits.device)
    check(lib().pp_acq_pick(_ptr(ws.buf), ws.nbytes, n_img, HW, k, _ptr(pos), n, _ptr(out), _stream(logits)), "pp_acq_pick")
    return out


def acq_gather(topk_idx, pos):
    n, k = topk_idx.shape
    nsel = pos.shape[1]
    out = torch.empty((n, nsel), dtype=torch.int32, device=topk_idx.device)
    pos = pos.to(device=topk_idx.device, dtype=torch.int32).contiguous()
    check(lib().pp_acq_gather(_ptr(topk_idx), n, k, _ptr(pos), nsel, _ptr(out), _stream(topk_idx)), "pp_acq_gather")
    return out


def acq_entropy_at(logits, px_idx):
    _need_cuda(logits, px_idx)
    n, Cc, H, W = logits.shape
    if logits.stride(3) != 1:
        logits = logits.contiguous()
    px_idx = px_idx.to(torch.int32).contiguous()
    nsel = px_idx.shape[1]
    out = torch.empty((n, nsel), dtype=torch.float32, device=logits.device)
    check(lib().pp_acq_entropy_at(_ptr(logits), _dtype_code(logits), n, Cc, H, W, logits.stride(0), logits.stride(1),
                                  logits.stride(2), _ptr(px_idx), nsel, _ptr(out), _stream(logits)),
          "pp_acq_entropy_at")
    return out


def acq_entropy_at_upsampled(logits_lowres, size, px_idx):
    _need_cuda(logits_lowres, px_idx)
    n, Cc, h, w = logits_lowres.shape
    H, W = size
    x = logits_lowres.float().contiguous()
    px_idx = px_idx.to(torch.int32).contiguous()
    nsel = px_idx.shape[1]
    out = torch.empty((n, nsel), dtype=torch.float32, device=x.device)
    check(lib().pp_acq_entropy_at_upsampled(_ptr(x), n, Cc, h, w, H, W, _ptr(px_idx), nsel, _ptr(out), _stream(x)),
          "pp_acq_entropy_at_upsampled")
    return out


def query_stats_at(sel_sorted, W, HW, labels=None, n_classes=1, label_hist=None):
    """QueryStats at the picks on the device: sel_sorted int64 [n_img, n] (ascending flat indices), labels uint8 [n_img, HW]
    or None -> (x_coords, y_coords int64 [n_img, n], labels_at int32 | None, n_unique int32 | None, coverage f64 [n_img]);
    label_hist (int64 [n_classes], device) is incremented in place."""
    _need_cuda(sel_sorted, labels, label_hist)
    assert sel_sorted.dtype == torch.int64 and sel_sorted.is_contiguous()
    n_img, n = sel_sorted.shape
    dev = sel_sorted.device
    xs = torch.empty((n_img, n), dtype=torch.int64, device=dev)
    ys = torch.empty((n_img, n), dtype=torch.int64, device=dev)
    cov = torch.empty(n_img, dtype=torch.float64, device=dev)
    lab_at = uniq = None
    if labels is not None:
        assert labels.dtype == torch.uint8 and labels.is_contiguous() and labels.numel() == n_img * HW
        assert label_hist is not None and label_hist.dtype == torch.int64 and label_hist.numel() == n_classes
        lab_at = torch.empty((n_img, n), dtype=torch.int32, device=dev)
        uniq = torch.empty(n_img, dtype=torch.int32, device=dev)
    check(lib().pp_query_stats_at(_ptr(sel_sorted), n_img, n, W, HW, _ptr(labels), n_classes, _ptr(xs), _ptr(ys), _ptr(lab_at),
                                  _ptr(label_hist), _ptr(uniq), _ptr(cov), _stream(sel_sorted)), "pp_query_stats_at")
    return xs, ys, lab_at, uniq, cov


def sparse_ce(logits_lowres, size, px_img, px_idx, px_label, grad_scale=1.0, want_grad=True, want_pred=False,
              n_valid=None):
    """(loss[1], grad_lowres | None, pred_at | None); see pp_sparse_ce in the header.  n_valid: optional device int32
    [1] with the number of valid list entries (the lists then have a fixed capacity: CUDA-graph friendly)."""
    _need_cuda(logits_lowres, px_img, px_idx, px_label)
    n, Cc, h, w = logits_lowres.shape
    H, W = size
    x = logits_lowres.float().contiguous()
    loss = torch.empty(1, dtype=torch.float32, device=x.device)
    grad = torch.zeros_like(x) if want_grad else None
    n_px = int(px_idx.numel())
    pred = torch.empty(n_px, dtype=torch.int32, device=x.device) if want_pred else None
    check(lib().pp_sparse_ce(_ptr(x), n, Cc, h, w, H, W, _ptr(px_img), _ptr(px_idx), _ptr(px_label), n_px,
                             _ptr(n_valid), float(grad_scale), _ptr(loss), _ptr(grad), _ptr(pred), _stream(x)),
          "pp_sparse_ce")
    return loss, grad, pred


class DeviceMetrics:
    """Confusion matrix / loss sum / step count accumulated on the device (pp_metrics_accumulate); read() once per epoch."""

    def __init__(self, n_classes, device):
        self.n_classes = n_classes
        self.confusion = torch.zeros((n_classes, n_classes), dtype=torch.int64, device=device)
        self.loss_sum = torch.zeros(1, dtype=torch.float64, device=device)
        self.n_steps = torch.zeros(1, dtype=torch.int64, device=device)

    def accumulate(self, labels, preds, loss=None, n_valid=None):
        _need_cuda(labels, preds)
        assert labels.dtype == torch.int32 and preds.dtype == torch.int32 and labels.numel() == preds.numel()
        check(lib().pp_metrics_accumulate(_ptr(labels), _ptr(preds), _ptr(n_valid), int(labels.numel()), self.n_classes,
                                          _ptr(loss), _ptr(self.confusion), _ptr(self.loss_sum), _ptr(self.n_steps),
                                          _stream(labels)), "pp_metrics_accumulate")

    def read(self):
        """(confusion [C, C] numpy int64, loss sum, steps) — one device->host sync."""
        return self.confusion.cpu().numpy(), float(self.loss_sum.item()), int(self.n_steps.item())

    def reset(self):
        self.confusion.zero_()
        self.loss_sum.zero_()
        self.n_steps.zero_()


def eval_confusion_upsampled(logits_lowres, size, labels, confusion, want_pred=False):
    """confusion (device int64 [C, C]) += confusion matrix of argmax(upsample(logits_lowres)) vs labels [n, H, W] (device,
    int64 / int32 / uint8); returns the int32 prediction map when want_pred."""
    _need_cuda(logits_lowres, labels, confusion)
    n, Cc, h, w = logits_lowres.shape
    H, W = size
    x = logits_lowres.float().contiguous()
    code = {torch.int64: 0, torch.int32: 1, torch.uint8: 2}.get(labels.dtype)
    if code is None or tuple(labels.shape) != (n, H, W) or not labels.is_contiguous():
        raise PixelPickError("labels must be a contiguous [n, H, W] int64 / int32 / uint8 tensor")
    assert confusion.dtype == torch.int64 and confusion.numel() == Cc * Cc and confusion.is_contiguous()
    pred = torch.empty((n, H, W), dtype=torch.int32, device=x.device) if want_pred else None
    check(lib().pp_eval_confusion_upsampled(_ptr(x), n, Cc, h, w, H, W, _ptr(labels), code, _ptr(confusion), _ptr(pred),
                                            _stream(x)), "pp_eval_confusion_upsampled")
    return pred


def upsample_bilinear_ac(x, size):
    _need_cuda(x)
    n, Cc, h, w = x.shape
    H, W = size
    x = x.float().contiguous()
    out = torch.empty((n, Cc, H, W), dtype=torch.float32, device=x.device)
    check(lib().pp_upsample_bilinear_ac(_ptr(x), n, Cc, h, w, _ptr(out), H, W, _stream(x)), "pp_upsample_bilinear_ac")
    return out


def upsample_bilinear_ac_bwd(grad_out, in_size):
    _need_cuda(grad_out)
    n, Cc, H, W = grad_out.shape
    h, w = in_size
    g = grad_out.float().contiguous()
    gin = torch.zeros((n, Cc, h, w), dtype=torch.float32, device=g.device)
    check(lib().pp_upsample_bilinear_ac_bwd(_ptr(g), n, Cc, H, W, _ptr(gin), h, w, _stream(g)),
          "pp_upsample_bilinear_ac_bwd")
    return gin


def pack_conv_weight(w, cin_pad=None, cout_pad=None, transpose_for_dgrad=False):
    """torch conv weight [Cout, Cin, kh, kw] -> packed bf16 [taps][Cout_pad][Cin_pad] (tap = ky*kw + kx).
    transpose_for_dgrad: weights of the data-gradient convolution, [taps][Cin_pad][Cout_pad] with flipped taps."""
    if transpose_for_dgrad:
        w = w.flip(2, 3).transpose(0, 1)
    co, ci, kh, kw = w.shape
    cin_pad = cin_pad or -(-ci // 64) * 64
    cout_pad = cout_pad or -(-co // 32) * 32
    out = torch.zeros((kh * kw, cout_pad, cin_pad), dtype=torch.bfloat16, device=w.device)
    out[:, :co, :ci] = w.permute(2, 3, 0, 1).reshape(kh * kw, co, ci).to(torch.bfloat16)
    return out


def pack_conv_weights(w, cin=None, fwd_pad=None, dgrad_pad=None, dgrad_out=None):
    """One launch: f32 conv weight [Cout, Cin_total, kh, kw] (first `cin` input channels) -> bf16 operand tensors.
    fwd_pad = (Cout_pad, Cin_pad) -> [taps, Cout_pad, Cin_pad]; dgrad_pad = (Cin_rows, Cout_cols) -> flipped/transposed."""
    _need_cuda(w)
    w = w.detach()
    if w.dtype != torch.float32 or not w.is_contiguous():
        w = w.float().contiguous()
    co, ci_tot, kh, kw = w.shape
    cin = ci_tot if cin is None else cin
    taps = kh * kw
    fwd = torch.empty((taps,) + tuple(fwd_pad), dtype=torch.bfloat16, device=w.device) if fwd_pad else None
    dgr = torch.empty((taps,) + tuple(dgrad_pad), dtype=torch.bfloat16, device=w.device) if dgrad_pad else None
    if dgrad_out is not None:  # pack straight into a [taps, rows, cols] slice of a larger operand tensor
        assert dgrad_out.is_contiguous() and dgrad_out.dtype == torch.bfloat16 and tuple(dgrad_out.shape) == (taps,) + tuple(dgrad_pad)
        dgr = dgrad_out
    check(lib().pp_pack_conv_weight(_ptr(w), co, cin, ci_tot, taps, _ptr(fwd), fwd_pad[0] if fwd_pad else 0,
                                    fwd_pad[1] if fwd_pad else 0, _ptr(dgr), dgrad_pad[0] if dgrad_pad else 0,
                                    dgrad_pad[1] if dgrad_pad else 0, _stream(w)), "pp_pack_conv_weight")
    return fwd, dgr


def pack_conv_weights_tiles(fwd_pad, dgrad_pad, taps):
    """tiles (CTAs) pp_pack_conv_weights_batched spends on one conv; fwd_pad / dgrad_pad as in pack_conv_weights (or None)"""
    n = lib().pp_pack_conv_weights_tiles(*(fwd_pad or (0, 0)), *(dgrad_pad or (0, 0)), int(taps))
    if n < 0:
        raise PixelPickError(f"pp_pack_conv_weights_tiles: {lib().pp_last_error().decode()}")
    return n


def pack_conv_weights_batched(table, n, total_tiles):
    """table: device int64 [n, 12] (see pp_pack_conv_weights_batched): re-packs every listed conv weight in one launch."""
    _need_cuda(table)
    assert table.dtype == torch.int64 and table.is_contiguous() and table.numel() == 12 * n
    check(lib().pp_pack_conv_weights_batched(_ptr(table), n, int(total_tiles), _stream(table)), "pp_pack_conv_weights_batched")


def conv_igemm(x_nhwc, w_packed, cout, dil=1, pre_bias=None, scale=None, shift=None, relu=False, out=None,
               out_mode=0, c_off=0, cin=None, block_n=0):
    """x_nhwc: bf16 [N, H, W, ld_in]; returns bf16 NHWC [N, H, W, ld_out] (out_mode 0) or f32 NCHW (out_mode 1)."""
    _need_cuda(x_nhwc, w_packed)
    assert x_nhwc.dtype == torch.bfloat16 and x_nhwc.is_contiguous() and w_packed.dtype == torch.bfloat16
    N, H, W, ld_in = x_nhwc.shape
    taps, cout_pad, cin_w = w_packed.shape
    cin = cin_w if cin is None else cin
    assert cin == cin_w and cin <= ld_in
    if out is None:
        if out_mode == 0:
            out = torch.empty((N, H, W, cout), dtype=torch.bfloat16, device=x_nhwc.device)
        else:
            out = torch.empty((N, cout, H, W), dtype=torch.float32, device=x_nhwc.device)
    ld_out = out.shape[3] if out_mode == 0 else 0
    for t in (pre_bias, scale, shift):
        assert t is None or (t.dtype == torch.float32 and t.is_contiguous() and t.shape[-1] == cout_pad)
    check(lib().pp_conv_igemm(_ptr(x_nhwc), N, H, W, cin, ld_in, _ptr(w_packed), taps, dil, cout_pad, cout,
                              _ptr(pre_bias), _ptr(scale), _ptr(shift), int(relu), _ptr(out), out_mode, ld_out, c_off,
                              block_n, _stream(x_nhwc)), "pp_conv_igemm")
    return out


def conv_igemm_multi(x_nhwc, w_packed, entries, cout, out=None, block_n=0, pre_bias=None):
    """Generalised implicit GEMM: entries = [(dy, dx, c0), ...]; entry t convolves channels [c0, c0 + Cin) of x shifted by
    (dy, dx) with weight slice t of w_packed [n_entries][cout_pad][Cin]; all entries accumulate into one output."""
    _need_cuda(x_nhwc, w_packed)
    assert x_nhwc.dtype == torch.bfloat16 and x_nhwc.is_contiguous() and w_packed.dtype == torch.bfloat16
    N, H, W, ld_in = x_nhwc.shape
    n_e, cout_pad, cin = w_packed.shape
    assert n_e == len(entries)
    if out is None:
        out = torch.empty((N, H, W, cout), dtype=torch.bfloat16, device=x_nhwc.device)
    arr = lambda k: (C.c_int * n_e)(*[int(e[k]) for e in entries])
    if pre_bias is not None:  # per-image bias [N, cout_pad] added to every pixel of the image (fp32)
        assert pre_bias.dtype == torch.float32 and pre_bias.is_contiguous() and tuple(pre_bias.shape) == (N, cout_pad)
    check(lib().pp_conv_igemm_multi(_ptr(x_nhwc), N, H, W, ld_in, ld_in, cin, _ptr(w_packed), n_e, arr(0), arr(1), arr(2),
                                    cout_pad, cout, _ptr(pre_bias), None, None, 0, None, 0, _ptr(out), 0, out.shape[3], 0, block_n,
                                    _stream(x_nhwc)), "pp_conv_igemm_multi")
    return out


_TAPS = {}


def _tap_arrays(entries):
    n = len(entries)
    return tuple((C.c_int * n)(*[int(e[k]) for e in entries]) for k in range(3))


def conv_fused(x_nhwc, w_packed, cout, dil=1, scale=None, shift=None, act=0, res=None, out=None, flatten=True, stats=None,
               entries=None, c_off=0):
    """conv + per-channel affine (folded BatchNorm) + activation (+ residual) in one launch on activations whose channel
    count need not be a multiple of 64 (TMA zero-fills the K padding).  x_nhwc bf16 [N,H,W,C]; w_packed [taps][cout_pad][Cin_pad];
    act 0/1/2 = none/ReLU/ReLU6; res bf16 [N,H,W,cout].  1x1 convs are run on the flattened pixel list (no ragged tiles).
    stats: f32 [2, cout] zeroed accumulator -> += per-channel (sum, sum of squares) of the written values (TRAIN-mode
    BatchNorm statistics straight from the conv epilogue).  entries: explicit tap table [(dy, dx, c0), ...] (default: the
    1x1 / dilated 3x3 table of `dil`), e.g. the space-to-depth form of a stride-2 convolution."""
    _need_cuda(x_nhwc, w_packed, res, stats)
    assert x_nhwc.dtype == torch.bfloat16 and x_nhwc.is_contiguous() and w_packed.dtype == torch.bfloat16
    N, H, W, Cc = x_nhwc.shape
    taps, cout_pad, cin_pad = w_packed.shape
    if out is None:
        out = torch.empty((N, H, W, cout), dtype=torch.bfloat16, device=x_nhwc.device)
    if entries is None:
        key = (taps, dil)
        if key not in _TAPS:
            ent = [(0, 0, 0)] if taps == 1 else [((t // 3 - 1) * dil, (t % 3 - 1) * dil, 0) for t in range(9)]
            _TAPS[key] = _tap_arrays(ent)
        dy, dx, c0 = _TAPS[key]
    else:
        assert len(entries) == taps
        dy, dx, c0 = _tap_arrays(entries)
    n_, h_, w_ = N, H, W
    M = N * H * W
    if taps == 1 and flatten and M % 16 == 0 and entries is None:
        n_, h_, w_ = 1, M // 16, 16  # a 1x1 conv is a plain GEMM over pixels: 8x16 tiles of the flattened list
    for t in (scale, shift):
        assert t is None or (t.dtype == torch.float32 and t.is_contiguous() and t.numel() == cout_pad)
    if stats is not None:
        assert stats.dtype == torch.float32 and stats.is_contiguous() and tuple(stats.shape) == (2, cout)
    check(lib().pp_conv_igemm_stats(_ptr(x_nhwc), n_, h_, w_, Cc, Cc, cin_pad, _ptr(w_packed), taps, dy, dx, c0, cout_pad, cout,
                                    None, _ptr(scale), _ptr(shift), int(act), _ptr(res), res.shape[-1] if res is not None else 0,
                                    _ptr(out), 0, out.shape[3], c_off, 0, _ptr(stats), _stream(x_nhwc)), "pp_conv_igemm_stats")
    return out


def conv_wgrad_multi(x_nhwc, cin, dy_nhwc, cout, entries, splits=0, out=None):
    """weight gradient of a conv given as a tap table: f32 [taps][Cin_rows][ld] with entry t = sum_p x[p + (dy, dx), c0 + ci] *
    dy[p, co]; Cin_rows = cin rounded up to 8, ld = cout rounded up to the kernel's output-channel tile."""
    _need_cuda(x_nhwc, dy_nhwc)
    assert x_nhwc.dtype == torch.bfloat16 and dy_nhwc.dtype == torch.bfloat16 and x_nhwc.is_contiguous() and dy_nhwc.is_contiguous()
    N, H, W, ld_x = x_nhwc.shape
    assert tuple(dy_nhwc.shape[:3]) == (N, H, W) and dy_nhwc.shape[3] >= cout
    bn = 256 if cout > 128 else (128 if cout > 64 else 64)
    ld = -(-cout // bn) * bn
    rows = -(-cin // 8) * 8
    taps = len(entries)
    if out is None:
        dw = torch.zeros((taps, rows, ld), dtype=torch.float32, device=x_nhwc.device)
    else:  # a zeroed slice of a per-network gradient arena
        dw = out
        assert dw.dtype == torch.float32 and dw.is_contiguous() and tuple(dw.shape) == (taps, rows, ld)
    dy, dx, c0 = _tap_arrays(entries)
    check(lib().pp_conv_wgrad_multi(_ptr(x_nhwc), ld_x, ld_x, cin, _ptr(dy_nhwc), dy_nhwc.shape[3], cout, N, H, W, taps, dy, dx, c0,
                                    _ptr(dw), rows, ld, splits, _stream(x_nhwc)), "pp_conv_wgrad_multi")
    return dw


def conv_wgrad(x_nhwc, cin, dy_nhwc, cout_pad, taps, dil=1, splits=0):
    """dW as f32 [taps][Cin_rows][Cout_pad] (Cin_rows = Cin rounded up to 128).  dy_nhwc may be a channel-slice view
    [..., c0:c0+cout_pad] of a wider contiguous NHWC buffer."""
    _need_cuda(x_nhwc, dy_nhwc)
    assert x_nhwc.dtype == torch.bfloat16 and dy_nhwc.dtype == torch.bfloat16
    assert x_nhwc.is_contiguous() and dy_nhwc.stride(3) == 1 and dy_nhwc.stride(1) == dy_nhwc.shape[2] * dy_nhwc.stride(2) \
        and dy_nhwc.stride(0) == dy_nhwc.shape[1] * dy_nhwc.stride(1) and dy_nhwc.data_ptr() % 16 == 0
    N, H, W, ld_x = x_nhwc.shape
    ld_dy = dy_nhwc.stride(2)
    rows = -(-cin // 128) * 128
    dw = torch.zeros((taps, rows, cout_pad), dtype=torch.float32, device=x_nhwc.device)
    check(lib().pp_conv_wgrad(_ptr(x_nhwc), ld_x, cin, _ptr(dy_nhwc), ld_dy, cout_pad, N, H, W, taps, dil, _ptr(dw),
                              rows, splits, _stream(x_nhwc)), "pp_conv_wgrad")
    return dw


def bn_stats(raw, c_off, C):
    """per-channel (sum, sum of squares) of raw[..., c_off:c_off+C] as f32 [2, C]; raw is bf16 [..., ld]."""
    _need_cuda(raw)
    ld = raw.shape[-1]
    M = raw.numel() // ld
    sums = torch.empty((2, C), dtype=torch.float32, device=raw.device)
    check(lib().pp_bn_stats(_ptr(raw), M, ld, c_off, C, _ptr(sums), _stream(raw)), "pp_bn_stats")
    return sums


def bn_finalize(sums, M, bn, Cpad=None, update_running=True, count=True):
    """f32 [4, Cpad] = (scale, shift, mean, rstd) from bn_stats output and an nn.BatchNorm2d's parameters."""
    C = sums.shape[1]
    Cpad = Cpad or C
    out = torch.empty((4, Cpad), dtype=torch.float32, device=sums.device)
    upd = update_running and bn.track_running_stats and bn.running_mean is not None
    mom = 0.1 if bn.momentum is None else bn.momentum
    check(lib().pp_bn_finalize(_ptr(sums), C, M, bn.eps, mom, _ptr(bn.weight.detach()), _ptr(bn.bias.detach()),
                               _ptr(bn.running_mean) if upd else None, _ptr(bn.running_var) if upd else None,
                               _ptr(out), Cpad, _stream(sums)), "pp_bn_finalize")
    if upd and count:  # count=False: the caller advances the counters of many layers with one batched op
        bn.num_batches_tracked += 1
    return out


def bn_apply(raw, c_off_in, C, scale, shift, relu, out, c_off_out, drop_p=0.0, seed=0, offset=0, seed_dev=None, res=None):
    """out = dropout(act(raw * scale + shift [+ res])); res: bf16 [..., ld_res] residual added before the activation."""
    _need_cuda(raw, out, res)
    ld_in, ld_out = raw.shape[-1], out.shape[-1]
    M = raw.numel() // ld_in
    check(lib().pp_bn_apply_res(_ptr(raw), M, ld_in, c_off_in, C, _ptr(scale), _ptr(shift), int(relu), float(drop_p),
                                int(seed), int(offset), _ptr(seed_dev), _ptr(res), res.shape[-1] if res is not None else 0,
                                _ptr(out), ld_out, c_off_out, _stream(raw)), "pp_bn_apply")
    return out


def bn_apply_stats(raw, C, sums, bn, act, out, res=None, update_running=True):
    """finalize (sums [2, C] from the conv epilogue -> scale / shift / mean / rstd, running statistics) + normalise + activation
    (+ residual) in one launch; returns stats f32 [4, C] for bn_bwd.  num_batches_tracked is the caller's business."""
    _need_cuda(raw, out, res, sums)
    ld_in, ld_out = raw.shape[-1], out.shape[-1]
    M = raw.numel() // ld_in
    stats = torch.empty((4, C), dtype=torch.float32, device=raw.device)
    upd = update_running and bn.track_running_stats and bn.running_mean is not None
    mom = 0.1 if bn.momentum is None else bn.momentum
    check(lib().pp_bn_apply_stats(_ptr(raw), M, ld_in, 0, C, _ptr(sums), _ptr(bn.weight.detach()), _ptr(bn.bias.detach()), bn.eps,
                                  mom, _ptr(bn.running_mean) if upd else None, _ptr(bn.running_var) if upd else None,
                                  _ptr(stats), int(act), _ptr(res), res.shape[-1] if res is not None else 0, _ptr(out), ld_out, 0,
                                  _stream(raw)), "pp_bn_apply_stats")
    return stats


# Single-launch (cooperative) BatchNorm passes.  PP_BN_FUSED=0 falls back to the separate stats/finalize/apply kernels
# (same arithmetic; used by the tests to cross-check the two paths).
BN_FUSED = os.environ.get("PP_BN_FUSED", "1") != "0"


def bn_scratch(bn, C, device):
    """Zero-initialised per-layer scratch (2*C partial sums + 2 counters) kept on the nn.BatchNorm2d object; every fused
    launch leaves it zeroed again."""
    sc = getattr(bn, "_pp_scratch", None)
    if sc is None or sc.device != device or sc.numel() != 2 * C + 2:
        sc = torch.zeros(2 * C + 2, dtype=torch.float32, device=device)
        bn._pp_scratch = sc
    return sc


def bn_fwd_fused(raw, c_off_in, C, bn, relu, out, c_off_out, drop_p=0.0, seed=0, offset=0, seed_dev=None, res=None,
                 update_running=True):
    """train-mode BatchNorm statistics + normalise/activation/residual/dropout in ONE launch; returns stats f32 [4, C]."""
    _need_cuda(raw, out, res)
    ld_in, ld_out = raw.shape[-1], out.shape[-1]
    M = raw.numel() // ld_in
    stats = torch.empty((4, C), dtype=torch.float32, device=raw.device)
    upd = update_running and bn.track_running_stats and bn.running_mean is not None
    mom = 0.1 if bn.momentum is None else bn.momentum
    check(lib().pp_bn_fwd_fused(_ptr(raw), M, ld_in, c_off_in, C, _ptr(bn.weight.detach()), _ptr(bn.bias.detach()),
                                _ptr(bn.running_mean) if upd else None, _ptr(bn.running_var) if upd else None,
                                _ptr(bn.num_batches_tracked) if upd else None, bn.eps, mom, int(relu), float(drop_p),
                                int(seed), int(offset), _ptr(seed_dev), _ptr(res), res.shape[-1] if res is not None else 0,
                                _ptr(out), ld_out, c_off_out, _ptr(stats), _ptr(bn_scratch(bn, C, raw.device)),
                                _stream(raw)), "pp_bn_fwd_fused")
    return stats


def bn_bwd(dy, c_off_dy, raw, c_off_raw, C, scale, shift, mean, rstd, relu, drop_p=0.0, seed=0, offset=0,
           seed_dev=None, res=None, draw_out=None, draw_c_off=0, scratch=None):
    """returns (draw bf16 [M, C], sums f32 [2, C] = (d beta, d gamma)) — plus dres bf16 [M, C] (gradient wrt the
    residual that was added before the activation) when `res` is given."""
    _need_cuda(dy, raw, res)
    ld_dy, ld_raw = dy.shape[-1], raw.shape[-1]
    M = raw.numel() // ld_raw
    if draw_out is None:
        draw, ld_draw = torch.empty((M, C), dtype=torch.bfloat16, device=raw.device), C
    else:  # write into a channel slice of a wider gradient buffer
        draw, ld_draw = draw_out, draw_out.shape[-1]
        assert draw.dtype == torch.bfloat16 and draw.is_contiguous() and draw.numel() // ld_draw == M
    sums = torch.empty((2, C), dtype=torch.float32, device=raw.device)
    dres = torch.empty((M, C), dtype=torch.bfloat16, device=raw.device) if res is not None else None
    if scratch is not None and BN_FUSED:  # one cooperative launch (reduce -> grid barrier -> apply)
        check(lib().pp_bn_bwd_fused(_ptr(dy), ld_dy, c_off_dy, _ptr(raw), ld_raw, c_off_raw, M, C, _ptr(scale), _ptr(shift),
                                    _ptr(mean), _ptr(rstd), int(relu), float(drop_p), int(seed), int(offset),
                                    _ptr(seed_dev), _ptr(res), res.shape[-1] if res is not None else 0, _ptr(dres),
                                    _ptr(sums), _ptr(draw), ld_draw, draw_c_off, _ptr(scratch), _stream(raw)),
              "pp_bn_bwd_fused")
    else:
        check(lib().pp_bn_bwd_res(_ptr(dy), ld_dy, c_off_dy, _ptr(raw), ld_raw, c_off_raw, M, C, _ptr(scale), _ptr(shift),
                                  _ptr(mean), _ptr(rstd), int(relu), float(drop_p), int(seed), int(offset), _ptr(seed_dev),
                                  _ptr(res), res.shape[-1] if res is not None else 0, _ptr(dres),
                                  _ptr(sums), _ptr(draw), ld_draw, draw_c_off, _stream(raw)), "pp_bn_bwd")
    if res is not None:
        return draw, sums, dres
    return draw, sums


def maxpool3x3s2_fwd(x, want_code=True):
    """nn.MaxPool2d(3, 2, 1) on contiguous NHWC bf16 -> (y, code uint8 | None)"""
    _need_cuda(x)
    assert x.dtype == torch.bfloat16 and x.is_contiguous() and x.dim() == 4
    N, H, W, Cc = x.shape
    Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    y = torch.empty((N, Ho, Wo, Cc), dtype=torch.bfloat16, device=x.device)
    code = torch.empty((N, Ho, Wo, Cc), dtype=torch.uint8, device=x.device) if want_code else None
    check(lib().pp_maxpool3x3s2_fwd(_ptr(x), N, H, W, Cc, _ptr(y), _ptr(code), _stream(x)), "pp_maxpool3x3s2_fwd")
    return y, code


def maxpool3x3s2_bwd(dy, code, in_hw):
    _need_cuda(dy, code)
    assert dy.dtype == torch.bfloat16 and dy.is_contiguous() and code.dtype == torch.uint8 and code.is_contiguous()
    N, Ho, Wo, Cc = dy.shape
    H, W = in_hw
    dx = torch.empty((N, H, W, Cc), dtype=torch.bfloat16, device=dy.device)
    check(lib().pp_maxpool3x3s2_bwd(_ptr(dy), _ptr(code), N, H, W, Cc, _ptr(dx), _stream(dy)), "pp_maxpool3x3s2_bwd")
    return dx


def _dw_out(Hi, Wi, stride, dil):
    return (Hi - 2 * dil - 1) // stride + 1, (Wi - 2 * dil - 1) // stride + 1


def _dw_check(x, w):
    _need_cuda(x, w)
    if x.dtype != torch.bfloat16 or not x.is_contiguous() or w.dtype != torch.float32 or not w.is_contiguous():
        raise PixelPickError("depthwise conv: x must be contiguous bf16 NHWC and w contiguous f32 [C,1,3,3]")
    if w.numel() != x.shape[-1] * 9:
        raise PixelPickError("depthwise conv: weight does not match the channel count")


def dwconv_fwd(x, w, stride, dil, scale=None, shift=None, act=0):
    """valid depthwise 3x3 of bf16 NHWC x [N,Hi,Wi,C] with f32 w [C,1,3,3] -> bf16 [N,Ho,Wo,C]; with scale/shift (f32 [C])
    the epilogue applies a folded BatchNorm + activation (0 none, 1 ReLU, 2 ReLU6)."""
    _dw_check(x, w)
    N, Hi, Wi, Cc = x.shape
    Ho, Wo = _dw_out(Hi, Wi, stride, dil)
    y = torch.empty((N, Ho, Wo, Cc), dtype=torch.bfloat16, device=x.device)
    check(lib().pp_dwconv3x3_fwd_bnact(_ptr(x), _ptr(w), _ptr(scale), _ptr(shift), int(act), _ptr(y), N, Hi, Wi, Cc, stride,
                                       dil, _stream(x)), "pp_dwconv3x3_fwd")
    return y


def dwconv_dgrad(dy, w, in_hw, stride, dil):
    _dw_check(dy, w)
    N, Cc = dy.shape[0], dy.shape[-1]
    Hi, Wi = in_hw
    if tuple(dy.shape[1:3]) != _dw_out(Hi, Wi, stride, dil):
        raise PixelPickError("depthwise dgrad: dy does not match the input size")
    dx = torch.empty((N, Hi, Wi, Cc), dtype=torch.bfloat16, device=dy.device)
    check(lib().pp_dwconv3x3_dgrad(_ptr(dy), _ptr(w), _ptr(dx), N, Hi, Wi, Cc, stride, dil, _stream(dy)), "pp_dwconv3x3_dgrad")
    return dx


def dwconv_wgrad(x, dy, stride, dil):
    _need_cuda(x, dy)
    N, Hi, Wi, Cc = x.shape
    if tuple(dy.shape) != (N,) + _dw_out(Hi, Wi, stride, dil) + (Cc,) or not (x.is_contiguous() and dy.is_contiguous()) \
            or x.dtype != torch.bfloat16 or dy.dtype != torch.bfloat16:
        raise PixelPickError("depthwise wgrad: x / dy must be contiguous bf16 NHWC of matching sizes")
    dw = torch.empty((Cc, 1, 3, 3), dtype=torch.float32, device=x.device)
    check(lib().pp_dwconv3x3_wgrad(_ptr(x), _ptr(dy), _ptr(dw), N, Hi, Wi, Cc, stride, dil, _stream(x)), "pp_dwconv3x3_wgrad")
    return dw


def upsample_nhwc(x, out, c_off, C=None):
    """bilinear align_corners=True of bf16 NHWC x[..., :C] into out[..., c_off:c_off+C]."""
    _need_cuda(x, out)
    N, h, w, ld_in = x.shape
    _, H, W, ld_out = out.shape
    C = ld_in if C is None else C
    check(lib().pp_upsample_nhwc_bf16(_ptr(x), N, h, w, C, ld_in, _ptr(out), H, W, ld_out, c_off, _stream(x)),
          "pp_upsample_nhwc_bf16")
    return out


def upsample_nhwc_bwd(grad_out, c_off, C, in_hw):
    _need_cuda(grad_out)
    N, H, W, ld = grad_out.shape
    h, w = in_hw
    gin = torch.empty((N, h, w, C), dtype=torch.float32, device=grad_out.device)
    check(lib().pp_upsample_nhwc_bf16_bwd(_ptr(grad_out), N, H, W, ld, c_off, C, _ptr(gin), h, w, _stream(grad_out)),
          "pp_upsample_nhwc_bf16_bwd")
    return gin


def to_nhwc_bf16(x_nchw, out=None, c_off=0, ld=None):
    """any-strided [N, C, H, W] f32/bf16 -> bf16 NHWC (channel slice of `out`, zero padded when allocated here)."""
    _need_cuda(x_nchw)
    N, Cc, H, W = x_nchw.shape
    if out is None:
        ld = ld or -(-Cc // 64) * 64
        out = (torch.zeros if ld != Cc else torch.empty)((N, H, W, ld), dtype=torch.bfloat16, device=x_nchw.device)
    sn, sc, sh, sw = x_nchw.stride()
    check(lib().pp_to_nhwc_bf16(_ptr(x_nchw), _dtype_code(x_nchw), sn, sc, sh, sw, N, Cc, H, W, _ptr(out), out.shape[3],
                                c_off, _stream(x_nchw)), "pp_to_nhwc_bf16")
    return out


def session_check_host_buffers(what, n, Cc, H, W, k, n_sel, logits=None, masks=(), pos=None, sel=None, topk=None):
    """Host-buffer contract of pp_acq_session_* (include/pixelpick_b200.h): contiguous CPU tensors,
    logits f32 [n, C, H, W]; labelled / void masks 1 byte per pixel [n, H, W]; pick positions int32 [n, n_sel] in [0, k);
    outputs int32 [n, n_sel] (and [n, k] for the sorted list).  The library memcpy's exactly these sizes."""
    def want(name, t, shape, dtypes):
        if t is None:
            return
        if t.is_cuda or not t.is_contiguous():
            raise PixelPickError(f"{what}: {name} must be a contiguous host tensor")
        numel = 1
        for d in shape:
            numel *= d
        if t.numel() != numel or t.shape[0] != shape[0] or t.dtype not in dtypes:  # any view of the right block is fine
            raise PixelPickError(f"{what}: {name} is {t.dtype} {tuple(t.shape)}, expected {dtypes[0]} {tuple(shape)}")
    if n <= 0:
        raise PixelPickError(f"{what}: empty batch")
    want("logits", logits, (n, Cc, H, W), (torch.float32,))
    for name, m in zip(("labelled mask", "void mask"), masks):
        want(name, m, (n, H, W), (torch.uint8, torch.bool))
    want("pick positions", pos, (n, n_sel), (torch.int32,))
    if pos is not None and pos.numel() and (int(pos.min()) < 0 or int(pos.max()) >= k):
        raise PixelPickError(f"{what}: pick positions must lie in [0, {k})")
    want("selected indices (output)", sel, (n, n_sel), (torch.int32,))
    want("sorted top-k indices (output)", topk, (n, k), (torch.int32,))


class AcqSession:
    """Host-buffer acquisition session (pp_acq_session_*): numpy / pinned-torch in, numpy out."""

    def __init__(self, chunk_imgs, Cc, H, W, k, n_sel):
        h = _vp()
        check(lib().pp_acq_session_create(C.byref(h), chunk_imgs, Cc, H, W, k, n_sel), "pp_acq_session_create")
        self._h = h
        self.shape = (Cc, H, W)
        self.k, self.n_sel = k, n_sel

    def _check(self, what, n, logits=None, masks=(), pos=None, sel=None, topk=None):
        """The C side copies fixed-size blocks out of / into these host buffers: shapes and dtypes must be exact."""
        Cc, H, W = self.shape
        session_check_host_buffers(what, n, Cc, H, W, self.k, self.n_sel, logits, masks, pos, sel, topk)

    def run(self, h_logits, h_labelled, h_void, strategy, h_pos, h_sel, h_topk=None):
        """All arguments are CPU torch tensors (pinned for speed); h_sel (and h_topk) are outputs."""
        n = h_logits.shape[0]
        self._check("AcqSession.run", n, h_logits, (h_labelled, h_void), h_pos, h_sel, h_topk)
        check(lib().pp_acq_session_run_host(self._h, _ptr(h_logits), _ptr(h_labelled), _ptr(h_void), n,
                                            STRATEGIES[strategy], _ptr(h_pos), _ptr(h_sel), _ptr(h_topk)),
              "pp_acq_session_run_host")
        return h_sel

    def begin(self, h_logits, h_labelled, h_void, strategy):
        """asynchronous first half (H2D + scoring + radix select); draw the pick positions, then call finish()."""
        self._check("AcqSession.begin", h_logits.shape[0], h_logits, (h_labelled, h_void))
        check(lib().pp_acq_session_begin_host(self._h, _ptr(h_logits), _ptr(h_labelled), _ptr(h_void), h_logits.shape[0],
                                              STRATEGIES[strategy]), "pp_acq_session_begin_host")
        self._pending_n = h_logits.shape[0]

    def finish(self, h_pos, h_sel):
        self._check("AcqSession.finish", getattr(self, "_pending_n", None) or h_sel.shape[0], pos=h_pos, sel=h_sel)
        self._pending_n = None
        check(lib().pp_acq_session_finish_host(self._h, _ptr(h_pos), _ptr(h_sel)), "pp_acq_session_finish_host")
        return h_sel

    def close(self):
        if self._h is not None:
            lib().pp_acq_session_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class AdamPlan:
    """The argument arrays of pp_adam_step_multi for a fixed list of tensors (host arrays of device pointers): built once,
    reused while no tensor moves (`key` = the data pointers it was built from)."""

    def __init__(self, params, grads, exp_avg, exp_avg_sq, group, lrs, beta1, beta2, eps, weight_decay, step):
        _need_cuda(*params, *grads, *exp_avg, *exp_avg_sq, *lrs, step)
        for ts in (params, grads, exp_avg, exp_avg_sq):
            for t in ts:
                if t.dtype != torch.float32 or not t.is_contiguous():
                    raise PixelPickError("pp_adam_step_multi: parameters, gradients and moments must be contiguous fp32")
        if any(g.numel() != p.numel() or m.numel() != p.numel() or v.numel() != p.numel()
               for p, g, m, v in zip(params, grads, exp_avg, exp_avg_sq)):
            raise PixelPickError("pp_adam_step_multi: gradient / moment sizes differ from the parameter's")
        if step.dtype != torch.float32 or any(lr.dtype != torch.float32 for lr in lrs):
            raise PixelPickError("pp_adam_step_multi: step and learning rates must be fp32 device scalars")
        n, ng = len(params), len(lrs)
        self.n, self.ng = n, ng
        self.key = self.make_key(params, grads, exp_avg, exp_avg_sq, lrs, step)
        arr = lambda ts: (C.c_void_p * len(ts))(*[t.data_ptr() for t in ts])
        self.p, self.g, self.m, self.v, self.lr = arr(params), arr(grads), arr(exp_avg), arr(exp_avg_sq), arr(lrs)
        self.numel = (C.c_longlong * n)(*[t.numel() for t in params])
        self.group = (C.c_int * n)(*group)
        dbl = lambda xs: (C.c_double * ng)(*[float(x) for x in xs])
        self.b1, self.b2, self.eps, self.wd = dbl(beta1), dbl(beta2), dbl(eps), dbl(weight_decay)
        self.step = step
        self.keep = (params, grads, exp_avg, exp_avg_sq, lrs)

    @staticmethod
    def make_key(params, grads, exp_avg, exp_avg_sq, lrs, step):
        return tuple(t.data_ptr() for ts in (params, grads, exp_avg, exp_avg_sq, lrs) for t in ts) + (step.data_ptr(),)

    def launch(self):
        check(lib().pp_adam_step_multi(self.n, self.p, self.g, self.m, self.v, self.numel, self.group, self.ng, self.lr,
                                       self.b1, self.b2, self.eps, self.wd, _ptr(self.step), _stream(self.step)),
              "pp_adam_step_multi")
