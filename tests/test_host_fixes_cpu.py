"""CPU: host-side contracts added after the round-1 review - label validation / capacity overflow of the labelled-pixel
list (model.py:108-116 semantics of F.cross_entropy), the Poly scheduler on tensor learning rates, and the class-count
guard of the fused upsample+score path."""
import numpy as np
import pytest
import torch

from pixelpick_b200 import _lib
from pixelpick_b200.loss import LabelCapacityError, labelled_pixel_list_host
from pixelpick_b200.utils import Poly


def _batch(B=2, H=8, W=8, n_lab=5, C=4, ignore=255, seed=0):
    rs = np.random.RandomState(seed)
    y = rs.randint(0, C, size=(B, H, W)).astype(np.int64)
    q = np.zeros((B, H * W), dtype=np.uint8)
    for i in range(B):
        q[i, rs.choice(H * W, n_lab, replace=False)] = 1
    return torch.from_numpy(y), torch.from_numpy(q.reshape(B, H, W))


def test_label_list_matches_dense_mask():
    y, q = _batch()
    y[0, 0, 0] = 255
    q[0, 0, 0] = 1  # labelled but void: F.cross_entropy ignores it
    pi, px, pl, n = labelled_pixel_list_host(y, q, 255, capacity=32, n_classes=4)
    keep = (q.bool() & (y != 255)).reshape(2, -1)
    assert int(n) == int(keep.sum())
    img, idx = np.nonzero(keep.numpy())
    assert np.array_equal(pi[: int(n)].numpy(), img) and np.array_equal(px[: int(n)].numpy(), idx)
    assert np.array_equal(pl[: int(n)].numpy(), y.reshape(2, -1).numpy()[img, idx])
    assert not pi[int(n):].any() and not pl[int(n):].any()


def test_label_out_of_range_raises_like_cross_entropy():
    y, q = _batch()
    pos = np.argwhere(q.numpy()[1])[0]
    y[1, pos[0], pos[1]] = 9  # neither ignore_index nor a class
    with pytest.raises(IndexError, match="Target 9 is out of bounds"):
        labelled_pixel_list_host(y, q, 255, capacity=32, n_classes=4)
    ref = torch.where(q.bool(), y, torch.full_like(y, 255))
    with pytest.raises(IndexError, match="Target 9 is out of bounds"):
        torch.nn.functional.cross_entropy(torch.zeros(2, 4, 8, 8), ref, ignore_index=255)
    y[1, pos[0], pos[1]] = -3
    with pytest.raises(IndexError):
        labelled_pixel_list_host(y, q, 255, capacity=32, n_classes=4)


def test_capacity_overflow_is_a_distinct_error():
    y, q = _batch(n_lab=20)
    with pytest.raises(LabelCapacityError):
        labelled_pixel_list_host(y, q, 255, capacity=16, n_classes=4)
    assert issubclass(LabelCapacityError, _lib.PixelPickError)


def test_poly_reads_tensor_learning_rates_once():
    """capturable Adam keeps tensor learning rates; the scheduler must not read them back per step."""
    w = [torch.nn.Parameter(torch.zeros(3)), torch.nn.Parameter(torch.zeros(2))]

    class CountingLR(torch.Tensor):
        reads = 0

        def __float__(self):
            CountingLR.reads += 1
            return super().__float__()

    def mk(v):
        return torch.tensor(v).as_subclass(CountingLR)

    opt_t = torch.optim.Adam([{"params": [w[0]], "lr": mk(5e-5)}, {"params": [w[1]], "lr": mk(5e-4)}])
    opt_f = torch.optim.Adam([{"params": [w[0]], "lr": 5e-5}, {"params": [w[1]], "lr": 5e-4}])
    st, sf = Poly(opt_t, 2, 5), Poly(opt_f, 2, 5)
    after_init = CountingLR.reads
    for epoch in (1, 2):
        for _ in range(5):
            st.step(epoch=epoch - 1)
            sf.step(epoch=epoch - 1)
            for gt, gf in zip(opt_t.param_groups, opt_f.param_groups):
                assert abs(float(torch.as_tensor(gt["lr"]).as_subclass(torch.Tensor)) - gf["lr"]) < 1e-6 * gf["lr"] + 1e-12  # fp32 tensor
    assert CountingLR.reads == after_init  # no per-iteration read-back of the base learning rates


def test_upsampled_score_guard_lists_the_instantiated_class_counts():
    assert _lib.UPSAMPLED_SCORE_CLASSES == (11, 19, 21)


def test_topk_workspace_tracks_when_the_memset_may_be_skipped(monkeypatch):
    """include/pixelpick_b200.h, pp_acq_topk_prepare: a completed select hands the workspace back zeroed FOR ITS BATCH SIZE
    (the bucket state it leaves behind lies where a larger batch keeps histograms), so TopKWorkspace skips the per-batch
    memset only then.  The kernels are not called here: only the bookkeeping around them."""
    real = _lib.lib()
    calls = []

    class Fake:
        def __getattr__(self, name):
            return getattr(real, name)

        def pp_acq_topk_prepare(self, *a):
            calls.append(a[2])  # n_img of the zeroed region
            return 0

    monkeypatch.setattr(_lib, "_lib", Fake())
    monkeypatch.setattr(_lib, "_stream", lambda t: None)  # no CUDA stream on the CPU
    ws = object.__new__(_lib.TopKWorkspace)  # (the constructor wants a 256-byte aligned DEVICE buffer)
    ws.n_img, ws.HW, ws.k, ws.nbytes, ws.buf, ws._clean = 4, 64 * 128, 409, 1 << 20, torch.empty(1 << 20, dtype=torch.uint8), None
    ws.prepare()
    ws.prepare()
    assert calls == [4] and ws._clean == "all"          # always the full-capacity region, once
    ws._begin_fill(4)                                    # pp_acq_score(hist0=...) on a zeroed workspace: no memset
    assert calls == [4] and ws._clean is None
    ws._begin_select(4, True)
    ws._clean = 4                                        # what the wrappers set after a completed select
    ws.prepare()
    ws._begin_fill(4)                                    # same batch size again: still no memset
    assert calls == [4]
    ws._begin_fill(4)                                    # a fill that was NOT followed by a select: zero again
    assert calls == [4, 4]
    ws._begin_select(4, True)
    ws._clean = 4
    ws._begin_fill(3)                                    # another batch size: the old bucket state is in the way
    assert calls == [4, 4, 4]
    ws._clean = 3
    ws._begin_select(3, False)                           # the select builds the histogram itself: clean for 3 is enough
    assert calls == [4, 4, 4] and ws._clean is None
    ws._begin_select(3, False)                           # ... but not after an unfinished select
    assert calls == [4, 4, 4, 4]
    ws._clean = 3
    ws.prepare(force=True)
    assert calls == [4, 4, 4, 4, 4] and ws._clean == "all"
