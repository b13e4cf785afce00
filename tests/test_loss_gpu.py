"""GPU: sparse-pixel CE (+fused align_corners upsample) and the bilinear resize vs torch fp32 on CPU
(the reference's own arithmetic: model.py:108-116, deeplab.py:55).  Tolerance: loss 1e-5 abs, grads 1e-6 abs."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from pixelpick_b200 import _lib

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")


def _reference(lr, size, pi, px, pl, ignore=255):
    lr = lr.clone().requires_grad_(True)
    full = F.interpolate(lr, size=size, mode="bilinear", align_corners=True) if tuple(lr.shape[2:]) != tuple(size) else lr
    n = lr.shape[0]
    y = torch.full((n, size[0] * size[1]), ignore, dtype=torch.long)
    y[pi.long(), px.long()] = pl.long()
    loss = F.cross_entropy(full, y.view(n, *size), ignore_index=ignore)  # model.py:110,116
    loss.backward()
    pred = full.detach().argmax(dim=1).view(n, -1)[pi.long(), px.long()]
    return loss.detach(), lr.grad, pred


@pytest.mark.parametrize("n,C,lr,size,npx", [(4, 19, (64, 128), (256, 512), 40), (4, 11, (90, 120), (360, 480), 400),
                                             (2, 21, (80, 80), (320, 320), 7), (2, 19, (32, 48), (32, 48), 50),
                                             (1, 40, (9, 13), (33, 50), 64)])
def test_sparse_ce_matches_dense_reference(n, C, lr, size, npx):
    g = torch.Generator().manual_seed(npx)
    low = torch.randn((n, C, *lr), generator=g) * 2
    rs = np.random.RandomState(npx)
    flat = rs.choice(n * size[0] * size[1], npx, replace=False)
    pi = torch.from_numpy((flat // (size[0] * size[1])).astype(np.int32))
    px = torch.from_numpy((flat % (size[0] * size[1])).astype(np.int32))
    pl = torch.from_numpy(rs.randint(0, C, npx).astype(np.int32))
    loss, grad, pred = _lib.sparse_ce(low.to(DEV), size, pi.to(DEV), px.to(DEV), pl.to(DEV), want_pred=True)
    rl, rg, rp = _reference(low, size, pi, px, pl)
    assert abs(loss.item() - rl.item()) < 1e-5
    assert torch.allclose(grad.cpu(), rg, atol=1e-6, rtol=1e-5)
    assert (pred.cpu().long() == rp).float().mean() > 0.99  # argmax can differ only on exact near-ties


def test_sparse_ce_dense_labels_and_empty():
    g = torch.Generator().manual_seed(0)
    low = torch.randn((2, 19, 16, 24), generator=g)
    n_all = 2 * 64 * 96
    pi = torch.arange(n_all, dtype=torch.int32) // (64 * 96)
    px = torch.arange(n_all, dtype=torch.int32) % (64 * 96)
    pl = torch.randint(0, 19, (n_all,), generator=g, dtype=torch.int32)
    loss, grad, _ = _lib.sparse_ce(low.to(DEV), (64, 96), pi.to(DEV), px.to(DEV), pl.to(DEV))  # fully-supervised
    rl, rg, _ = _reference(low, (64, 96), pi, px, pl)
    assert abs(loss.item() - rl.item()) < 2e-5
    assert torch.allclose(grad.cpu(), rg, atol=2e-6, rtol=1e-4)
    e = torch.zeros(0, dtype=torch.int32, device=DEV)
    loss, _, _ = _lib.sparse_ce(low.to(DEV), (64, 96), e, e, e)
    assert torch.isnan(loss).all()  # F.cross_entropy with every target ignored is NaN


@pytest.mark.parametrize("shape,size", [((2, 19, 64, 128), (256, 512)), ((1, 256, 16, 32), (64, 128)),
                                        ((2, 5, 23, 30), (90, 120)), ((1, 3, 1, 1), (16, 32)), ((1, 2, 7, 9), (7, 9))])
def test_upsample_align_corners_fwd_bwd(shape, size):
    g = torch.Generator().manual_seed(1)
    x = torch.randn(shape, generator=g)
    got = _lib.upsample_bilinear_ac(x.to(DEV), size).cpu()
    xr = x.clone().requires_grad_(True)
    want = F.interpolate(xr, size=size, mode="bilinear", align_corners=True)
    assert torch.allclose(got, want.detach(), atol=1e-6, rtol=1e-5)
    go = torch.randn(want.shape, generator=g)
    want.backward(go)
    gin = _lib.upsample_bilinear_ac_bwd(go.to(DEV), shape[2:]).cpu()
    assert torch.allclose(gin, xr.grad, atol=1e-5, rtol=1e-4)


def test_device_metrics_equal_fast_hist():
    """pp_metrics_accumulate == the reference's RunningScore._fast_hist (utils/metrics.py:168-173) on (label, prediction)
    pairs, bit-exact, incl. ignore_index labels, a device-side valid count and accumulation over several launches."""
    import numpy as np
    from pixelpick_b200 import _lib
    from pixelpick_b200.utils import RunningScore
    dev = torch.device("cuda:0")
    C = 19
    rs = np.random.RandomState(0)
    dm = _lib.DeviceMetrics(C, dev)
    ref = RunningScore(C)
    loss_total = 0.0
    for step in range(4):
        n_max = 500
        lt = rs.randint(0, C + 2, size=n_max).astype(np.int32)  # C, C+1 play ignore_index: dropped
        lt[rs.rand(n_max) < 0.05] = 255
        lp = rs.randint(0, C, size=n_max).astype(np.int32)
        n_valid = int(rs.randint(1, n_max + 1)) if step % 2 else n_max
        loss = torch.tensor([0.5 + step], device=dev)
        dm.accumulate(torch.from_numpy(lt).to(dev), torch.from_numpy(lp).to(dev), loss=loss,
                      n_valid=torch.tensor([n_valid], dtype=torch.int32, device=dev) if step % 2 else None)
        ref.update_pairs(np.where(lt[:n_valid] < C, lt[:n_valid], C + 5), lp[:n_valid])
        loss_total += 0.5 + step
    conf, loss_sum, n_steps = dm.read()
    assert n_steps == 4 and abs(loss_sum - loss_total) < 1e-6
    assert np.array_equal(conf, ref.confusion_matrix.astype(np.int64))
    dm.reset()
    assert dm.read()[0].sum() == 0
