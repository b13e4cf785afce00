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


@pytest.mark.parametrize("C,lr,size,dtype", [(19, (16, 32), (64, 128), torch.int64), (11, (23, 30), (90, 120), torch.int32),
                                             (21, (10, 13), (40, 52), torch.uint8), (19, (64, 128), (256, 512), torch.int64)])
def test_eval_confusion_fused_upsample_argmax(C, lr, size, dtype):
    """pp_eval_confusion_upsampled == RunningScore.update(y, argmax(F.interpolate(logits, size, bilinear, align_corners=True)))
    (model.py:177-239 / eval.py:44-62): bit-exact on logits quantised to 2^-6 (ties resolved like torch.argmax: first
    maximum), <= 1e-4 of the pixels off on unconstrained fp32 logits (last-ulp differences of the interpolation)."""
    import numpy as np
    from pixelpick_b200 import _lib
    from pixelpick_b200.utils import RunningScore
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(C + lr[0])
    n = 3
    ignore = 255 if dtype == torch.uint8 else C
    y = torch.randint(0, C, (n,) + size, generator=g)
    y[torch.rand((n,) + size, generator=g) < 0.05] = ignore
    for quantised in (True, False):
        logits = torch.randn((n, C) + lr, generator=g) * 3
        if quantised:
            logits = torch.round(logits * 4) / 4  # coarse alphabet: exact ties after interpolation are common
        full = F.interpolate(logits, size=size, mode="bilinear", align_corners=True)
        ref = RunningScore(C)
        ref.update(y.numpy(), full.argmax(dim=1).numpy())
        conf = torch.zeros((C, C), dtype=torch.int64, device=dev)
        pred = _lib.eval_confusion_upsampled(logits.to(dev), size, y.to(dtype).to(dev), conf, want_pred=True)
        got = conf.cpu().numpy()
        assert got.sum() == int((y != ignore).sum())
        mism = (pred.cpu().long() != full.argmax(dim=1)).float().mean().item()
        if quantised:
            # the interpolation weights are multiples of 1/(out-1): products are not exact, so compare through the
            # mismatch rate as well, but ties on identical neighbours must resolve to the FIRST maximum
            assert mism < 1e-3, mism
        else:
            assert mism < 1e-4, mism
        assert np.abs(got - ref.confusion_matrix).sum() <= 2 * mism * y.numel() + 1e-9
    # accumulation: a second call adds to the same matrix
    before = conf.clone()
    _lib.eval_confusion_upsampled(logits.to(dev), size, y.to(dtype).to(dev), conf)
    assert torch.equal(conf, 2 * before)


def test_evaluate_mirror_matches_full_resolution_path(tmp_path):
    """pixelpick_b200.eval.evaluate (micro-batched, fused kernel) == the reference-style loop (one image, full-res pred,
    host RunningScore) on a synthetic validation set; also the CLI entry point."""
    import numpy as np
    from pixelpick_b200.args import Arguments
    from pixelpick_b200.eval import evaluate, main
    from pixelpick_b200.utils import RunningScore, get_dataloader, get_model
    from copy import deepcopy
    args = Arguments().parse_args(argv=["--dataset_name", "cs", "--dir_root", str(tmp_path), "--n_workers", "0",
                                        "--synthetic", "6", "64", "128"])
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    model = get_model(args).to(dev).eval()
    dl = get_dataloader(deepcopy(args), val=True, query=False, shuffle=False, batch_size=1, n_workers=0)
    ref = RunningScore(args.n_classes)
    with torch.no_grad():
        for d in dl:
            pred = model(d["x"].to(dev))["pred"].argmax(dim=1)
            ref.update(d["y"].numpy(), pred.cpu().numpy())
    want = ref.get_scores()[0]["Mean IoU"]
    got = evaluate(model, dl, "test", epoch=1, dir_ckpt=str(tmp_path / "ck"), device=dev, batch_imgs=4)
    assert abs(got - want) < 2e-3, (got, want)
    assert (tmp_path / "ck" / "e01" / "val" / "log_val.txt").exists()
    sd = tmp_path / "m.pt"
    torch.save({"model": model.state_dict()}, sd)
    got_cli = main(["--dataset_name", "cs", "--dir_root", str(tmp_path), "--n_workers", "0", "--synthetic", "6", "64", "128",
                    "--p_state_dict", str(sd)])
    assert abs(got_cli - want) < 2e-3
