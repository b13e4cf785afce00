"""GPU: the dense encoder convolutions on the tcgen05 kernels, training form - forward with the BatchNorm statistics taken in
the conv epilogue (pp_conv_igemm_stats), data gradient, and the tiled / tap-table weight gradient (pp_conv_wgrad_multi) -
against torch fp32 on the same bf16-rounded operands.  Shapes are the ones MobileNetV2 / ResNet-50 produce at 256x512
(ragged channel counts 24, 144, 960; Cout up to 2048; the stride-2 bottleneck through its space-to-depth form)."""
import pytest
import torch
import torch.nn.functional as F

from pixelpick_b200 import _lib
from pixelpick_b200.deeplab import _cpad, s2d_entries, space_to_depth

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")


def _rand(shape, seed, scale=1.0):
    return (torch.randn(shape, generator=torch.Generator().manual_seed(seed)) * scale)


@pytest.mark.parametrize("N,H,W,cin,cout,k,dil", [(2, 64, 128, 64, 256, 1, 1), (2, 32, 64, 512, 2048, 1, 1), (2, 32, 64, 256, 256, 3, 2),
                                                  (2, 66, 130, 24, 144, 1, 1), (2, 16, 32, 960, 320, 1, 1), (1, 20, 36, 144, 24, 1, 1),
                                                  (2, 32, 64, 512, 512, 3, 4), (3, 17, 23, 96, 576, 1, 1), (2, 18, 34, 64, 96, 3, 2),
                                                  (1, 13, 21, 64, 64, 3, 1)])
def test_forward_with_epilogue_statistics(N, H, W, cin, cout, k, dil):
    x = _rand((N, cin, H, W), 1).to(torch.bfloat16)
    w = _rand((cout, cin, k, k), 2, (2.0 / (cin * k * k)) ** 0.5)
    want = F.conv2d(x.float(), w.to(torch.bfloat16).float(), None, 1, dil * (k // 2), dil)
    xn = x.permute(0, 2, 3, 1).contiguous().to(DEV)
    wp, _ = _lib.pack_conv_weights(w.to(DEV), cin, fwd_pad=(_cpad(cout), -(-cin // 64) * 64))
    stats = torch.zeros((2, cout), dtype=torch.float32, device=DEV)
    got = _lib.conv_fused(xn, wp, cout, dil=dil, stats=stats)
    torch.cuda.synchronize()
    g = got.float().permute(0, 3, 1, 2).cpu()
    scale = want.abs().max().item()
    assert (g - want).abs().max().item() < 1e-2 * scale  # bf16 output rounding
    # the statistics are those of the ROUNDED values the kernel wrote
    M = N * H * W
    s_want, q_want = g.sum(dim=(0, 2, 3)), (g * g).sum(dim=(0, 2, 3))
    s_got, q_got = stats[0].cpu(), stats[1].cpu()
    assert torch.allclose(s_got, s_want, rtol=1e-3, atol=1e-3 * (q_want.max().item() * M) ** 0.5)
    assert torch.allclose(q_got, q_want, rtol=1e-3, atol=1e-5 * q_want.max().item())
    # the same launch without statistics, and through the direct register -> global epilogue, must write the same tensor
    assert torch.equal(got, _lib.conv_fused(xn, wp, cout, dil=dil))
    prev = _lib.lib().pp_conv_set_epilogue(0)
    try:
        assert torch.equal(got, _lib.conv_fused(xn, wp, cout, dil=dil))
    finally:
        _lib.lib().pp_conv_set_epilogue(prev)


@pytest.mark.parametrize("N,H,W,cin,cout,k,dil", [(2, 64, 128, 64, 256, 1, 1), (2, 32, 64, 512, 2048, 1, 1), (2, 32, 64, 256, 256, 3, 2),
                                                  (2, 66, 130, 24, 144, 1, 1), (2, 16, 32, 960, 320, 1, 1), (1, 20, 36, 144, 24, 1, 1),
                                                  (2, 32, 64, 128, 128, 3, 1)])
def test_weight_gradient_any_channel_count(N, H, W, cin, cout, k, dil):
    x = _rand((N, cin, H, W), 3).to(torch.bfloat16)
    dy = _rand((N, cout, H, W), 4).to(torch.bfloat16)
    w = torch.zeros((cout, cin, k, k), requires_grad=True)
    F.conv2d(x.float(), w, None, 1, dil * (k // 2), dil).backward(dy.float())
    xn, dyn = x.permute(0, 2, 3, 1).contiguous().to(DEV), dy.permute(0, 2, 3, 1).contiguous().to(DEV)
    ent = [(0, 0, 0)] if k == 1 else [((t // 3 - 1) * dil, (t % 3 - 1) * dil, 0) for t in range(9)]
    dw = _lib.conv_wgrad_multi(xn, cin, dyn, cout, ent)
    got = dw[:, :cin, :cout].permute(2, 1, 0).reshape(cout, cin, k, k).cpu()
    scale = w.grad.abs().max().item()
    assert (got - w.grad).abs().max().item() < 2e-3 * scale


def test_stride_two_bottleneck_convs_through_space_to_depth():
    """resnet_models.py:74-83 at layer2.0: 3x3 stride 2 pad 1 (128 -> 128) and the 1x1 stride 2 downsample."""
    N, H, W, cin, cout = 2, 64, 128, 128, 128
    x = _rand((N, cin, H, W), 5).to(torch.bfloat16)
    w = _rand((cout, cin, 3, 3), 6, (2.0 / (cin * 9)) ** 0.5)
    wr = w.to(torch.bfloat16).float().requires_grad_(True)
    xr = x.float().requires_grad_(True)
    want = F.conv2d(xr, wr, None, 2, 1)
    dy = _rand(tuple(want.shape), 7).to(torch.bfloat16)
    want.backward(dy.float())
    xn = x.permute(0, 2, 3, 1).contiguous().to(DEV)
    xs = space_to_depth(xn)
    assert xs.shape == (N, H // 2, W // 2, 4 * cin)
    ent = s2d_entries(cin)
    wp, _ = _lib.pack_conv_weights(w.to(DEV), cin, fwd_pad=(_cpad(cout), cin))
    got = _lib.conv_fused(xs, wp, cout, entries=ent)
    g = got.float().permute(0, 3, 1, 2).cpu()
    assert (g - want.detach()).abs().max().item() < 1e-2 * want.abs().max().item()
    dyn = dy.permute(0, 2, 3, 1).contiguous().to(DEV)
    dw = _lib.conv_wgrad_multi(xs, cin, dyn, cout, ent)
    gotw = dw[:, :cin, :cout].permute(2, 1, 0).reshape(cout, cin, 3, 3).cpu()
    assert (gotw - wr.grad).abs().max().item() < 2e-3 * wr.grad.abs().max().item()


@pytest.mark.parametrize("N,H,W,cin,cout,k,dil", [(2, 64, 128, 64, 256, 1, 1), (2, 32, 64, 512, 2048, 1, 1), (3, 17, 23, 256, 1024, 1, 1),
                                                  (1, 13, 21, 64, 64, 3, 1)])
def test_residual_added_in_the_raw_epilogue(N, H, W, cin, cout, k, dil):
    """the data gradient that lands on a residual fork (kEpiRawRes): out = bf16(conv + res), one rounding, equal to the
    generic epilogue's result on the direct path"""
    x = _rand((N, cin, H, W), 11).to(torch.bfloat16)
    w = _rand((cout, cin, k, k), 12, (2.0 / (cin * k * k)) ** 0.5)
    res = _rand((N, cout, H, W), 13).to(torch.bfloat16)
    want = F.conv2d(x.float(), w.to(torch.bfloat16).float(), None, 1, dil * (k // 2), dil) + res.float()
    xn = x.permute(0, 2, 3, 1).contiguous().to(DEV)
    rn = res.permute(0, 2, 3, 1).contiguous().to(DEV)
    wp, _ = _lib.pack_conv_weights(w.to(DEV), cin, fwd_pad=(_cpad(cout), -(-cin // 64) * 64))
    got = _lib.conv_fused(xn, wp, cout, dil=dil, res=rn)
    torch.cuda.synchronize()
    g = got.float().permute(0, 3, 1, 2).cpu()
    assert (g - want).abs().max().item() < 1e-2 * want.abs().max().item()
    prev = _lib.lib().pp_conv_set_epilogue(0)
    try:
        assert torch.equal(got, _lib.conv_fused(xn, wp, cout, dil=dil, res=rn))
    finally:
        _lib.lib().pp_conv_set_epilogue(prev)


def test_identity_shortcut_gradient_through_the_link():
    """resnet_models.py:74-94 in train mode: with the shortcut's gradient handed to conv1's dgrad epilogue (_ResLink) the
    input / weight gradients are those of the autograd-add form.  The step's fp32 atomics make two runs of the SAME form
    differ (more so towards the stem: 50 train-mode BatchNorm layers amplify the last bit), so the yardstick is that
    run-to-run floor, measured here, not zero."""
    import pixelpick_b200.deeplab as dl
    from pixelpick_b200.deeplab import ResNet50Dilated8

    torch.manual_seed(3)
    bb = ResNet50Dilated8().to(DEV).train()
    with torch.no_grad():
        for n, p in bb.named_parameters():
            if n.endswith("bn3.weight"):
                p.mul_(0.05)  # the residual branches start small, as in a trained / zero-init-residual network
    x = torch.randn(4, 3, 128, 256, device=DEV)
    grads = []
    for mode in (True, False, False):
        dl._RES_LINK = mode
        try:
            bb.zero_grad(set_to_none=True)
            xi = x.clone().requires_grad_(True)
            t, c2 = dl._rn50_train_forward(bb, xi, torch.bfloat16, {})
            (t.float().square().mean() + c2.float().square().mean()).backward()
            g = {n: p.grad.detach().float().clone() for n, p in bb.named_parameters()}
            g["x"] = xi.grad.detach().float().clone()
            grads.append(g)
        finally:
            dl._RES_LINK = True
    link, ref, ref2 = grads
    worst = 0.0
    for n, g in link.items():
        cos = F.cosine_similarity(g.flatten(), ref[n].flatten(), dim=0).item()
        floor = F.cosine_similarity(ref2[n].flatten(), ref[n].flatten(), dim=0).item()
        worst = max(worst, 1 - cos)
        assert 1 - cos <= 3 * (1 - floor) + 0.02, (n, cos, floor)
        assert abs(g.norm().item() / max(ref[n].norm().item(), 1e-20) - 1) < 0.1 + 3 * abs(ref2[n].norm().item() / max(ref[n].norm().item(), 1e-20) - 1), n
    print("shortcut link: worst 1 - cos over the parameters", worst)
