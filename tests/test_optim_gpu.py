"""GPU: pp_adam_step_multi / optim.FusedAdam against torch.optim.Adam (the optimiser utils/utils.py:112-141 builds for `cs`)
on the same tensors: ragged sizes, an unaligned view, two parameter groups with different learning rates / weight decay,
tensor learning rates changed between steps (what the Poly scheduler does), state-dict round trip, graph capture."""
import copy

import pytest
import torch

from pixelpick_b200.optim import FusedAdam

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")
SIZES = [(1,), (3,), (19, 304, 1, 1), (256, 2048, 3, 3), (17,), (16385,), (64, 3, 7, 7), (4099,), (2048,), (5, 16384)]


def _params(seed):
    g = torch.Generator().manual_seed(seed)
    ps = [torch.randn(s, generator=g).to(DEV).requires_grad_(True) for s in SIZES]
    # a parameter that starts 4 bytes off a 16-byte boundary (the scalar path)
    store = torch.randn(1 + 1001, generator=g).to(DEV)
    ps.append(store[1:].detach().requires_grad_(True))
    return ps


def _groups(ps, tensor_lr):
    mk = (lambda v: torch.tensor(v, dtype=torch.float32, device=DEV)) if tensor_lr else (lambda v: v)
    return [{"params": ps[:4], "lr": mk(5e-5), "weight_decay": 2e-4},
            {"params": ps[4:], "lr": mk(5e-4), "weight_decay": 0.0, "betas": (0.8, 0.99), "eps": 1e-7}]


def _set_grads(ps, seed):
    g = torch.Generator().manual_seed(seed)
    for p in ps:
        p.grad = (torch.randn(p.shape, generator=g) * 0.1).to(DEV)


def test_equals_torch_adam_over_steps_and_lr_changes():
    pa, pb = _params(1), _params(1)
    ours = FusedAdam(_groups(pa, True))
    ref = torch.optim.Adam(_groups(pb, True), fused=True, capturable=True)
    for k in range(6):
        _set_grads(pa, 10 + k)
        _set_grads(pb, 10 + k)
        for o in (ours, ref):
            for g in o.param_groups:
                g["lr"].fill_(float(g["lr"]) * 0.9)  # a scheduler rewriting the device scalars
        ours.step()
        ref.step()
    for a, b in zip(pa, pb):
        assert torch.allclose(a, b, rtol=1e-6, atol=1e-8), (a.shape, (a - b).abs().max().item())
        sa, sb = ours.state[a], ref.state[b]
        assert float(sa["step"]) == float(sb["step"]) == 6.0
        assert torch.allclose(sa["exp_avg"], sb["exp_avg"], rtol=1e-6, atol=1e-10)
        assert torch.allclose(sa["exp_avg_sq"], sb["exp_avg_sq"], rtol=1e-6, atol=1e-12)


def test_state_dict_round_trip_and_fallback():
    pa, pb = _params(2), _params(2)
    ours = FusedAdam(_groups(pa, True))
    for k in range(2):
        _set_grads(pa, 20 + k)
        ours.step()
    sd = copy.deepcopy(ours.state_dict())
    assert set(sd["state"][0]) == {"step", "exp_avg", "exp_avg_sq"}  # torch.optim.Adam's layout
    with torch.no_grad():
        for a, b in zip(pa, pb):
            b.copy_(a)
    other = FusedAdam(_groups(pb, True))
    other.load_state_dict(sd)
    _set_grads(pa, 30)
    _set_grads(pb, 30)
    ours.step()
    other.step()  # the loaded step counters move into the shared buffer
    for a, b in zip(pa, pb):
        assert torch.equal(a, b)
        assert float(other.state[b]["step"]) == 3.0
    # a parameter without a gradient: torch's own step runs (and skips it), the state stays interchangeable
    _set_grads(pa, 31)
    _set_grads(pb, 31)
    pa[2].grad = None
    pb[2].grad = None
    before = pa[2].detach().clone()
    ours.step()
    other.step()
    assert torch.equal(pa[2], before)
    for a, b in zip(pa, pb):
        assert torch.equal(a, b)


def test_inside_a_captured_graph():
    pa, pb = _params(3), _params(3)
    ours = FusedAdam(_groups(pa, True))
    ref = torch.optim.Adam(_groups(pb, True), fused=True, capturable=True)
    _set_grads(pa, 40)
    _set_grads(pb, 40)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        ours.step()  # warm-up step outside the graph (state creation)
    torch.cuda.current_stream().wait_stream(s)
    ref.step()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        ours.step()
    for k in range(3):
        g = torch.Generator().manual_seed(50 + k)
        for a, b in zip(pa, pb):
            new = (torch.randn(a.shape, generator=g) * 0.1).to(DEV)
            a.grad.copy_(new)  # static gradient buffers, as in the captured train step
            b.grad = new.clone()
        graph.replay()
        ref.step()
    torch.cuda.synchronize()
    for a, b in zip(pa, pb):
        assert torch.allclose(a, b, rtol=1e-6, atol=1e-8)
        assert float(ours.state[a]["step"]) == 4.0
