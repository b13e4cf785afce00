"""CPU: pins oracle/deeplab_oracle.py (functional fp32 restatement) against tests/golden/model_golden.npz, which
was produced by the UNMODIFIED reference modules (tests/golden/make_golden_model.py)."""
import os
import sys
from argparse import Namespace

import numpy as np
import pytest
import torch

from oracle import deeplab_oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ARGS = Namespace(use_mc_dropout=False, mc_dropout_p=0.2, n_classes=19)


@pytest.fixture(scope="module")
def mg():
    return np.load(os.path.join(ROOT, "tests", "golden", "model_golden.npz"))


def _shapes(backbone):
    from pixelpick_b200.deeplab import DeepLab  # parameter names/shapes only (CPU construction, no kernels)
    m = DeepLab(ARGS, backbone=backbone)
    return {k: tuple(v.shape) for k, v in m.state_dict().items()}


def _x(seed, shape):
    return torch.randn(shape, generator=torch.Generator().manual_seed(seed))


def test_mobilenet_deeplab_eval_matches_reference(mg):
    sd = orc.synthetic_state_dict(_shapes("mobilenet"), seed=1)
    with torch.no_grad():
        out = orc.deeplab_forward(sd, _x(10, (1, 3, 48, 64)))
    assert np.allclose(out["pred"].numpy(), mg["mnv2_eval_pred"], atol=1e-5, rtol=1e-5)


def test_rn50_composition_eval_matches_reference(mg):
    sd = orc.synthetic_state_dict(_shapes("resnet"), seed=2)
    with torch.no_grad():
        out = orc.deeplab_forward(sd, _x(12, (1, 3, 64, 96)), backbone="resnet")
    assert np.allclose(out["pred"].numpy(), mg["rn50_eval_pred"], atol=2e-5, rtol=1e-4)
    assert sum(v.numel() for k, v in sd.items() if not k.endswith(("running_mean", "running_var", "num_batches_tracked"))) \
        == int(mg["rn50_n_params"]) == 40351667


def test_train_step_loss_grads_and_running_stats(mg):
    shapes = _shapes("mobilenet")
    sd = orc.synthetic_state_dict(shapes, seed=1)
    # the reference registers features / low_level_features / high_level_features as aliases of the same tensors
    params = {}
    for k in sorted(sd):
        if k.endswith(("running_mean", "running_var", "num_batches_tracked")):
            continue
        if k.startswith(("backbone.low_level_features.", "backbone.high_level_features.")):
            continue
        sd[k] = sd[k].clone().requires_grad_(True)
        params[k] = sd[k]
    x = _x(11, (2, 3, 64, 64))
    rs = np.random.RandomState(11)
    y = torch.from_numpy(rs.randint(0, 19, size=(2, 64, 64)).astype(np.int64))
    q = torch.zeros((2, 64 * 64), dtype=torch.bool)
    for i in range(2):
        q[i, torch.from_numpy(rs.choice(64 * 64, 10, replace=False))] = True
    out = orc.deeplab_forward(sd, x, training=True, return_ctx=True)
    loss = orc.sparse_ce_loss(out["pred"], y, q.view(2, 64, 64), 19)
    loss.backward()
    assert np.allclose(out["pred"].detach().numpy(), mg["mnv2_train_pred"], atol=2e-5, rtol=1e-4)
    assert abs(loss.item() - float(mg["mnv2_train_loss"])) < 1e-5
    names, norms = list(mg["mnv2_train_grad_names"]), mg["mnv2_train_grad_norms"]
    assert set(names) == set(params)
    for n, ref in zip(names, norms):
        got = params[n].grad.norm().item()
        assert abs(got - ref) <= 1e-3 * max(ref, 1e-6) + 1e-7, (n, got, ref)
    for key in mg.files:
        if key.startswith("mnv2_train_grad::"):
            n = key.split("::")[1]
            g = params[n.split("[")[0]].grad
            if n.endswith("[:8]"):
                g = g[:8]
            elif n.endswith("[:4]"):
                g = g[:4]
            assert np.allclose(g.numpy(), mg[key], atol=1e-6 + 1e-4 * np.abs(mg[key]).max(), rtol=1e-3), n
    ctx = out["ctx"]
    assert np.allclose(ctx.new_stats["aspp.bn1"][0].numpy(), mg["mnv2_train_running_mean::aspp.bn1"], atol=1e-6)
    assert np.allclose(ctx.new_stats["seg_head.segment_head.1"][1].numpy(),
                       mg["mnv2_train_running_var::seg_head.segment_head.1"], atol=1e-6, rtol=1e-5)


@pytest.mark.parametrize("backbone", ["mobilenet", "resnet"])
def test_state_dict_layout_equals_the_reference(backbone):
    """A checkpoint of the reference (`{"model": model.state_dict()}`, model.py:207-212) loads into the drop-in with strict=True
    and vice versa: same keys in the same order, same shapes and dtypes (tests/golden/state_dict_keys.json, written from the
    reference modules by tests/golden/make_golden_keys.py)."""
    import json
    from pixelpick_b200.deeplab import DeepLab
    want = json.load(open(os.path.join(ROOT, "tests", "golden", "state_dict_keys.json")))[backbone]
    sd = DeepLab(ARGS, backbone=backbone).state_dict()
    got = [[k, list(v.shape), str(v.dtype).replace("torch.", "")] for k, v in sd.items()]
    assert [g[0] for g in got] == [w[0] for w in want]
    assert got == want


def test_seeded_construction_starts_from_the_reference_weights():
    """DeepLab(args) built right after torch.manual_seed(0) holds, bit for bit, the tensors the reference's DeepLab holds after
    the same seed (aspp.py:14,62 draws every ASPP branch twice; deeplab.py:23-26 leaves low_level_conv at nn.Conv2d's default):
    a seeded run of the drop-in starts where the reference starts.  Exact checksums of the raw bit patterns from
    tests/golden/make_golden_keys.py; skipped where this host's torch CPU normal stream differs from the generating host's."""
    import json
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    from pixelpick_b200.deeplab import DeepLab
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "state_dict_keys.json")))

    def bits_checksum(t):
        a = t.detach().contiguous().cpu().numpy()
        a = a.view(np.uint32) if a.dtype == np.float32 else a.astype(np.int64).view(np.uint64)
        return int(a.astype(np.uint64).sum() % (1 << 63))

    torch.manual_seed(123)
    w = torch.empty(64, 32, 3, 3)
    torch.nn.init.kaiming_normal_(w)
    if [bits_checksum(torch.randn(4096)), bits_checksum(w)] != gold["canary"]:
        pytest.skip("torch's CPU normal stream on this host differs from the host that generated the golden")
    torch.manual_seed(0)
    got = {k: bits_checksum(v) for k, v in DeepLab(ARGS).state_dict().items()}
    diff = [k for k in gold["init_seed0"] if got.get(k) != gold["init_seed0"][k]]
    assert not diff, diff[:10]
