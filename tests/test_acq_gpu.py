"""GPU parity tests of the Q path: CUDA (through the C-ABI) vs the oracle and the reference goldens.

Tolerances (SURVEY.md §8c contract):
  scores     : |gpu - oracle| <= 2e-6 + 1e-5 * |oracle|   (fp32 logits; NaN positions identical)
  top-k      : BIT-EXACT ordered index list vs the order contract (`topk_indices_spec`) evaluated on
               the GPU-produced score map; equal to the reference's torch.topk outside tie groups
  selection  : BIT-EXACT selected pixels under the same np.random seed
"""
import numpy as np
import pytest
import torch

from oracle import acq_oracle as orc
from pixelpick_b200 import _lib

pytestmark = pytest.mark.gpu
STRATS = ["entropy", "least_confidence", "margin_sampling"]
ATOL, RTOL = 2e-6, 1e-5
DEV = torch.device("cuda:0")


def _close(a, b):
    return np.allclose(a, b, atol=ATOL, rtol=RTOL, equal_nan=True)


def _rand_logits(seed, n, C, h, w, scale=3.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn((n, C, h, w), generator=g) * scale).float()


def _masks(seed, n, h, w, n_lab=10, void_frac=0.02):
    rs = np.random.RandomState(seed)
    lab = np.zeros((n, h * w), dtype=bool)
    for i in range(n):
        lab[i, rs.choice(h * w, min(n_lab, h * w), replace=False)] = True
    void = rs.rand(n, h, w) < void_frac
    return lab.reshape(n, h, w), void


# ------------------------------------------------------------------------------------------ scores
@pytest.mark.parametrize("C", [11, 19, 21])
@pytest.mark.parametrize("strat", STRATS)
def test_scores_vs_reference_golden(golden, C, strat):
    logits = torch.from_numpy(golden[f"logits_c{C}"])
    got = _lib.acq_score(logits.to(DEV), strat).cpu().numpy()
    assert _close(got, golden[f"scores_{strat}_c{C}"])
    y, lab = golden[f"y_c{C}"], golden[f"lab_c{C}"]
    got = _lib.acq_score(logits.to(DEV), strat, torch.from_numpy(lab).to(DEV),
                         torch.from_numpy(y == C).to(DEV)).cpu().numpy()
    want = golden[f"uc_{strat}_c{C}"]
    assert _close(got, want)
    excl = lab | (y == C)
    assert np.all(got[excl] == orc.FILL[strat])  # fills are exact


def test_entropy_nan_semantics(golden):
    logits = torch.from_numpy(golden["logits_nan"])
    got = _lib.acq_score(logits.to(DEV), "entropy").cpu().numpy()
    want = golden["scores_entropy_nan"]
    assert np.array_equal(np.isnan(got), np.isnan(want))
    assert _close(got, want)
    idx = _lib.acq_topk(torch.from_numpy(got).to(DEV).view(1, -1), 4, True).cpu().numpy()[0]
    sel = np.zeros(8 * 16, dtype=bool)
    sel[idx] = True
    assert np.array_equal(sel.reshape(8, 16), golden["sel_nan"])
    assert set(idx[:2].tolist()) == {2 * 16 + 5, 7 * 16 + 1} and idx[0] < idx[1]  # NaNs first, by index


@pytest.mark.parametrize("strat", STRATS)
@pytest.mark.parametrize("shape", [(1, 7, 23, 37), (2, 19, 17, 30), (3, 5, 8, 8), (1, 33, 9, 12)])
def test_scores_scalar_fallback_shapes(strat, shape):
    n, C, h, w = shape
    logits = _rand_logits(11, n, C, h, w)
    lab, void = _masks(3, n, h, w, n_lab=5)
    got = _lib.acq_score(logits.to(DEV), strat, torch.from_numpy(lab).to(DEV), torch.from_numpy(void).to(DEV))
    for i in range(n):
        want = orc.score_map(logits[i:i + 1], strat, lab[i], void[i]).numpy()
        assert _close(got[i].cpu().numpy(), want)


@pytest.mark.parametrize("strat", STRATS)
def test_scores_strided_view_of_padded_forward(strat):
    # VOC path: model runs on a reflect-padded image and the logits are sliced [:, :, :h, :w] (query.py:171-190)
    full = _rand_logits(5, 2, 21, 40, 48)
    view = full.to(DEV)[:, :, :36, :44]
    got = _lib.acq_score(view, strat).cpu().numpy()
    for i in range(2):
        want = orc.score_map(full[i:i + 1, :, :36, :44], strat).numpy()
        assert _close(got[i], want)
    view2 = full.to(DEV)[:, :, :35, :41]  # unaligned width -> scalar kernel on a strided view
    got = _lib.acq_score(view2, strat).cpu().numpy()
    for i in range(2):
        assert _close(got[i], orc.score_map(full[i:i + 1, :, :35, :41], strat).numpy())


@pytest.mark.parametrize("strat", STRATS)
def test_scores_bf16_logits(strat):
    logits = _rand_logits(21, 2, 19, 32, 64).to(torch.bfloat16)
    got = _lib.acq_score(logits.to(DEV), strat).cpu().numpy()
    for i in range(2):
        assert _close(got[i], orc.score_map(logits[i:i + 1].float(), strat).numpy())


@pytest.mark.parametrize("strat", STRATS)
def test_keep_mask_is_reverse_order_sampling(strat):
    logits = _rand_logits(8, 1, 19, 16, 32)
    rs = np.random.RandomState(1)
    keep = rs.rand(1, 16, 32) < 0.05
    got = _lib.acq_score(logits.to(DEV), strat, keep=torch.from_numpy(keep).to(DEV)).cpu().numpy()[0]
    want = orc.score_map(logits, strat).numpy()
    want[~keep[0]] = orc.FILL[strat]
    assert _close(got, want)


@pytest.mark.parametrize("strat", STRATS)
@pytest.mark.parametrize("C,lr,size", [(19, (16, 32), (64, 128)), (11, (23, 30), (90, 120)), (21, (10, 13), (40, 52))])
def test_scores_fused_upsample(strat, C, lr, size):
    low = _rand_logits(2, 2, C, *lr)
    lab, void = _masks(4, 2, *size)
    got = _lib.acq_score_upsampled(low.to(DEV), size, strat, torch.from_numpy(lab).to(DEV),
                                   torch.from_numpy(void).to(DEV)).cpu().numpy()
    full = torch.nn.functional.interpolate(low, size=size, mode="bilinear", align_corners=True)  # deeplab.py:55
    for i in range(2):
        want = orc.score_map(full[i:i + 1], strat, lab[i], void[i]).numpy()
        # the interpolation itself is re-associated (fma) on the GPU: logits agree to ~1e-6 abs
        assert np.allclose(got[i], want, atol=1e-5, rtol=1e-4, equal_nan=True)


# ------------------------------------------------------------------------------------------- top-k
def _check_topk(scores: torch.Tensor, k, largest):
    n = scores.shape[0]
    idx, val = _lib.acq_topk(scores.to(DEV), k, largest, return_values=True)
    idx, val = idx.cpu().numpy(), val.cpu().numpy()
    for i in range(n):
        want = orc.topk_indices_spec(scores[i], k, largest)
        assert np.array_equal(idx[i], want), f"image {i}: first mismatch at {np.flatnonzero(idx[i] != want)[:5]}"
        assert np.array_equal(val[i], scores[i].numpy().reshape(-1)[want] + np.float32(0), equal_nan=True)
    return idx


@pytest.mark.parametrize("largest", [True, False])
@pytest.mark.parametrize("n,hw,k", [(3, 768, 38), (2, 131072, 6553), (1, 172800, 8640), (2, 1000, 1), (2, 999, 999),
                                     (1, 5000, 4097), (1, 2097152, 104857), (5, 4096, 33)])
def test_topk_random_scores(n, hw, k, largest):
    g = torch.Generator().manual_seed(hw + k)
    _check_topk(torch.rand((n, hw), generator=g), k, largest)


@pytest.mark.parametrize("largest", [True, False])
def test_topk_heavy_ties_and_specials(largest):
    g = torch.Generator().manual_seed(1)
    s = (torch.randint(0, 7, (2, 20000), generator=g).float() / 8.0)  # 7 distinct values -> long tie groups
    _check_topk(s, 1000, largest)
    _check_topk(torch.full((1, 9000), 0.25), 500, largest)  # everything ties: the index levels decide
    s = torch.rand((1, 6000), generator=g)
    s[0, [5, 77, 4000]] = float("nan")
    s[0, [9, 10]] = float("inf")
    s[0, [11]] = -float("inf")
    s[0, [100, 200]] = 0.0
    s[0, [150]] = -0.0
    _check_topk(s, 300, largest)
    _check_topk(s, 6000, largest)


def test_topk_narrow_distribution_exercises_all_levels():
    # near-uniform softmax (untrained net): every entropy shares the top radix digits
    g = torch.Generator().manual_seed(2)
    s = 2.944 + torch.rand((2, 131072), generator=g) * 1e-4
    _check_topk(s, 6553, True)
    s = torch.rand((1, 131072), generator=g).mul(1e-6)
    _check_topk(s, 6553, False)


@pytest.mark.parametrize("strat", STRATS)
def test_fused_hist0_equals_standalone(strat):
    logits = _rand_logits(3, 4, 19, 64, 128).to(DEV)
    k = 409
    ws = _lib.TopKWorkspace(4, 64 * 128, k, DEV)
    ws.prepare()
    score = _lib.acq_score(logits, strat, hist0_ws=ws)
    a = _lib.acq_topk(score.view(4, -1), k, _lib.LARGEST[strat], ws=ws, hist0_valid=True)
    b = _lib.acq_topk(score.view(4, -1), k, _lib.LARGEST[strat])
    assert torch.equal(a, b)
    ws.prepare()  # workspace reuse for the next batch
    score = _lib.acq_score(logits, strat, hist0_ws=ws)
    c = _lib.acq_topk(score.view(4, -1), k, _lib.LARGEST[strat], ws=ws, hist0_valid=True)
    assert torch.equal(a, c)


@pytest.mark.parametrize("strat", STRATS)
def test_workspace_is_prepared_again_after_a_select(strat):
    """pp_acq_topk_prepare contract: a completed select zeroes what it consumed, so the wrapper skips the memset between
    batches.  Batches of different data, sizes and paths (heavy ties -> whole-bucket selections, short boundary lists
    ranked directly, picks with and without the sort) through ONE workspace must equal fresh workspaces, and the zeroed
    histograms must read back as zeros.  (The bucket state a select leaves behind sits where a LARGER batch keeps
    histograms, so the wrapper zeroes again when the batch size changes.)"""
    g = torch.Generator().manual_seed(17)
    h, w, k = 64, 128, 409
    ws = _lib.TopKWorkspace(4, h * w, k, DEV)
    largest = _lib.LARGEST[strat]
    for step in range(5):
        n = 4 if step != 2 else 3  # a smaller batch in between
        logits = torch.randn((n, 19, h, w), generator=g) * (3.0 if step != 3 else 0.0)  # step 3: every score equal
        logits = logits.to(DEV)
        ws.prepare()  # a no-op from the second batch on
        assert ws._clean is not None
        score = _lib.acq_score(logits, strat, hist0_ws=ws)
        assert ws._clean is None
        pos = torch.from_numpy(np.stack([np.random.RandomState(step * 7 + i).permutation(k)[:10] for i in range(n)]).astype(np.int32))
        if step % 2 == 0:
            got = _lib.acq_select_pick(score.view(n, -1), k, largest, pos, ws=ws, hist0_valid=True)
            want = _lib.acq_gather(_lib.acq_topk(score.view(n, -1), k, largest), pos)
        else:
            got = _lib.acq_topk(score.view(n, -1), k, largest, ws=ws, hist0_valid=True)
            want = _lib.acq_topk(score.view(n, -1), k, largest)
        assert ws._clean == n
        assert torch.equal(got, want), step
        torch.cuda.synchronize()
        assert int(ws.buf[: n * 2048 * 4].count_nonzero()) == 0, step  # the level-0 histograms of the batch
    # a score that is not followed by a select leaves the histogram filled: the next fill prepares by itself
    _lib.acq_score(logits, strat, hist0_ws=ws)
    score = _lib.acq_score(logits, strat, hist0_ws=ws)
    got = _lib.acq_topk(score.view(n, -1), k, largest, ws=ws, hist0_valid=True)
    assert torch.equal(got, _lib.acq_topk(score.view(n, -1), k, largest))


# --------------------------------------------------------------------------------------- selection
@pytest.mark.parametrize("strat", STRATS)
@pytest.mark.parametrize("C,h,w", [(19, 256, 512), (11, 360, 480)])
def test_selection_bit_exact_on_gpu_score_map(strat, C, h, w):
    """§8c contract step 2: same (GPU) score map -> reference-style selection on CPU == GPU selection."""
    n = 2
    logits = _rand_logits(40 + C, n, C, h, w)
    lab, void = _masks(6, n, h, w)
    score = _lib.acq_score(logits.to(DEV), strat, torch.from_numpy(lab).to(DEV), torch.from_numpy(void).to(DEV))
    k = int(h * w * 0.05)
    topk = _lib.acq_topk(score.view(n, -1), k, _lib.LARGEST[strat])
    np.random.seed(11)
    pos = np.stack([np.random.permutation(k)[:10] for _ in range(n)]).astype(np.int32)
    sel = _lib.acq_gather(topk, torch.from_numpy(pos)).cpu().numpy()
    score_h = score.cpu()
    np.random.seed(11)
    for i in range(n):
        vals = score_h[i].flatten()[torch.from_numpy(orc.topk_indices_spec(score_h[i].flatten(), k + 1, _lib.LARGEST[strat]))]
        topk_fn = orc.topk_indices_torch if len(np.unique(vals.numpy())) == k + 1 else orc.topk_indices_spec
        want = orc.select_queries(score_h[i], strat, 10, 0.05, topk=topk_fn)
        got = np.zeros(h * w, dtype=bool)
        got[sel[i]] = True
        assert np.array_equal(got.reshape(h, w), want)


@pytest.mark.parametrize("strat", STRATS)
def test_full_size_image_vs_reference_golden(golden, strat):
    """Reference output on a 256x512 image (its own CPU scores): value sequences agree to tolerance and the
    ordered index lists agree except where CPU/GPU rounding swaps near-equal neighbours."""
    # NumPy's legacy RandomState is bit-identical on every platform (tests/golden/make_golden.py)
    logits = torch.from_numpy((np.random.RandomState(900).standard_normal((1, 19, 256, 512)) * 3.0).astype(np.float32))
    bits = logits.numpy().view(np.uint32).astype(np.uint64)  # exact, order-independent checksum of the raw bit patterns
    chk = np.array([bits.sum(), (bits * (np.arange(bits.size, dtype=np.uint64).reshape(bits.shape) % 65521)).sum()], dtype=np.uint64)
    assert np.array_equal(chk, golden["big_logits_checksum"]), "golden inputs could not be regenerated"
    rs = np.random.RandomState(900)
    y = rs.randint(0, 19, size=(1, 256, 512)).astype(np.int64)
    y[rs.rand(1, 256, 512) < 0.02] = 19
    lab = np.zeros((1, 256 * 512), dtype=bool)
    lab[0, rs.choice(256 * 512, 10, replace=False)] = True
    score = _lib.acq_score(logits.to(DEV), strat, torch.from_numpy(lab.reshape(1, 256, 512)).to(DEV),
                           torch.from_numpy(y == 19).to(DEV))
    k = int(256 * 512 * 0.05)
    idx, val = _lib.acq_topk(score.view(1, -1), k, _lib.LARGEST[strat], return_values=True)
    idx, val = idx.cpu().numpy()[0], val.cpu().numpy()[0]
    assert _close(val, golden[f"big_topk_val_{strat}"])
    ref_idx = golden[f"big_topk_idx_{strat}"]
    agree = (idx == ref_idx).mean()
    assert agree > 0.98, agree
    assert len(set(idx.tolist()) ^ set(ref_idx.tolist())) <= 4  # only the k-th boundary can differ
    if agree == 1.0:
        np.random.seed(3)
        sel = np.sort(idx[np.random.permutation(k)[:10]])
        assert np.array_equal(sel, golden[f"big_sel_{strat}"])


@pytest.mark.parametrize("strat", STRATS)
def test_full_size_properties(strat):
    """Size-independent properties at the BASELINE sizes: sortedness, threshold consistency, idempotence."""
    for (n, C, h, w) in [(4, 19, 256, 512), (1, 19, 1024, 2048)]:
        logits = _rand_logits(77, n, C, h, w).to(DEV)
        score = _lib.acq_score(logits, strat)
        k = int(h * w * 0.05)
        largest = _lib.LARGEST[strat]
        idx, val = _lib.acq_topk(score.view(n, -1), k, largest, return_values=True)
        sc = score.view(n, -1)
        assert torch.equal(sc.gather(1, idx.long()), val)
        d = val[:, 1:] - val[:, :-1]
        assert bool((d <= 0).all() if largest else (d >= 0).all())
        ties = d == 0
        assert bool((idx[:, 1:][ties] > idx[:, :-1][ties]).all())
        assert all(len(set(r.tolist())) == k for r in idx.cpu().numpy())
        rest = sc.clone()
        rest.scatter_(1, idx.long(), float("-inf") if largest else float("inf"))
        if largest:
            assert bool((rest.max(dim=1).values <= val[:, -1]).all())
        else:
            assert bool((rest.min(dim=1).values >= val[:, -1]).all())
        idx2 = _lib.acq_topk(score.view(n, -1), k, largest)
        assert torch.equal(idx, idx2)


def test_entropy_at_selected_pixels():
    logits = _rand_logits(9, 2, 19, 32, 48)
    px = torch.tensor([[0, 17, 1535], [5, 700, 1000]], dtype=torch.int32)
    got = _lib.acq_entropy_at(logits.to(DEV), px.to(DEV)).cpu().numpy()
    low = _rand_logits(10, 2, 19, 8, 12)
    got_up = _lib.acq_entropy_at_upsampled(low.to(DEV), (32, 48), px.to(DEV)).cpu().numpy()
    full = torch.nn.functional.interpolate(low, size=(32, 48), mode="bilinear", align_corners=True)
    for i in range(2):
        q = np.zeros(32 * 48, dtype=bool)
        q[px[i].numpy()] = True
        assert _close(got[i], np.array(orc.entropy_at(logits[i:i + 1], q)))
        assert np.allclose(got_up[i], np.array(orc.entropy_at(full[i:i + 1], q)), atol=1e-5)


@pytest.mark.parametrize("strat", STRATS)
def test_host_buffer_session_equals_device_api(strat):
    n, C, h, w = 5, 19, 64, 128
    logits = _rand_logits(13, n, C, h, w)
    lab, void = _masks(2, n, h, w)
    k, nsel = int(h * w * 0.05), 10
    np.random.seed(4)
    pos = np.stack([np.random.permutation(k)[:nsel] for _ in range(n)]).astype(np.int32)
    sess = _lib.AcqSession(2, C, h, w, k, nsel)  # chunk of 2 -> 3 chunks, both slots reused
    h_sel = torch.empty((n, nsel), dtype=torch.int32).pin_memory()
    h_topk = torch.empty((n, k), dtype=torch.int32).pin_memory()
    sess.run(logits.pin_memory(), torch.from_numpy(lab).view(torch.uint8).pin_memory(),
             torch.from_numpy(void).view(torch.uint8).pin_memory(), strat, torch.from_numpy(pos), h_sel, h_topk)
    score = _lib.acq_score(logits.to(DEV), strat, torch.from_numpy(lab).to(DEV), torch.from_numpy(void).to(DEV))
    topk = _lib.acq_topk(score.view(n, -1), k, _lib.LARGEST[strat])
    assert torch.equal(h_topk, topk.cpu())
    assert torch.equal(h_sel, _lib.acq_gather(topk, torch.from_numpy(pos)).cpu())
    sess.close()
    # split form: begin (async H2D + score + select) / host draws / finish (pick + D2H); needs all chunks resident
    sess = _lib.AcqSession(3, C, h, w, k, nsel)
    h_sel2 = torch.empty((n, nsel), dtype=torch.int32).pin_memory()
    for _ in range(2):  # twice: the slots and the pending state are reusable
        sess.begin(logits.pin_memory(), torch.from_numpy(lab).view(torch.uint8).pin_memory(),
                   torch.from_numpy(void).view(torch.uint8).pin_memory(), strat)
        sess.finish(torch.from_numpy(pos), h_sel2)
        assert torch.equal(h_sel2, h_sel)
    with pytest.raises(_lib.PixelPickError):
        sess.finish(torch.from_numpy(pos), h_sel2)  # nothing pending
    small = _lib.AcqSession(2, C, h, w, k, nsel)
    with pytest.raises(_lib.PixelPickError):
        small.begin(logits.pin_memory(), None, None, strat)  # 5 images > 2 resident chunks of 2
    small.close()
    sess.close()


def test_errors_are_loud():
    with pytest.raises(_lib.PixelPickError):
        _lib.acq_score(torch.zeros(1, 19, 8, 8), "entropy")  # CPU tensor: no fallback
    with pytest.raises(_lib.PixelPickError):
        _lib.acq_topk(torch.zeros(1, 16, device=DEV), 17, True)  # k > HW


@pytest.mark.parametrize("largest", [True, False])
@pytest.mark.parametrize("n,hw,k,nsel", [(3, 768, 38, 10), (2, 131072, 6553, 10), (1, 172800, 8640, 10), (2, 1000, 1, 1),
                                          (2, 999, 999, 30), (1, 2097152, 104857, 10), (4, 4096, 33, 13)])
def test_pick_ranks_equals_sorted_topk_gather(n, hw, k, nsel, largest):
    """pp_acq_select + pp_acq_pick (order statistics, no sort) == pp_acq_topk + pp_acq_gather, incl. heavy ties."""
    g = torch.Generator().manual_seed(hw + k)
    for scores in (torch.rand((n, hw), generator=g), (torch.randint(0, 5, (n, hw), generator=g).float() / 4.0)):
        rs = np.random.RandomState(k)
        pos = torch.from_numpy(np.stack([rs.permutation(k)[:nsel] for _ in range(n)]).astype(np.int32))
        topk = _lib.acq_topk(scores.to(DEV), k, largest)
        want = _lib.acq_gather(topk, pos)
        got = _lib.acq_select_pick(scores.to(DEV), k, largest, pos)
        assert torch.equal(got, want)
        first = _lib.acq_select_pick(scores.to(DEV), k, largest, None, n=nsel)
        assert torch.equal(first, topk[:, :nsel])


@pytest.mark.parametrize("strat", ["entropy", "least_confidence", "margin_sampling"])
@pytest.mark.parametrize("n,C,H,W", [(5, 19, 256, 512), (3, 11, 64, 128), (2, 21, 128, 256)])
def test_fused_score_select_equals_the_two_call_path(strat, n, C, H, W):
    """pp_acq_score_select (one pass over the logits, scores in shared memory, cluster-merged histogram) leaves the same candidate
    set as pp_acq_score + pp_acq_select: identical picks at random ranks, identical sorted top-k list, identical optional score
    map - incl. masked pixels, exact ties and a NaN score."""
    from pixelpick_b200 import _lib
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(H + C)
    logits = (torch.randn((n, C, H, W), generator=g) * 3).float()
    logits[0, :, :4, :64] = logits[0, :, 4:8, :64]  # exact ties
    logits[1, :, 10, 10] = 0.0                       # uniform pixel: margin 0, entropy ln C
    if strat == "entropy":
        logits[n - 1, 0, 3, 3] = -200.0              # an underflowing class: NaN entropy ranks first
    rs = np.random.RandomState(1)
    lab = torch.from_numpy(rs.rand(n, H, W) < 0.001)
    void = torch.from_numpy(rs.rand(n, H, W) < 0.01)
    k = int(H * W * 0.05)
    assert _lib.acq_score_select_supported(logits.to(dev), C, H, W)
    pos = torch.from_numpy(np.stack([rs.permutation(k)[:10] for _ in range(n)]).astype(np.int32))
    lg, lb, vd = logits.to(dev), lab.to(dev), void.to(dev)
    ws_a = _lib.TopKWorkspace(n, H * W, k, dev)
    ws_a.prepare()
    score = _lib.acq_score(lg, strat, lb, vd, hist0_ws=ws_a)
    want = _lib.acq_select_pick(score.view(n, -1), k, _lib.LARGEST[strat], pos, ws=ws_a, hist0_valid=True)
    ws_b = _lib.TopKWorkspace(n, H * W, k, dev)
    ws_b.prepare()
    score_b = torch.empty_like(score)
    got = _lib.acq_score_select_pick(lg, strat, k, pos, lb, vd, ws=ws_b, score_out=score_b)
    torch.cuda.synchronize()
    assert torch.equal(got.cpu(), want.cpu())
    assert torch.equal(torch.nan_to_num(score_b, nan=-1.0), torch.nan_to_num(score, nan=-1.0))
    # without the optional score map, and the full sorted list out of the fused candidates
    ws_c = _lib.TopKWorkspace(n, H * W, k, dev)
    ws_c.prepare()
    got2 = _lib.acq_score_select_pick(lg, strat, k, None, lb, vd, n=k, ws=ws_c)
    ref_sorted = _lib.acq_topk(score.view(n, -1), k, _lib.LARGEST[strat])
    torch.cuda.synchronize()
    assert torch.equal(got2.cpu(), ref_sorted.cpu())
