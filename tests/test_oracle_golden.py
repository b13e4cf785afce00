"""CPU: pins oracle/acq_oracle.py against golden vectors produced by the UNMODIFIED reference
(tests/golden/make_golden.py imports /root/reference/query.py in the build container)."""
import numpy as np
import pytest
import torch

from oracle import acq_oracle as orc

STRATS = ["entropy", "least_confidence", "margin_sampling"]
CASES = [11, 19, 21]


def _t(a):
    """torch-allocated (64-byte aligned) copy: torch's vectorised CPU kernels peel unaligned heads with scalar code, so
    bit-exact comparisons must not depend on NumPy's 16-byte malloc alignment of the .npz arrays."""
    return torch.from_numpy(a).clone()


@pytest.mark.parametrize("C", CASES)
@pytest.mark.parametrize("strat", STRATS)
def test_scores_bit_exact(golden, C, strat):
    logits = _t(golden[f"logits_c{C}"])
    got = orc.uncertainty(orc.probabilities(logits), strat).numpy()
    assert np.array_equal(got, golden[f"scores_{strat}_c{C}"], equal_nan=True)


@pytest.mark.parametrize("C", CASES)
@pytest.mark.parametrize("strat", STRATS)
def test_masked_scores_bit_exact(golden, C, strat):
    logits = _t(golden[f"logits_c{C}"])
    y, lab = golden[f"y_c{C}"], golden[f"lab_c{C}"]
    uc = orc.uncertainty(orc.probabilities(logits), strat)  # same batch shape as the golden run
    for i in range(2):
        got = orc.apply_masks(uc[i], strat, lab[i], y[i] == C).numpy()
        assert np.array_equal(got, golden[f"uc_{strat}_c{C}"][i], equal_nan=True)
        # torch's CPU softmax is not bit-stable across batch shapes: per-image evaluation is only close
        one = orc.score_map(logits[i:i + 1], strat, lab[i], y[i] == C).numpy()
        assert np.allclose(one, got, rtol=1e-6, atol=1e-7, equal_nan=True)


@pytest.mark.parametrize("C", CASES)
@pytest.mark.parametrize("strat", STRATS)
@pytest.mark.parametrize("tag,kw", [("top5", dict(top_n_percent=0.05)), ("topn", dict(top_n_percent=0.0)),
                                    ("rev", dict(top_n_percent=0.05, reverse_order=True))])
def test_select_queries_matches_reference(golden, C, strat, tag, kw):
    uc = _t(golden[f"uc_{strat}_c{C}"])
    for topk in (orc.topk_indices_torch, orc.topk_indices_spec):
        np.random.seed(7)
        sel = np.stack([orc.select_queries(uc[i], strat, 10, topk=topk, **kw) for i in range(2)])
        assert np.array_equal(sel, golden[f"sel_{tag}_{strat}_c{C}"]), topk.__name__


def test_entropy_nan_ranks_first(golden):
    logits = _t(golden["logits_nan"])
    uc = orc.uncertainty(orc.probabilities(logits), "entropy")
    assert np.array_equal(uc.numpy(), golden["scores_entropy_nan"], equal_nan=True)
    assert np.isnan(uc.numpy()).sum() >= 2
    for topk in (orc.topk_indices_torch, orc.topk_indices_spec):
        sel = orc.select_queries(uc[0], "entropy", 4, 0.0, topk=topk)
        assert np.array_equal(sel, golden["sel_nan"])
        assert sel[2, 5] and sel[7, 1]


@pytest.mark.parametrize("strat", STRATS)
def test_query_call_end_to_end(golden, strat):
    logits = _t(golden["call_logits"])
    y, lab = golden["call_y"], golden["call_lab"]
    np.random.seed(0)
    d = orc.query_images([logits[i:i + 1] for i in range(3)], strat, lab, y == 19,
                         [f"img_{i:04d}.png" for i in range(3)])
    for i, (p, info) in enumerate(sorted(d.items())):
        assert np.array_equal(np.stack([info["x_coords"], info["y_coords"]]), golden[f"call_{strat}_{i}_xy"])


@pytest.mark.parametrize("strat", STRATS)
def test_full_size_image(golden, strat):
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
    uc = orc.score_map(logits, strat, lab.reshape(1, 256, 512)[0], y[0] == 19)
    k = int(256 * 512 * 0.05)
    largest = orc.LARGEST[strat]
    vals = golden[f"big_topk_val_{strat}"]
    idx_t = orc.topk_indices_torch(uc.flatten(), k, largest)
    assert np.array_equal(idx_t, golden[f"big_topk_idx_{strat}"])
    idx_s = orc.topk_indices_spec(uc.flatten(), k, largest)
    assert np.array_equal(uc.flatten().numpy()[idx_s], vals)  # same value sequence
    distinct = np.ones(k, dtype=bool)
    distinct[1:] &= vals[1:] != vals[:-1]
    distinct[:-1] &= vals[:-1] != vals[1:]
    assert np.array_equal(idx_s[distinct], idx_t[distinct])  # identical outside tie groups
    np.random.seed(3)
    sel = np.flatnonzero(orc.select_queries(uc, strat, 10, 0.05))
    assert np.array_equal(sel, golden[f"big_sel_{strat}"])


def test_ord_key_properties():
    s = np.array([np.nan, np.inf, 3.0, 1e-38, 0.0, -0.0, -1e-38, -2.0, -np.inf], dtype=np.float32)
    k = orc.ord_key(s, largest=True)
    assert k[4] == k[5]
    assert list(np.argsort(k, kind="stable")) == list(range(9))
    k2 = orc.ord_key(s, largest=False)
    assert list(np.argsort(k2, kind="stable")) == [8, 7, 6, 4, 5, 3, 2, 1, 0]
