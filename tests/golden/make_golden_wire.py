"""Generate golden vectors for the query WIRE FORMAT by running the UNMODIFIED reference here.

    python tests/golden/make_golden_wire.py       # needs /root/reference (build container only)

The reference (NoelShin/PixelPick @ 43c2981) is imported from /root/reference; nothing is copied.  Output:
`tests/golden/wire_golden.pkl`, one small pickle holding

  masks          seeded boolean query masks (ragged sizes) and their image paths
  encoded        QuerySelector.encode_query on each (query.py:72-88), merged into one dict = the `queries.pkl` payload
  encoded_bytes  pickle.dumps(encoded, protocol=4): the byte image of `queries.pkl` for byte-compatibility checks
  decoded_list / decoded_dict / decoded_one
                 QuerySelector.decode_queries (query.py:90-142) as list, as dict, on a single entry
  human          the same entries with a `category_id` list (the annotation tool's output, query.py:97-105), decoded
                 with ignore_index 255 and 19
  merged         merge_previous_query_files (query.py:316-351) over three rounds of `*/queries.pkl` files whose
                 pixels overlap with DIFFERENT labels (later file wins), with ignore_index 255
  human_call     QuerySelector.__call__(nth_query, model, human_labels=True) (query.py:144-222) on the inputs of
                 query_golden.npz's `call_*` case: the masks are `dataset.list_labelled_queries` (int64 label maps, ignore_index
                 where unlabelled), no statistics are written and label_queries is not called; picks per image for
                 margin_sampling and entropy under np.random.seed(0)
  stats          QueryStats.update / save (query.py:250-308) over three images: the inputs (query masks, labels, logits)
                 and the `query_stats.pkl` dict the reference writes
"""
import os
import pickle
import sys
import tempfile

import numpy as np
import torch

REF = "/root/reference"
sys.dont_write_bytecode = True
sys.path.insert(0, REF)
import query as refq  # noqa: E402  (the reference module)

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "wire_golden.pkl")


def main():
    rs = np.random.RandomState(1234)
    sizes = [(12, 20), (7, 9), (16, 16), (1, 5)]
    paths = [f"/data/leftImg8bit/train/city_{i}/im_{i:03d}.png" for i in (3, 0, 2, 1)]  # unsorted on purpose
    masks = []
    for (h, w) in sizes:
        m = np.zeros((h, w), dtype=bool)
        n = min(h * w, 6)
        m.reshape(-1)[rs.choice(h * w, n, replace=False)] = True
        masks.append(m)

    encoded = {}
    for p, m in zip(paths, masks):
        encoded.update(refq.QuerySelector.encode_query(p, m.shape, m))
    decoded_list = refq.QuerySelector.decode_queries(encoded)
    decoded_dict = refq.QuerySelector.decode_queries(encoded, return_as_dict=True)
    decoded_one = refq.QuerySelector.decode_queries({paths[0]: encoded[paths[0]]})

    human = {}
    for p, info in encoded.items():
        d = dict(info)
        d["category_id"] = [int(c) for c in rs.randint(0, 19, size=len(info["x_coords"]))]
        human[p] = d
    human_255 = refq.QuerySelector.decode_queries(human, ignore_index=255, return_as_dict=True)
    human_19 = refq.QuerySelector.decode_queries(human, ignore_index=19)

    # three rounds of human-labelled query files; rounds overlap on some pixels with different labels
    rounds = []
    for r in range(3):
        d = {}
        for p, (h, w) in list(zip(paths, sizes))[: 4 - r]:
            n = min(h * w, 5)
            flat = rs.choice(h * w, n, replace=False)
            if r > 0 and p in rounds[0]:  # force an overlap with round 0
                flat[0] = rounds[0][p]["y_coords"][0] * w + rounds[0][p]["x_coords"][0]
            d[p] = {"height": h, "width": w, "y_coords": flat // w, "x_coords": flat % w,
                    "category_id": [int(c) for c in rs.randint(0, 19, size=n)]}
        rounds.append(d)
    with tempfile.TemporaryDirectory() as tmp:
        files = []
        for r, d in enumerate(rounds):
            os.makedirs(os.path.join(tmp, f"{r}_query"))
            f = os.path.join(tmp, f"{r}_query", "queries.pkl")
            pickle.dump(d, open(f, "wb"))
            files.append(f)
        found = refq.gather_previous_query_files(tmp)
        assert sorted(found) == sorted(files)
        merged = refq.merge_previous_query_files(files, ignore_index=255, verbose=False)

    # QueryStats over three images (C = 11, 24x32)
    from argparse import Namespace
    g = torch.Generator().manual_seed(77)
    st_q, st_y, st_logits = [], [], []
    with tempfile.TemporaryDirectory() as tmp:
        qs = refq.QueryStats(Namespace(dir_root=tmp, experim_name="golden", n_classes=11))
        for i in range(3):
            logits = torch.randn((1, 11, 24, 32), generator=g) * 2.0
            y = rs.randint(0, 11, size=(24, 32)).astype(np.int64)
            qm = np.zeros(24 * 32, dtype=bool)
            qm[rs.choice(24 * 32, 10, replace=False)] = True
            qm = qm.reshape(24, 32)
            qs.update(qm, y, torch.softmax(logits, dim=1))
            st_q.append(qm), st_y.append(y), st_logits.append(logits.numpy())
        qs.save(0)
        st_saved = pickle.load(open(os.path.join(tmp, "checkpoints", "golden", "0_query", "query_stats.pkl"), "rb"))
    stats = {"queries": st_q, "y": st_y, "logits": st_logits, "saved": st_saved, "list_entropy": list(qs.list_entropy)}

    # QuerySelector.__call__ with human labels, on the call_* inputs of query_golden.npz
    qg = np.load(os.path.join(os.path.dirname(OUT), "query_golden.npz"))
    c_logits, c_y, c_lab = torch.from_numpy(qg["call_logits"]), qg["call_y"], qg["call_lab"]

    class HumanDS:
        list_labelled_queries = [np.where(c_lab[i], c_y[i], 19).astype(np.int64) for i in range(3)]
        called = False

        def label_queries(self, *a, **k):
            HumanDS.called = True

    class HumanLoader:
        dataset = HumanDS()

        def __iter__(self):
            for i in range(3):
                yield {"x": c_logits[i:i + 1], "y": torch.from_numpy(c_y[i:i + 1]), "p_img": [f"img_{i:04d}.png"]}

    class PredIsInput:
        def eval(self):
            return self

        def __call__(self, x):
            return {"pred": x}

    human_call = {}
    for strat in ("margin_sampling", "entropy"):
        with tempfile.TemporaryDirectory() as tmp:
            a = Namespace(dataset_name="cs", debug=False, dir_root=tmp, experim_name="golden", ignore_index=19, mc_n_steps=20,
                          n_classes=19, n_pixels_by_us=10, network_name="deeplab", query_strategy=strat, reverse_order=False,
                          stride_total=8, top_n_percent=0.05, use_mc_dropout=False, vote_type="soft")
            sel = refq.QuerySelector(a, HumanLoader(), device=torch.device("cpu"))
            np.random.seed(0)
            d = sel(1, PredIsInput(), human_labels=True)
            assert not HumanDS.called and not os.path.exists(os.path.join(tmp, "checkpoints"))
        human_call[strat] = [np.stack([d[f"img_{i:04d}.png"]["x_coords"], d[f"img_{i:04d}.png"]["y_coords"]]) for i in range(3)]

    out = {"human_call": human_call, "stats": stats, "paths": paths, "masks": masks, "encoded": encoded, "encoded_bytes": pickle.dumps(encoded, protocol=4),
           "decoded_list": decoded_list, "decoded_dict": decoded_dict, "decoded_one": decoded_one, "human": human,
           "human_255": human_255, "human_19": human_19, "rounds": rounds, "merged": merged}
    pickle.dump(out, open(OUT, "wb"), protocol=4)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
