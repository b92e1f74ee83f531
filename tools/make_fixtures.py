#!/usr/bin/env python
"""Generate tests/golden/ from the reference tree.  Runs only where /root/reference exists.

What it writes (all consumed by tests/ on machines that do NOT have the reference):

* tests/golden/weights/<MODEL>/...   the trained weight blobs, in the reference's own file formats
                                     (inputs the host loads);
* tests/golden/<dataset>.npz          packed graphs: all 4,113 shipped molhiv graphs, the first 4,113
                                     molpcba graphs, the first 1,000 hep10k graphs (with DGN eigenvectors);
* tests/golden/golden_<dataset>.npz   per-graph predictions of the UNMODIFIED reference kernels compiled
                                     against oracle/shim (oracle/_ref, canonical -O2 -ffp-contract=off
                                     build), one array per model.  `gat` is the whole dataset as ONE batch
                                     (the reference's missing node offset active);
                                     `gat_per_graph` evaluates each graph as its own batch (offset bug
                                     cannot trigger).  `ginvn` runs on the virtual-node-augmented batch.

Usage:  make -C oracle ref && python tools/make_fixtures.py [dataset ...]     (default: all three)
"""
from __future__ import annotations

import os
import shutil
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from flowgnn_b200.dataset import load_dataset_zip  # noqa: E402
from flowgnn_b200.weights import load_weights  # noqa: E402
from oracle.refbind import run_reference  # noqa: E402

REF = os.environ.get("FLOWGNN_REFERENCE", "/root/reference")
GOLD = os.path.join(ROOT, "tests", "golden")

WEIGHT_FILES = {
    "GIN": ["gin_ep1_noBN_dim100.weights.all.bin"],
    "GCN": ["gcn_ep1_dim100.weights.all.bin"],
    "PNA": ["pna_ep1_noBN_dim80.weights.all.bin"],
    "DGN": ["dgn_ep1_noBN_dim100.weights.all.bin"],
    "GAT": [f"gat_ep1_{n}_layer5.bin" for n in (
        "pred_weights", "pred_bias", "scoring_fn_target", "scoring_fn_source",
        "linear_proj_weight_0", "linear_proj_weight_1", "skip_proj_weight_0", "skip_proj_weight_1")],
}
MODEL_DIR = {"gin": "GIN", "ginvn": "GIN", "gcn": "GCN", "gat": "GAT", "pna": "PNA", "dgn": "DGN"}
DATASETS = {"molhiv": 4113, "molpcba": 4113, "hep10k": 1000}


def main() -> None:
    os.makedirs(GOLD, exist_ok=True)
    for d, files in WEIGHT_FILES.items():
        os.makedirs(os.path.join(GOLD, "weights", d), exist_ok=True)
        for f in files:
            shutil.copyfile(os.path.join(REF, d, f), os.path.join(GOLD, "weights", d, f))
            os.chmod(os.path.join(GOLD, "weights", d, f), 0o644)

    # the split GIN files the reference host actually reads must equal the packed blob we ship
    from flowgnn_b200 import weights as W
    a = W._load_gin(os.path.join(GOLD, "weights", "GIN"))
    tmp = os.path.join(GOLD, "_tmp_gin_split")
    os.makedirs(tmp, exist_ok=True)
    for f in os.listdir(os.path.join(REF, "GIN")):
        if f.startswith("gin_ep1_") and f.endswith("_dim100.bin"):
            shutil.copyfile(os.path.join(REF, "GIN", f), os.path.join(tmp, f))
    b = W._load_gin(tmp)
    shutil.rmtree(tmp)
    for k in a:
        assert np.array_equal(a[k], b[k]), f"GIN split file vs .all.bin mismatch in {k}"
    print("GIN: .weights.all.bin == the nine split files")

    weights = {m: load_weights(m, os.path.join(GOLD, "weights", d)) for m, d in MODEL_DIR.items()}

    only = set(sys.argv[1:])
    for ds, count in DATASETS.items():
        if only and ds not in only:
            continue
        t = time.time()
        batch = load_dataset_zip(os.path.join(REF, f"{ds}.zip"), count, with_eigen=True)
        batch.save_npz(os.path.join(GOLD, f"{ds}.npz"))
        print(f"{ds}: {batch.num_graphs} graphs, {batch.total_nodes} nodes, {batch.total_edges} edges "
              f"({time.time() - t:.1f}s to load)")
        gold = {}
        for m in MODEL_DIR:
            t = time.time()
            bb = batch.with_virtual_node() if m == "ginvn" else batch
            gold[m] = run_reference(m, bb, weights[m])
            print(f"  {m:6s} {time.time() - t:6.1f}s  finite={np.isfinite(gold[m]).sum()}/{count} "
                  f"range=[{np.nanmin(gold[m]):.4f}, {np.nanmax(gold[m]):.4f}]")
        per_graph = np.zeros(count, dtype=np.float32)
        for g in range(count):
            per_graph[g] = run_reference("gat", batch.slice(g, g + 1), weights["gat"])[0]
        gold["gat_per_graph"] = per_graph
        np.savez(os.path.join(GOLD, f"golden_{ds}.npz"), **gold)


if __name__ == "__main__":
    main()
