import os, sys, numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from conftest import assert_parity, GOLDEN, MODEL_WEIGHT_DIR
from flowgnn_b200.capi import Context
from flowgnn_b200.dataset import load_npz
from flowgnn_b200.weights import load_weights
with Context(0) as c:
    for ds in ("molhiv", "molpcba", "hep10k"):
        b = load_npz(os.path.join(GOLDEN, f"{ds}.npz")); g = dict(np.load(os.path.join(GOLDEN, f"golden_{ds}.npz")))
        for m in ("gin", "ginvn", "gcn", "gat", "pna", "dgn"):
            if m == "gat" and ds == "hep10k": continue
            w = load_weights(m, os.path.join(GOLDEN, "weights", MODEL_WEIGHT_DIR[m]))
            bb = b.with_virtual_node() if m == "ginvn" else b
            y = c.run("gin" if m == "ginvn" else m, bb, w)
            print(ds, m, "%.2e" % assert_parity(y, g[m]), flush=True)
