import os, sys
sys.path.insert(0, "/root/repo")
import numpy as np
from flowgnn_b200.capi import Context
from flowgnn_b200.dataset import load_npz
from flowgnn_b200.weights import load_weights
gold = "/root/repo/tests/golden"
b = load_npz(os.path.join(gold, "molhiv.npz")).slice(0, 300)
h = load_npz(os.path.join(gold, "hep10k.npz")).slice(0, 40)
with Context(0) as c:
    for model, d in (("gin", "GIN"), ("pna", "PNA"), ("gcn", "GCN")):
        w = load_weights(model, os.path.join(gold, "weights", d))
        for v in ("tc2", "tc1", "tc3"):
            c.set_option("gin_tc1", int(v == "tc1")); c.set_option("gin_tc3", int(v == "tc3"))
            y = c.run(model, b, w); y2 = c.run(model, h, w)
            print(model, v, float(np.abs(y).max()), float(np.abs(y2).max()), flush=True)
            if model != "gin": break
