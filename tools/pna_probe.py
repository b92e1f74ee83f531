"""PNA layer time, FFMA kernel vs tensor-core path, on a molpcba-shaped batch.  usage: python tools/pna_probe.py [graphs]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from flowgnn_b200.capi import Context
from flowgnn_b200.dataset import synthetic_molecules
from flowgnn_b200.weights import load_weights
G = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
w = load_weights("pna", os.path.join(ROOT, "tests", "golden", "weights", "PNA"))
big = synthetic_molecules(4096, "molpcba", seed=11).tile(G)
print(f"N {big.total_nodes} E {big.total_edges}", flush=True)
ys = {}
only = sys.argv[2] if len(sys.argv) > 2 else ""
for name, v in (("ffma", 0), ("tc", 1)):
    if only and name != only:
        continue
    with Context(0) as c:
        c.set_option("time_layers", 1); c.set_option("pna_tc", v)
        c.load_weights("pna", w); c.upload(big)
        for _ in range(2): c.compute("pna")
        ms = []
        for _ in range(5):
            c.compute("pna"); ms += c.last_layer_ms()[:4]
        ys[name] = c.download()
        print(f"{name:5s} layer {float(np.mean(ms)):8.3f} ms  y[0:3] {ys[name][:3]}", flush=True)
if only:
    sys.exit(0)
d = np.abs(ys["tc"] - ys["ffma"]) / np.maximum(1, np.abs(ys["ffma"]))
print("max scaled diff tc vs ffma", float(np.nanmax(d)), "nonfinite", int((~np.isfinite(ys["tc"])).sum()), int((~np.isfinite(ys["ffma"])).sum()))
