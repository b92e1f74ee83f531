"""GCN / DGN layer time, FFMA kernel vs tensor-core path, molhiv-shaped batch.  usage: python tools/tc_probe_models.py [graphs]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from flowgnn_b200.capi import Context
from flowgnn_b200.dataset import synthetic_molecules
from flowgnn_b200.weights import load_weights
G = int(sys.argv[1]) if len(sys.argv) > 1 else 41127
for model, d in (("gcn", "GCN"), ("dgn", "DGN")):
    w = load_weights(model, os.path.join(ROOT, "tests", "golden", "weights", d))
    big = synthetic_molecules(4096, "molhiv", seed=11, with_eigen=(model == "dgn")).tile(G)
    ys = {}
    for name, v in (("ffma", 0), ("tc", 1)):
        with Context(0) as c:
            c.set_option("time_layers", 1); c.set_option(model + "_tc", v)
            c.load_weights(model, w); c.upload(big)
            for _ in range(2): c.compute(model)
            ms, tot = [], []
            for _ in range(5):
                tot.append(c.compute(model)); ms += c.last_layer_ms()[:4]
            ys[name] = c.download()
            print(f"{model} {name:5s} layer {float(np.mean(ms)):8.3f} ms  forward {float(np.mean(tot)):8.3f} ms", flush=True)
    d = np.abs(ys["tc"] - ys["ffma"]) / np.maximum(1, np.abs(ys["ffma"]))
    print(model, "max scaled diff tc vs ffma", float(np.nanmax(d)), "nonfinite", int((~np.isfinite(ys["tc"])).sum()), int((~np.isfinite(ys["ffma"])).sum()), flush=True)
