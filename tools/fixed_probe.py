#!/usr/bin/env python
"""Time the ap_fixed flavours (option fixed_point) of GIN and DGN on the BASELINE batch (41,127 molhiv-shaped graphs), resident in HBM."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from flowgnn_b200.capi import Context  # noqa: E402
from flowgnn_b200.dataset import load_npz  # noqa: E402
from flowgnn_b200.weights import load_weights  # noqa: E402

G = 41127
b = load_npz(os.path.join(ROOT, "tests", "golden", "molhiv.npz")).tile(G)
res = {}
with Context(0) as c:
    for model, d in (("gin", "GIN"), ("dgn", "DGN")):
        w = load_weights(model, os.path.join(ROOT, "tests", "golden", "weights", d))
        c.load_weights(model, w)
        c.upload(b)
        for fixed in (0, 1):
            c.set_option("fixed_point", fixed)
            c.set_option("time_layers", 1)
            for _ in range(3):
                c.compute(model)
            ms = [c.compute(model) for _ in range(5)]
            res[f"{model}_fixed{fixed}"] = {"ms": float(np.median(ms)), "graphs_per_s": G / (float(np.median(ms)) * 1e-3),
                                           "layer_ms": [round(x, 4) for x in c.last_layer_ms()]}
        c.set_option("fixed_point", 0)
        c.set_option("time_layers", 0)
print(json.dumps(res))
