"""Roofline sweep of the GIN / GIN-VN layer over the batch size (BASELINE.json config 5: hep10k-shaped graphs).
For every batch size: device-timed layer time (all layer launches of a forward, one event pair), graphs/s of the whole
forward, achieved algorithmic GB/s of a layer and of the edge gather alone (mp_only), as fractions of the measured HBM peak.
usage: python tools/roofline_sweep.py [hep10k|molhiv] [out.json]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from flowgnn_b200.capi import Context
from flowgnn_b200.dataset import synthetic_hep, synthetic_molecules
from flowgnn_b200.weights import load_weights

shape = sys.argv[1] if len(sys.argv) > 1 else "hep10k"
out = sys.argv[2] if len(sys.argv) > 2 else None
model = "ginvn" if shape == "hep10k" else "gin"
w = load_weights(model, os.path.join(ROOT, "tests", "golden", "weights", "GIN"))
peak = 6650.0
pp = os.path.join(ROOT, "MEASURED_PEAKS.json")
if os.path.isfile(pp):
    peak = float(json.load(open(pp))["hbm_gbs"])
base = synthetic_hep(2048, seed=11) if shape == "hep10k" else synthetic_molecules(2048, "molhiv", seed=11)
rows = []
with Context(0) as c:
    c.load_weights(model, w)
    for G in (256, 1024, 4096, 16384, 40000 if shape == "hep10k" else 41127):
        b = base.tile(G)
        if model == "ginvn":
            b = b.with_virtual_node()
        N, E = b.total_nodes, b.total_edges
        lb = 8 * 100 * N + 20 * E
        c.upload(b)
        res = {"graphs": G, "nodes": int(N), "edges": int(E), "algorithmic_MB_per_layer": lb / 1e6}
        for key, mp in (("layer", 0), ("gather", 1)):
            c.set_option("mp_only", mp)
            c.set_option("time_layers", 2)
            for _ in range(3):
                c.compute(model, timed=True)
            ms, tot = [], []
            for _ in range(10):
                tot.append(c.compute(model, timed=True))
                ms.append(c.last_layer_ms()[0] / 5)
            m = float(np.mean(ms))
            res[key + "_ms"] = m
            res[key + "_GBps"] = lb / m / 1e6
            res[key + "_frac"] = lb / m / 1e6 / peak
            if not mp:
                res["forward_ms"] = float(np.mean(tot))
                res["graphs_per_s"] = G / (float(np.mean(tot)) * 1e-3)
        c.set_option("mp_only", 0)
        rows.append(res)
        print(json.dumps(res), flush=True)
if out:
    json.dump({"shape": shape, "model": model, "hbm_peak_GBps": peak, "rows": rows}, open(out, "w"), indent=1)
