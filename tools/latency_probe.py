"""BASELINE config 1: ONE molhiv graph through the host-pointer entry point -- the latency of a single-graph call (H2D, every launch of the
forward, D2H, synchronisation), for all six models.   python tools/latency_probe.py [reps=200]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from flowgnn_b200.capi import ReferenceCall  # noqa: E402
from flowgnn_b200.dataset import load_npz  # noqa: E402
from flowgnn_b200.weights import load_weights  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
gold = os.path.join(ROOT, "tests", "golden")
g1 = load_npz(os.path.join(gold, "molhiv.npz")).slice(0, 1)
ref = dict(np.load(os.path.join(gold, "golden_molhiv.npz")))
for model, d in (("gin", "GIN"), ("ginvn", "GIN"), ("gcn", "GCN"), ("gat", "GAT"), ("pna", "PNA"), ("dgn", "DGN")):
    b = g1.with_virtual_node() if model == "ginvn" else g1
    call = ReferenceCall(model, b, load_weights(model, os.path.join(gold, "weights", d)))
    for _ in range(20):
        y = call.run()
    ts = []
    for _ in range(reps):
        t = time.perf_counter(); y = call.run(); ts.append(time.perf_counter() - t)
    print(f"{model:6s} g1 ({b.total_nodes} nodes, {b.total_edges} edges): median {np.median(ts) * 1e6:7.1f} us, min {min(ts) * 1e6:7.1f} us per call; "
          f"y = {y[0]:.6f} (reference {ref[model][0]:.6f})", flush=True)
