"""Stand-alone edge gather (mp_only) timing on the bench workload.  usage: python tools/gather_probe.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from flowgnn_b200.capi import Context
from flowgnn_b200.dataset import synthetic_molecules
from flowgnn_b200.weights import load_weights
w = load_weights("gin", os.path.join(ROOT, "tests", "golden", "weights", "GIN"))
big = synthetic_molecules(2048, "molhiv", seed=11).tile(41127)
N, E = big.total_nodes, big.total_edges
bytes_per_launch = 8 * 100 * N + E * 20
with Context(0) as c:
    c.set_option("time_layers", 1); c.set_option("mp_only", 1)
    c.load_weights("gin", w); c.upload(big)
    for _ in range(5): c.compute("gin")
    ms = []
    for _ in range(30):
        c.compute("gin"); ms += c.last_layer_ms()[:5]
    m = float(np.mean(ms))
    print(f"gather launch {m*1e3:.1f} us -> {bytes_per_launch / m / 1e6:.0f} GB/s algorithmic", flush=True)
