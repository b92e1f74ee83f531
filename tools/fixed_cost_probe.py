"""Fixed cost of one forward: device time of GIN on tiny batches, per kernel (per-layer events).  usage: python tools/fixed_cost_probe.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from flowgnn_b200.capi import Context
from flowgnn_b200.dataset import synthetic_molecules
from flowgnn_b200.weights import load_weights
w = load_weights("gin", os.path.join(ROOT, "tests", "golden", "weights", "GIN"))
base = synthetic_molecules(2048, "molhiv", seed=11)
with Context(0) as c:
    c.set_option("time_layers", 1)
    c.load_weights("gin", w)
    for n in (16, 512, 2048, 6855, 13709, 20564, 41127):
        b = base.tile(n)
        c.upload(b)
        for _ in range(3): c.compute("gin")
        ms = [c.compute("gin") for _ in range(20)]
        print(f"graphs={n:6d} nodes={b.total_nodes:8d} step {np.median(ms)*1e3:8.1f} us  layers {np.round(np.array(c.last_layer_ms())*1e3,1)} us", flush=True)
