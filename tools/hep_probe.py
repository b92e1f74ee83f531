"""GIN-VN on hep10k-shaped graphs: layer time of every GIN kernel variant.  usage: python tools/hep_probe.py [graphs]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from flowgnn_b200.capi import Context
from flowgnn_b200.dataset import synthetic_hep
from flowgnn_b200.weights import load_weights
G = int(sys.argv[1]) if len(sys.argv) > 1 else 40000
w = load_weights("ginvn", os.path.join(ROOT, "tests", "golden", "weights", "GIN"))
big = synthetic_hep(4096, seed=11).tile(G).with_virtual_node()
N, E = big.total_nodes, big.total_edges
bytes_per_launch = 8 * 100 * N + E * 20
print(f"N {N} E {E} algorithmic MB {bytes_per_launch/1e6:.0f}", flush=True)
VARIANTS = (("staged+mlp", {}), ("tc2 fused", {"gin_staged": 0}), ("mp_only staged", {"mp_only": 1}), ("mp_only rows", {"mp_only": 1, "gin_staged": 0}))
if len(sys.argv) > 2:
    VARIANTS += (("tc1", {"gin_tc1": 1}), ("tc3", {"gin_tc3": 1}), ("ffma", {"gin_ffma": 1}))
for name, opts in VARIANTS:
    with Context(0) as c:
        c.set_option("time_layers", 1)
        for k, v in opts.items(): c.set_option(k, v)
        c.load_weights("ginvn", w); c.upload(big)
        for _ in range(3): c.compute("ginvn")
        ms = []
        for _ in range(10):
            c.compute("ginvn"); ms += c.last_layer_ms()[:4]
        m = float(np.mean(ms))
        y = c.download()
        print(f"{name:15s} layer {m*1e3:8.1f} us -> {bytes_per_launch / m / 1e6:6.0f} GB/s algorithmic; y[0:3] {y[:3]}", flush=True)
