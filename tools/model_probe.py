"""Per-layer device time of one model on its bench workload, resident in HBM (GPU box).
    python tools/model_probe.py <model> [graphs=41127] [reps=10] [option=value ...]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from flowgnn_b200.capi import Context  # noqa: E402
from flowgnn_b200.weights import load_weights  # noqa: E402

model = sys.argv[1]
G = int(sys.argv[2]) if len(sys.argv) > 2 else 41127
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
opts = dict(a.split("=") for a in sys.argv[4:])
wdir = {"gin": "GIN", "ginvn": "GIN", "gcn": "GCN", "gat": "GAT", "pna": "PNA", "dgn": "DGN"}[model]
w = load_weights(model, os.path.join(ROOT, "tests", "golden", "weights", wdir))
with Context(0) as ctx:
    ctx.load_weights(model, w)
    b = bench.make_workload(model, G, base_graphs=4096)
    ctx.upload(b)
    for k, v in opts.items():
        ctx.set_option(k, int(v))
    ctx.set_option("time_layers", 1)
    for _ in range(3):
        ctx.compute(model)
    ms, lay = [], []
    for _ in range(reps):
        ms.append(ctx.compute(model))
        lay.append(ctx.last_layer_ms())
    lay = np.array(lay).mean(0)
    lb = bench.layer_bytes(model, b.total_nodes, b.total_edges)
    print(f"{model} {opts} {b.num_graphs} graphs {b.total_nodes} nodes {b.total_edges} edges: step {np.mean(ms):.3f} ms (min {np.min(ms):.3f}) = "
          f"{b.num_graphs / np.mean(ms) / 1e3:.2f} M graphs/s; layer intervals {np.round(lay, 4).tolist()} ms; layer bytes {lb / 1e6:.1f} MB -> "
          f"{lb / (np.mean(lay) * 1e-3) / 1e9:.0f} GB/s", flush=True)
