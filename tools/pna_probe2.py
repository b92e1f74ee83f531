"""PNA layer probe (GPU box): per-layer device time of the fused kernel (pna_fused.cu) and of the aggregate -> GEMM path
(pna_tc.cu) on the molpcba-shaped workload, plus agreement of the two.
    python tools/pna_probe2.py [graphs=437929] [reps=5]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from flowgnn_b200.capi import Context  # noqa: E402
from flowgnn_b200.weights import load_weights  # noqa: E402

G = int(sys.argv[1]) if len(sys.argv) > 1 else 437929
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
w = load_weights("pna", os.path.join(ROOT, "tests", "golden", "weights", "PNA"))
b = bench.make_workload("pna", G, base_graphs=8192)
print(f"molpcba-shaped: {b.num_graphs} graphs, {b.total_nodes} nodes, {b.total_edges} edges; layer bytes {bench.layer_bytes('pna', b.total_nodes, b.total_edges) / 1e6:.1f} MB", flush=True)
out = {}
with Context(0) as ctx:
    ctx.load_weights("pna", w)
    ctx.upload(b)
    ctx.set_option("time_layers", 1)
    for label, fused in (("fused (pna_fused.cu)", 1), ("aggregate -> GEMM (pna_tc.cu)", 0)):
        ctx.set_option("pna_fused", fused)
        for _ in range(2):
            ctx.compute("pna")
        ms, lay = [], []
        for _ in range(reps):
            ms.append(ctx.compute("pna"))
            lay.append(ctx.last_layer_ms())
        out[fused] = ctx.download().copy()
        print(f"{label:34s} step {np.mean(ms):8.3f} ms  layers {np.round(np.mean(lay, 0), 3).tolist()}  launches {ctx.last_launch_count}", flush=True)
err = np.abs(out[1] - out[0]) / np.maximum(1, np.abs(out[0]))
print("fused vs aggregate->GEMM: max scaled difference %.3e, finite %d / %d" % (np.nanmax(err), np.isfinite(out[1]).sum(), len(err)))
