"""GIN layer-kernel probe (GPU box): per-layer device time of the fused kernel, its mp_only mode and the round-1 pair kernel on
the molhiv-shaped bench workload, and of the dense (hep10k-shaped, virtual node) workload fused vs staged.
    python tools/gin_probe.py [graphs=41127] [reps=10] [modes=fused,mp,tc2,hep]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from flowgnn_b200.capi import Context  # noqa: E402
from flowgnn_b200.weights import load_weights  # noqa: E402

G = int(sys.argv[1]) if len(sys.argv) > 1 else 41127
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
modes = (sys.argv[3] if len(sys.argv) > 3 else "fused,mp,tc2,hep").split(",")
w = load_weights("gin", os.path.join(ROOT, "tests", "golden", "weights", "GIN"))


def run(ctx, label, opts, grouped):
    for k, v in opts.items():
        ctx.set_option(k, v)
    ctx.set_option("time_layers", 2 if grouped else 1)
    for _ in range(3):
        ctx.compute("gin")
    ms, lay = [], []
    for _ in range(reps):
        ms.append(ctx.compute("gin"))
        lay.append(ctx.last_layer_ms())
    lay = np.array(lay)
    per = lay.mean(0) / (5 if grouped else 1)
    print(f"{label:42s} step {np.mean(ms):7.3f} ms (min {np.min(ms):.3f})  layer launches {np.round(per, 4).tolist()}", flush=True)
    for k in opts:
        ctx.set_option(k, -1 if k == "gin_staged" else 0)


with Context(0) as ctx:
    ctx.load_weights("gin", w)
    if any(m in modes for m in ("fused", "mp", "tc2")):
        b = bench.make_workload("gin", G, base_graphs=4096)
        ctx.upload(b)
        print(f"molhiv-shaped: {b.num_graphs} graphs, {b.total_nodes} nodes, {b.total_edges} edges; layer bytes {bench.layer_bytes('gin', b.total_nodes, b.total_edges) / 1e6:.1f} MB")
        if "fused" in modes:
            run(ctx, "fused (grouped, PDL)", {}, True)
            run(ctx, "fused (per-layer events)", {}, False)
        if "mp" in modes:
            run(ctx, "fused mp_only", {"mp_only": 1}, False)
            run(ctx, "stand-alone gather kernel", {"mp_only": 2}, False)
        if "tc2" in modes:
            run(ctx, "round-1 pair kernel (grouped, PDL)", {"gin_tc2": 1}, True)
    if "hep" in modes:
        b = bench.make_workload("ginvn", min(G, 40000), base_graphs=2048)
        ctx.upload(b)
        print(f"hep10k-shaped + VN: {b.num_graphs} graphs, {b.total_nodes} nodes, {b.total_edges} edges; layer bytes {bench.layer_bytes('ginvn', b.total_nodes, b.total_edges) / 1e6:.1f} MB")
        run(ctx, "dense: fused single launch", {"gin_staged": 0}, False)
        run(ctx, "dense: staged gather + MLP launch", {"gin_staged": 1}, False)
        run(ctx, "dense: fused mp_only", {"gin_staged": 0, "mp_only": 1}, False)
