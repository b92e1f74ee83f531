"""End-to-end time of the host-pointer entry point (pinned caller buffers) for a few chunk counts (env FLOWGNN_B200_CHUNKS).
    python tools/e2e_probe.py [model=gin] [reps=30]"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

if len(sys.argv) > 3 and sys.argv[3] == "child":
    import numpy as np
    import bench
    from flowgnn_b200.capi import ReferenceCall, pin_host
    from flowgnn_b200.weights import load_weights
    model, reps = sys.argv[1], int(sys.argv[2])
    d = {"gin": "GIN", "ginvn": "GIN", "gcn": "GCN", "gat": "GAT", "pna": "PNA", "dgn": "DGN"}[model]
    w = load_weights(model, os.path.join(ROOT, "tests", "golden", "weights", d))
    b = bench.make_workload(model, bench.WORKLOADS[model][1], base_graphs=4096)
    for a in (b.node_feature, b.edge_list, b.edge_attr, b.node_eigen):
        if a is not None:
            pin_host(a)
    call = ReferenceCall(model, b, w)
    for _ in range(5):
        call.run()
    ts = []
    for _ in range(reps):
        t = time.perf_counter(); call.run(); ts.append(time.perf_counter() - t)
    ms = float(np.median(ts)) * 1e3
    print(f"chunks={os.environ.get('FLOWGNN_B200_CHUNKS', 'default')}: {ms:.3f} ms per call = {b.num_graphs / ms / 1e3:.2f} M graphs/s (min {min(ts) * 1e3:.3f} ms)", flush=True)
else:
    model = sys.argv[1] if len(sys.argv) > 1 else "gin"
    reps = sys.argv[2] if len(sys.argv) > 2 else "30"
    for c in ("", "1", "2", "3", "4", "5", "6", "8"):
        env = dict(os.environ)
        if c:
            env["FLOWGNN_B200_CHUNKS"] = c
        subprocess.run([sys.executable, __file__, model, reps, "child"], env=env)
