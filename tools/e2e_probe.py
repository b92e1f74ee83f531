"""End-to-end time of the host-pointer entry point for chunk counts (FLOWGNN_B200_CHUNKS), the narrowed upload on / off
(FLOWGNN_B200_HOST_STAGE), host thread counts (FLOWGNN_B200_HOST_THREADS) and pageable / page-locked caller arrays.
    python tools/e2e_probe.py [model=gin] [reps=20] [trace ['{"HOST_STAGE": 0}' ...]]   (trace: pinned arrays, timeline of one call per setting)"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import bench
from flowgnn_b200.capi import ReferenceCall, pin_host
from flowgnn_b200.weights import load_weights

model = sys.argv[1] if len(sys.argv) > 1 else "gin"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
d = {"gin": "GIN", "ginvn": "GIN", "gcn": "GCN", "gat": "GAT", "pna": "PNA", "dgn": "DGN"}[model]
w = load_weights(model, os.path.join(ROOT, "tests", "golden", "weights", d))
b = bench.make_workload(model, bench.WORKLOADS[model][1], base_graphs=4096)
call = ReferenceCall(model, b, w)


def measure(tag, **env):
    for k in ("FLOWGNN_B200_CHUNKS", "FLOWGNN_B200_HOST_STAGE", "FLOWGNN_B200_HOST_THREADS", "FLOWGNN_B200_GRADE"):
        os.environ.pop(k, None)
    for k, v in env.items():
        os.environ["FLOWGNN_B200_" + k] = str(v)
    for _ in range(3):
        call.run()
    ts = []
    for _ in range(reps):
        t = time.perf_counter(); call.run(); ts.append(time.perf_counter() - t)
    ms = float(np.median(ts)) * 1e3
    print(f"{tag:10s} {str(env):60s} {ms:7.3f} ms = {b.num_graphs / ms / 1e3:6.2f} M graphs/s (min {min(ts) * 1e3:.3f} ms)", flush=True)


if len(sys.argv) > 3 and sys.argv[3] == "trace":
    for a in (b.node_feature, b.edge_list, b.edge_attr, b.node_eigen):
        if a is not None:
            pin_host(a)
    import json
    envs = [json.loads(a) for a in sys.argv[4:]] or [{}, {"HOST_STAGE": 7}, {"GRADE": "1,2,4"}, {"GRADE": "1,3,6"}, {"GRADE": "1,2,3,4"}, {"HOST_STAGE": 0},
                                                      {"HOST_STAGE": 5}, {"HOST_THREADS": 8}, {"HOST_THREADS": 16}]
    for env in envs:
        measure("pinned", **env)
        if len(sys.argv) > 4 or len(env) == 0 or env.get("GRADE") == "1,2,4":
            os.environ["FLOWGNN_B200_E2E_TRACE"] = "1"
            call.run()
            del os.environ["FLOWGNN_B200_E2E_TRACE"]
    sys.exit(0)
if len(sys.argv) > 3 and sys.argv[3] == "pageable-trace":
    measure("pageable")
    os.environ["FLOWGNN_B200_E2E_TRACE"] = "1"
    call.run()
    sys.exit(0)
measure("pageable")
measure("pageable", HOST_STAGE=0)
for t in (6, 8, 12):
    measure("pageable", HOST_STAGE=7, HOST_THREADS=t)
for c in (2, 3, 4, 5, 6, 8):
    measure("pageable", HOST_STAGE=7, CHUNKS=c)
for a in (b.node_feature, b.edge_list, b.edge_attr, b.node_eigen):
    if a is not None:
        pin_host(a)
measure("pinned")
measure("pinned", HOST_STAGE=0)
for m in (7, 5, 4, 1, 6, 3):
    measure("pinned", HOST_STAGE=m)
for m in (5, 4, 1):
    for c in (2, 4, 5, 6):
        measure("pinned", HOST_STAGE=m, CHUNKS=c)
for m in (7, 5):
    for t in (4, 6, 12):
        measure("pinned", HOST_STAGE=m, HOST_THREADS=t)
