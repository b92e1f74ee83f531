#!/bin/bash
# usage (GPU box): bash tools/r2_gin.sh <tag> -- GIN tests + bench summary (+ optional env FUSED_NCU=1: ncu capture of the layer kernel)
tag=${1:-r2b}
timeout 900 python -m pytest tests -m gpu -x -q -k "gin or mp_only or synthetic or tile" 2>&1 | tail -8
timeout 300 python bench.py --no-cpu-baseline --no-extras --no-pageable --steps 20 --warmup 5 > gpurun_out/${tag}_bench_gin.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/${tag}_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/${tag}_bench_gin.json"))
print("value %.0f graphs/s  ms/step %.3f  e2e %.0f  layer_ms %.4f  frac %.3f  mp_only_ms %.4f frac %.3f  side %.4f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["mean_launch_ms"], d["roofline"]["frac"], d["edge_gather"]["mean_launch_ms"], d["edge_gather"]["frac"], d["edge_gather_side_kernel"]["mean_launch_ms"]))
PY
for st in 0 -1; do
FLOWGNN_STAGED=$st timeout 200 python - <<PY
import os, sys, numpy as np
sys.path.insert(0, ".")
import bench
from flowgnn_b200.capi import Context
from flowgnn_b200.weights import load_weights
st = int(os.environ["FLOWGNN_STAGED"])
b = bench.make_workload("ginvn", 40000, base_graphs=2048)
w = load_weights("gin", "tests/golden/weights/GIN")
with Context(0) as c:
    c.load_weights("gin", w); c.upload(b); c.set_option("gin_staged", st); c.set_option("time_layers", 1)
    for _ in range(3): c.compute("gin")
    ms = [c.compute("gin") for _ in range(8)]
    print("ginvn hep10k 40000 graphs, gin_staged=%d: %.3f ms/step, layer launches %s" % (st, np.mean(ms), ["%.3f" % x for x in c.last_layer_ms()]))
PY
done
if [ -n "$FUSED_NCU" ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gin_layer_fused -s 6 -c 2 -f -o gpurun_out/${tag}_fused python bench.py --no-cpu-baseline --no-extras --no-pageable --steps 3 --warmup 3 > gpurun_out/${tag}_ncu.log 2>&1; tail -2 gpurun_out/${tag}_ncu.log
fi
