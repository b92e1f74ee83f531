#!/bin/bash
# usage: tools/quick_bench.sh <tag>   (run on the GPU box: parity tests for GIN + bench line summary)
tag=${1:-x}
timeout 600 python -m pytest tests -m gpu -x -q -k "gin or mp_only or synthetic" 2>&1 | tail -5
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; tail -3 gpurun_out/bench_$tag.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_$tag.json"))
print("value %.0f graphs/s  ms/step %.3f  e2e %.0f  layer_ms %.4f  frac %.3f  mp_only_ms %.4f frac %.3f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["mean_launch_ms"], d["roofline"]["frac"], d["edge_gather"]["mean_launch_ms"], d["edge_gather"]["frac"]))
PY
