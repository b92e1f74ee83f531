#!/bin/bash
# usage (GPU box): bash tools/r2_checkpoint.sh <tag>  -- full GPU suite, smoke, bench (both arms), ncu launch list + full capture of the GIN layer kernel
tag=${1:-r2p}
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/${tag}_gputests.txt
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -3
timeout 500 python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench_gin.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/${tag}_bench.err
python - <<PY
import json
d = json.load(open("gpurun_out/${tag}_bench_gin.json"))
print("value %.0f ms/step %.3f e2e %.0f layer_ms %.4f frac %.3f mp_only %.4f ms frac %.3f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["mean_launch_ms"], d["roofline"]["frac"], d["edge_gather"]["mean_launch_ms"], d["edge_gather"]["frac"]))
print("pageable", d.get("e2e_pageable"))
for k, v in (d.get("extras") or {}).items():
    print(k, "%.0f graphs/s %.3f ms/step frac %.3f" % (v["value"], v["ms_per_step"], v["roofline"]["frac"]))
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 200 --csv --log-file gpurun_out/${tag}_launches_gin_bench.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras --no-pageable > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gin_layer_fused -s 8 -c 1 -f -o gpurun_out/${tag}_ncu_gin_layer_fused python tools/gin_probe.py 41127 1 fused > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gin_layer_fused -s 8 -c 1 -f -o gpurun_out/${tag}_ncu_gin_layer_fused_mp_only python tools/gin_probe.py 41127 1 mp > /dev/null 2>&1
ls -la gpurun_out/ | grep ${tag}
