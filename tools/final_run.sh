#!/bin/bash
# usage (on the GPU box): bash tools/final_run.sh <tag>  -- end-of-round evidence: GPU tests, smoke, bench lines, sweep, ncu
tag=${1:-r1z}
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -3
timeout 300 python bench.py > gpurun_out/${tag}_bench_gin.json 2> gpurun_out/${tag}_bench.err; tail -2 gpurun_out/${tag}_bench.err
timeout 200 python bench.py --model ginvn --steps 20 --warmup 3 --base-graphs 8192 > gpurun_out/${tag}_bench_ginvn.json 2> gpurun_out/${tag}_bench_ginvn.err; tail -2 gpurun_out/${tag}_bench_ginvn.err
timeout 200 python tools/roofline_sweep.py hep10k gpurun_out/${tag}_roofline_sweep_ginvn_hep10k.json 2>&1 | tail -6
timeout 200 ncu --set full --clock-control none --import-source on -k regex:gin_gather_staged -s 3 -c 1 -f -o gpurun_out/${tag}_sg python tools/prof_hep.py 1 1 > gpurun_out/${tag}_sg.log 2>&1; tail -1 gpurun_out/${tag}_sg.log
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/${tag}_launches_ginvn.csv python tools/prof_hep.py 0 2 > /dev/null 2>&1
python - <<PY
import json
for m in ("gin", "ginvn"):
    d = json.load(open("gpurun_out/${tag}_bench_%s.json" % m))
    print("%s value %.0f ms/step %.3f e2e %.0f layer_ms %.4f frac %.3f gather %.4f ms frac %.3f" % (m, d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["mean_launch_ms"], d["roofline"]["frac"], d["edge_gather"]["mean_launch_ms"], d["edge_gather"]["frac"]))
PY
