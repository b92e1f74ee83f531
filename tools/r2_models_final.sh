#!/bin/bash
# usage (GPU box): bash tools/r2_models_final.sh <tag> -- GPU suite + per-layer times and one ncu --set full capture of the GAT / GCN / DGN fused kernels
tag=${1:-r2y}
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/${tag}_gputests.txt
for m in gat gcn dgn; do timeout 120 python tools/model_probe.py $m 41127 10 2>&1 | tail -1 | tee -a gpurun_out/${tag}_model_probe.txt; done
timeout 400 ncu --set full --clock-control none --import-source on -k regex:fused_kernel -s 2 -c 1 -f -o gpurun_out/${tag}_ncu_gat python tools/model_probe.py gat 41127 1 > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:fused_kernel -s 6 -c 1 -f -o gpurun_out/${tag}_ncu_gcn python tools/model_probe.py gcn 41127 1 > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:fused_kernel -s 2 -c 1 -f -o gpurun_out/${tag}_ncu_dgn python tools/model_probe.py dgn 41127 1 > /dev/null 2>&1
for m in gat gcn dgn; do timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${tag}_launches_${m}.csv python tools/model_probe.py $m 41127 1 > /dev/null 2>&1; done
ls -la gpurun_out | grep ${tag}
