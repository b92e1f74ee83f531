#!/bin/bash
# usage: bash tools/r2_probe.sh <tag> [ncu]
tag=${1:-p}
timeout 300 python tools/gin_probe.py 41127 10 2>&1 | tee gpurun_out/${tag}_probe.txt
if [ "$2" = ncu ]; then
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:gin_layer_fused -s 8 -c 2 -f -o gpurun_out/${tag}_fused python tools/gin_probe.py 41127 1 fused > gpurun_out/${tag}_ncu1.log 2>&1; tail -1 gpurun_out/${tag}_ncu1.log
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:gin_layer_fused -s 8 -c 2 -f -o gpurun_out/${tag}_mp python tools/gin_probe.py 41127 1 mp > gpurun_out/${tag}_ncu2.log 2>&1; tail -1 gpurun_out/${tag}_ncu2.log
fi
