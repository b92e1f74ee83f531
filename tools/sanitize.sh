#!/bin/bash
# usage (GPU box): bash tools/sanitize.sh <tag>   -- memcheck, racecheck and synccheck over tools/sanitize_all.py;
# full logs to gpurun_out/<tag>_sanitizer_<tool>.log, the ERROR SUMMARY lines are what profiles/ keeps
tag=${1:-r2}
for tool in memcheck racecheck synccheck; do
  lim=600; [ $tool = racecheck ] && lim=900
  SAN_MOL=${SAN_MOL:-150} SAN_HEP=${SAN_HEP:-12} timeout $lim compute-sanitizer --tool $tool --print-limit 30 \
      python tools/sanitize_all.py > gpurun_out/${tag}_sanitizer_${tool}.log 2>&1
  echo "$tool rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|sanitize workload done' gpurun_out/${tag}_sanitizer_${tool}.log | tr '\n' ' ')"
done
