#!/bin/bash
# usage (GPU box): bash tools/r2_final.sh <tag> -- the round's evidence run: GPU suite, smoke, bench (both arms), launch list,
# ncu full captures of the GIN and PNA layer kernels, per-model probes, sanitizer
tag=${1:-r2z}
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/${tag}_gputests.txt
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -3
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench_gin.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/${tag}_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_reference_gin.json 2>> gpurun_out/${tag}_bench.err; echo "ref rc=$?"
for m in gcn gat dgn pna ginvn; do timeout 200 python tools/model_probe.py $m $( [ $m = pna ] && echo 437929 || ( [ $m = ginvn ] && echo 40000 || echo 41127 ) ) 5 2>&1 | tail -1 | tee -a gpurun_out/${tag}_model_probe.txt; done
timeout 120 python tools/fixed_probe.py 2>&1 | tail -1 > gpurun_out/${tag}_fixed_probe.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 200 --csv --log-file gpurun_out/${tag}_launches_gin_bench.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras --no-pageable > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gin_layer_fused -s 8 -c 1 -f -o gpurun_out/${tag}_ncu_gin_layer_fused python tools/gin_probe.py 41127 1 fused > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pna_layer_fused -s 1 -c 1 -f -o gpurun_out/${tag}_ncu_pna_layer_fused python tools/model_probe.py pna 100000 1 > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:fused_kernel -s 2 -c 1 -f -o gpurun_out/${tag}_ncu_gat python tools/model_probe.py gat 41127 1 > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:fused_kernel -s 6 -c 1 -f -o gpurun_out/${tag}_ncu_gcn python tools/model_probe.py gcn 41127 1 > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:fused_kernel -s 2 -c 1 -f -o gpurun_out/${tag}_ncu_dgn python tools/model_probe.py dgn 41127 1 > /dev/null 2>&1
for m in gat gcn dgn pna; do timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${tag}_launches_${m}.csv python tools/model_probe.py $m $( [ $m = pna ] && echo 100000 || echo 41127 ) 1 > /dev/null 2>&1; done
timeout 200 python tools/e2e_probe.py gin 20 > gpurun_out/${tag}_e2e_probe.txt 2>&1
timeout 100 python tools/e2e_probe.py gin 10 trace > gpurun_out/${tag}_e2e_trace.txt 2>&1
for m in dgn pna ginvn; do timeout 100 python tools/e2e_probe.py $m 5 2>&1 | head -3 >> gpurun_out/${tag}_e2e_probe.txt; done
bash tools/hs_probe/run.sh > gpurun_out/${tag}_host_probe.txt 2>&1
bash tools/sanitize.sh ${tag}
ls -la gpurun_out/ | grep ${tag}
