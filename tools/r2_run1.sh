#!/bin/bash
# first GPU call of round 2: topology probe, GPU tests, smoke, bench (both arms), sanitizer
mkdir -p gpurun_out
{ nvidia-smi topo -m; echo; nproc; cat /sys/fs/cgroup/cpuset.cpus.effective 2>/dev/null; cat /sys/fs/cgroup/cpuset.mems.effective 2>/dev/null;
  lscpu | grep -iE 'numa|socket|model name|^cpu\(s\)'; for d in /sys/bus/pci/devices/*; do if [ "$(cat $d/vendor)" = 0x10de ] && [ -e $d/numa_node ]; then echo "$d $(cat $d/class) numa $(cat $d/numa_node) cpus $(cat $d/local_cpulist)"; fi; done;
  python -c "import os;print('affinity',sorted(os.sched_getaffinity(0)))"; } > gpurun_out/r2a_topology.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/r2a_gputests.txt
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -3
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_bench_gin.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2a_bench.err
timeout 200 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/r2a_bench_reference_gin.json 2>/dev/null; echo "ref rc=$?"
python - <<PY
import json
d = json.load(open("gpurun_out/r2a_bench_gin.json"))
print("value %.0f ms/step %.3f e2e %.0f layer_ms %.4f frac %.3f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["mean_launch_ms"], d["roofline"]["frac"]))
print("pageable", d.get("e2e_pageable")); print("affinity", d.get("affinity"))
for k, v in (d.get("extras") or {}).items():
    print(k, "%.0f graphs/s %.3f ms/step frac %.3f" % (v["value"], v["ms_per_step"], v["roofline"]["frac"]))
PY
bash tools/sanitize.sh r2a
