timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 60 python __graft_entry__.py smoke 2>&1 | tail -3
timeout 200 python bench.py --model pna --steps 10 --warmup 3 --no-cpu-baseline --base-graphs 8192 > gpurun_out/r1z2_bench_pna.json 2> gpurun_out/r1z2_bench_pna.err; tail -2 gpurun_out/r1z2_bench_pna.err
python - <<PY
import json
d=json.load(open("gpurun_out/r1z2_bench_pna.json"))
print("pna value %.0f ms/step %.3f e2e %.0f layer_ms %.4f frac %.3f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["mean_launch_ms"], d["roofline"]["frac"]))
PY
timeout 100 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r1z2_launches_pna.csv python tools/pna_probe.py 100000 > /dev/null 2>&1
grep -v '^==' gpurun_out/r1z2_launches_pna.csv | tail -14 | cut -d, -f5,15 | cut -c1-90
