for m in ginvn gcn gat dgn pna; do
  timeout 150 python bench.py --model $m --steps 20 --warmup 3 --no-cpu-baseline --base-graphs 8192 > gpurun_out/r1y_bench_$m.json 2> gpurun_out/r1y_bench_$m.err
  echo "$m rc=$?"; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r1y_bench_$m.json"))
    print("$m value %.0f ms/step %.3f e2e %.0f layer_ms %.4f frac %.3f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["mean_launch_ms"], d["roofline"]["frac"]))
except Exception as e:
    print("$m fail", e); print(open("gpurun_out/r1y_bench_$m.err").read()[-600:])
PY
done
