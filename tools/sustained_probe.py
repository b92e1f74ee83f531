"""Sustained back-to-back GIN forwards with nvidia-smi clock/power sampling.  usage: python tools/sustained_probe.py [tc1]"""
import os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from flowgnn_b200.capi import Context
from flowgnn_b200.dataset import synthetic_molecules
from flowgnn_b200.weights import load_weights
w = load_weights("gin", os.path.join(ROOT, "tests", "golden", "weights", "GIN"))
big = synthetic_molecules(2048, "molhiv", seed=11).tile(41127)
with Context(0) as c:
    c.set_option("gin_tc1", int(sys.argv[1]) if len(sys.argv) > 1 else 0)
    c.load_weights("gin", w); c.upload(big)
    for _ in range(5): c.compute("gin", timed=False)
    c.synchronize()
    mon = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,clocks.mem,power.draw,temperature.gpu,clocks_throttle_reasons.active", "--format=csv,noheader", "-lms", "50"], stdout=subprocess.PIPE, text=True)
    for rep in range(6):
        t0 = time.perf_counter()
        for _ in range(200): c.compute("gin", timed=False)
        c.synchronize()
        print(f"rep {rep}: ms/step {(time.perf_counter() - t0) * 5:.3f}", flush=True)
    mon.terminate()
    lines = mon.stdout.read().strip().splitlines()
    print(len(lines), "samples; every 8th:")
    for l in lines[::8]: print("  ", l)
