"""Workload for ncu: GIN forward on the bench batch, a few passes.  usage: python tools/prof_gin.py [tc1|tc2|tc3|0|1] [passes]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from flowgnn_b200.capi import Context
from flowgnn_b200.dataset import synthetic_molecules
from flowgnn_b200.weights import load_weights

w = load_weights("gin", os.path.join(ROOT, "tests", "golden", "weights", "GIN"))
big = synthetic_molecules(2048, "molhiv", seed=11).tile(41127)
v = sys.argv[1] if len(sys.argv) > 1 else "tc2"
with Context(0) as c:
    c.set_option("gin_tc1", int(v in ("tc1", "1")))
    c.set_option("gin_tc3", int(v == "tc3"))
    c.load_weights("gin", w)
    c.upload(big)
    for _ in range(int(sys.argv[2]) if len(sys.argv) > 2 else 2):
        c.compute("gin")
    c.synchronize()
