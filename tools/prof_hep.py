"""Workload for ncu: GIN-VN forward on hep10k-shaped graphs.  usage: python tools/prof_hep.py [mp_only 0|1] [passes]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from flowgnn_b200.capi import Context
from flowgnn_b200.dataset import synthetic_hep
from flowgnn_b200.weights import load_weights
w = load_weights("ginvn", os.path.join(ROOT, "tests", "golden", "weights", "GIN"))
big = synthetic_hep(4096, seed=11).tile(40000).with_virtual_node()
with Context(0) as c:
    c.set_option("mp_only", int(sys.argv[1]) if len(sys.argv) > 1 else 0)
    c.load_weights("ginvn", w)
    c.upload(big)
    for _ in range(int(sys.argv[2]) if len(sys.argv) > 2 else 2):
        c.compute("ginvn")
    c.synchronize()
