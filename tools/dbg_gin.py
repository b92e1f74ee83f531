"""GPU debugging aid: GIN through the CTA-pair kernel (default) and the single-CTA kernel (gin_tc1) against the
golden outputs, plus per-layer device times on the bench workload.  usage: python tools/dbg_gin.py [n_graphs]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from flowgnn_b200.capi import Context
from flowgnn_b200.dataset import load_npz, synthetic_molecules
from flowgnn_b200.weights import load_weights

gold = os.path.join(ROOT, "tests", "golden")
w = load_weights("gin", os.path.join(gold, "weights", "GIN"))
full = load_npz(os.path.join(gold, "molhiv.npz"))
g = np.load(os.path.join(gold, "golden_molhiv.npz"))["gin"]
sizes = [int(a) for a in sys.argv[1:]] or [1, 5, 64, 4113]
VARIANTS = os.environ.get("VARIANTS", "tc1 tc2 tc3").split()


def select(c, v):
    c.set_option("gin_tc1", int(v == "tc1"))
    c.set_option("gin_tc3", int(v == "tc3"))


with Context(0) as c:
    for n in sizes:
        b = full.slice(0, n)
        for v in VARIANTS:
            select(c, v)
            y = c.run("gin", b, w)
            err = np.abs(y - g[:n]) / np.maximum(1, np.abs(g[:n]))
            print(f"n={n} {v} max scaled err {err.max():.3e} at {int(err.argmax())} nonfinite {int((~np.isfinite(y)).sum())}", flush=True)
    if os.environ.get("DBG_HEP"):
        hb = load_npz(os.path.join(gold, "hep10k.npz"))
        hg = np.load(os.path.join(gold, "golden_hep10k.npz"))["gin"]
        for v in VARIANTS:
            select(c, v)
            y = c.run("gin", hb, w)
            err = np.abs(y - hg) / np.maximum(1, np.abs(hg))
            print(f"hep10k {v} max scaled err {err.max():.3e}", flush=True)
    big = synthetic_molecules(2048, "molhiv", seed=11).tile(41127)
    c.set_option("time_layers", 1)
    for v in VARIANTS:
        select(c, v)
        c.load_weights("gin", w)
        c.upload(big)
        for _ in range(3):
            c.compute("gin")
        ms = [c.compute("gin") for _ in range(10)]
        print(f"bench {v} step ms {np.mean(ms):.3f} layers {np.round(c.last_layer_ms(), 4)}", flush=True)
    # sustained (bench-like) comparison: 200 back-to-back forwards, host clock around the lot
    import time
    c.set_option("time_layers", 0)
    for v in VARIANTS + VARIANTS:
        select(c, v)
        for _ in range(5):
            c.compute("gin", timed=False)
        c.synchronize()
        t0 = time.perf_counter()
        for _ in range(200):
            c.compute("gin", timed=False)
        c.synchronize()
        print(f"sustained {v} ms/step {(time.perf_counter() - t0) * 5:.3f}", flush=True)
