#!/usr/bin/env python
"""Generate tests/golden/golden_fixed_<dataset>.npz: raw int16 predictions (value = raw / 2^F) of the UNMODIFIED reference
kernels compiled against oracle/shim_fixed (the ap_fixed<16,I> emulation) on the packed graphs already in
tests/golden/.  Runs only where oracle/_ref/*_fixed.so exist (`make -C oracle ref`, needs /root/reference).

Usage:  python tools/make_fixed_fixtures.py [--force] [dataset ...]     (default: all three; existing arrays are kept unless --force)
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from flowgnn_b200.dataset import load_npz  # noqa: E402
from flowgnn_b200.weights import load_weights  # noqa: E402
from oracle.refbind import have_ref_fixed, run_reference_fixed  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
MODEL_DIR = {"gin": "GIN", "ginvn": "GIN", "dgn": "DGN"}
COUNTS = {"molhiv": 4113, "molpcba": 4113, "hep10k": 1000}


def main() -> None:
    only = {a for a in sys.argv[1:] if not a.startswith("--")}
    weights = {m: load_weights(m, os.path.join(GOLD, "weights", d)) for m, d in MODEL_DIR.items()}
    for ds, count in COUNTS.items():
        if only and ds not in only:
            continue
        batch = load_npz(os.path.join(GOLD, f"{ds}.npz")).slice(0, count)
        path = os.path.join(GOLD, f"golden_fixed_{ds}.npz")
        gold = dict(np.load(path)) if os.path.isfile(path) else {}
        for m in MODEL_DIR:
            if m in gold and len(gold[m]) == count and "--force" not in sys.argv:
                continue                                   # the emulation takes minutes per model: keep what is there
            if not have_ref_fixed(m):
                print(f"  {m}: oracle/_ref/libflowgnn_ref_{m}_fixed.so missing, skipped")
                continue
            t = time.time()
            bb = batch.with_virtual_node() if m == "ginvn" else batch
            gold[m] = run_reference_fixed(m, bb, weights[m])
            print(f"{ds} {m:6s} {time.time() - t:6.1f}s  raw range [{gold[m].min()}, {gold[m].max()}]", flush=True)
        np.savez(path, **gold)


if __name__ == "__main__":
    main()
