#!/usr/bin/env python
"""The counterpart of the reference's run_experiments.sh (run_experiments.sh:28-66): for every `<dataset>:<model>`
pair, load the dataset (a reference zip, an extracted directory, a packed .npz/.fgb, or `synthetic-<shape>:<graphs>`),
run the model NUM_TRIALS times on the GPU and print ms/graph = mean device time / #graphs, exactly the figure the
reference derives from the XRT "Kernel Execution" average (run_experiments.sh:44-47).

    tools/run_experiments.py molhiv.zip:gin molhiv.zip:gat tests/golden/molpcba.npz:pna synthetic-hep10k:10000:ginvn
        [--weights DIR_WITH_MODEL_SUBDIRS] [--trials 25] [--device 0] [--out-dir DIR]

With --out-dir the per-graph predictions are written as <dataset>.<model>.B200_output.txt in the reference's
`g%d: %.8f` format (GIN/src/host.cc:213-222).
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from flowgnn_b200.capi import Context  # noqa: E402
from flowgnn_b200.dataset import (load_dataset_dir, load_dataset_zip, load_npz, load_packed, synthetic_hep,  # noqa: E402
                                  synthetic_molecules)
from flowgnn_b200.models import get_model  # noqa: E402
from flowgnn_b200.weights import load_weights  # noqa: E402

WEIGHT_DIRS = {"gin": "GIN", "ginvn": "GIN", "gcn": "GCN", "gat": "GAT", "pna": "PNA", "dgn": "DGN"}


def load_any(spec: str, need_eigen: bool):
    if spec.startswith("synthetic-"):
        shape, n = spec[len("synthetic-"):].split(":")
        return synthetic_hep(int(n)) if shape == "hep10k" else synthetic_molecules(int(n), shape, with_eigen=need_eigen)
    if spec.endswith(".zip"):
        return load_dataset_zip(spec, with_eigen=need_eigen)
    if spec.endswith(".npz"):
        return load_npz(spec)
    if spec.endswith(".fgb"):
        return load_packed(spec)
    return load_dataset_dir(spec, with_eigen=need_eigen)


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("experiments", nargs="+", help="<dataset>:<model>")
    ap.add_argument("--weights", default=os.path.join(ROOT, "tests", "golden", "weights"))
    ap.add_argument("--trials", type=int, default=25)          # NUM_TRIALS, GIN/src/host.h:8
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--out-dir", default=None)
    args = ap.parse_args()
    print(f"{'dataset':40s} {'model':6s} {'graphs':>8s} {'ms/batch':>10s} {'us/graph':>9s} {'graphs/s':>12s}")
    with Context(args.device) as ctx:
        for exp in args.experiments:
            ds, model = exp.rsplit(":", 1)
            spec = get_model(model)
            batch = load_any(ds, spec.uses_eigen)
            if spec.virtual_node:
                batch = batch.with_virtual_node()
            ctx.load_weights(model, load_weights(model, os.path.join(args.weights, WEIGHT_DIRS[spec.name])))
            ctx.upload(batch)
            ms = [ctx.compute(model, timed=True) for _ in range(args.trials + 1)][1:]       # first run warms up
            y = ctx.download()
            mean = float(np.mean(ms))
            print(f"{ds[-40:]:40s} {spec.name:6s} {batch.num_graphs:8d} {mean:10.3f} {1e3 * mean / batch.num_graphs:9.3f} "
                  f"{batch.num_graphs / (mean * 1e-3):12.0f}", flush=True)
            if args.out_dir:
                os.makedirs(args.out_dir, exist_ok=True)
                name = os.path.basename(ds.rstrip("/")).replace(":", "_")
                with open(os.path.join(args.out_dir, f"{name}.{spec.name}.B200_output.txt"), "w") as f:
                    for g, v in enumerate(y, 1):
                        f.write(f"g{g}: {v:.8f}\n")


if __name__ == "__main__":
    main()
