#!/usr/bin/env python
"""Headline benchmark: graphs/s of the GIN dim100 forward on synthetic molhiv-shaped batches.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

A *step* is one pass of the hot path (on-device load_graph + embedding + 5 fused GIN layers + pool/head,
i.e. what the reference's single kernel enqueue does, GIN/src/GIN_compute.cc:44-98) over one batch of
41,127 synthetic molhiv-shaped graphs (BASELINE.json configs[1]).  Per-GPU work is fixed ("weak"
scaling): under torchrun every rank owns its own batch of that size (graphs shard by index, no
collective on the data path; NCCL only tallies graphs and the max time).

Printed (rank 0, one JSON line):
  value        graphs/s over all ranks, inputs resident in HBM, CUDA events on the context's stream
  e2e          the same through the reference-compatible entry point GIN_compute_graphs(...) with
               pinned HOST buffers: H2D of the batch, compute, D2H of the predictions, every step
  roofline     dominant kernel (gin_layer_kernel): algorithmic bytes per launch / mean launch time
  edge_gather  the mp_only variant of the same kernel (node transform = identity) -- the edge
               gather-scatter figure of BASELINE.json's metric
  cpu_baseline the reference's own kernel sources (oracle/_ref, fp32 flavour) on this box's host cores,
               bounded sample of the same workload (rank 0, N = 1 only)
  e2e_pageable the end-to-end call again with the caller's arrays in pageable memory, and after pinning them in place
  extras       BASELINE configs C4 / C5 on the same N GPUs: ONE 437,929-graph PNA batch sharded by graph index
               (strong scaling) and GIN-VN on hep10k-shaped graphs (weak scaling + batch-size sweep)

--impl reference times that CPU build alone (one process per core; its state is in file-scope
globals, */src/globals.cc, so it is not re-entrant) and prints the same line with "impl": "reference".
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
WEIGHT_DIRS = {"gin": "GIN", "ginvn": "GIN", "gcn": "GCN", "gat": "GAT", "pna": "PNA", "dgn": "DGN"}
#: graphs per GPU per step and generator shape, per model (BASELINE.json configs)
WORKLOADS = {
    "gin": ("molhiv", 41127), "gcn": ("molhiv", 41127), "gat": ("molhiv", 41127), "dgn": ("molhiv", 41127),
    "pna": ("molpcba", 437929), "ginvn": ("hep10k", 40000),
}
#: (D, L, attr words A, extra bytes per node X) of SURVEY.md 8(d): B_layer = 8*D*N + E*(8+4A) + X*N
ALGO = {"gin": (100, 5, 3, 0), "ginvn": (100, 5, 3, 0), "gcn": (100, 5, 3, 0), "gat": (64, 5, 0, 64), "pna": (80, 4, 0, 0),
        "dgn": (100, 4, 0, 0)}
LAYER_KERNEL = {"gin": "gin_layer_fused_kernel", "ginvn": "gin_layer_fused_kernel", "gcn": "tcf::fused_kernel<GcnFused> (one launch per step: gather -> tcgen05 GEMM -> bias)", "gat": "tcf::fused_kernel<GatFused> (one launch per layer: attention gather -> ELU -> tcgen05 GEMM [W_proj ; W_skip])",
                "pna": "pna_layer_fused_kernel (+ pna_exact_rows_kernel for the few non-finite rows)", "dgn": "tcf::fused_kernel<DgnFused> (+ dgn_exact_rows_kernel for the few non-finite rows)"}


def metric_name(model: str) -> str:
    """The SAME string on both arms (the driver matches them to form its ratios)."""
    return f"graphs/sec ({WORKLOADS[model][0]}-shaped, {model.upper()} forward)"


def layer_bytes(model: str, total_nodes: int, total_edges: int) -> int:
    D, _, A, X = ALGO[model]
    return 8 * D * total_nodes + total_edges * (8 + 4 * A) + X * total_nodes


def graph_bytes(model: str, num_graphs: int, total_nodes: int, total_edges: int) -> int:
    return 36 * total_nodes + ALGO[model][1] * layer_bytes(model, total_nodes, total_edges) + 4 * num_graphs


def make_workload(model: str, num_graphs: int, seed_offset: int = 0, base_graphs: int = 0):
    """Synthetic batch of the model's workload shape.  `base_graphs` > 0 generates
    that many distinct graphs and tiles them (used where the Python generator would take minutes)."""
    from flowgnn_b200.dataset import BASE_SEED, synthetic_hep, synthetic_molecules
    shape, _ = WORKLOADS[model]
    n_gen = min(num_graphs, base_graphs) if base_graphs else num_graphs
    seed = BASE_SEED + seed_offset
    if shape == "hep10k":
        b = synthetic_hep(n_gen, seed=seed)
    else:
        b = synthetic_molecules(n_gen, shape, seed=seed, with_eigen=(model == "dgn"))
    if n_gen < num_graphs:
        b = b.tile(num_graphs)
    if model == "ginvn":
        b = b.with_virtual_node()
    return b


# ---- clocks during the timed region (B200_PROFILING.md "clocks line") ---------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    REASONS = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")

    def __init__(self, gpu_id: str):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", gpu_id, f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, windows):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        time.sleep(0.06)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        for t, line in self.rows:
            if not any(a <= t <= b for a, b in windows):
                continue
            parts = [p.strip() for p in line.split(",")]
            try:
                sm.append(float(parts[0])); smax.append(float(parts[1])); power.append(float(parts[2]))
            except (ValueError, IndexError):
                continue
            for name, val in zip(self.REASONS, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "reasons": sorted(reasons), "samples": len(sm)}


# ---- the reference's CPU implementation, one process per core -----------------------------------------
_cpu_state = {}


def _cpu_init(model, kind, fast):
    from flowgnn_b200.weights import load_weights
    _cpu_state.update(model=model, kind=kind, fast=fast, w=load_weights(model, os.path.join(GOLDEN, "weights", WEIGHT_DIRS[model])))


def _cpu_run(shard):
    from oracle import refbind
    st = _cpu_state
    t0 = time.perf_counter()
    if st["kind"] == "reference":
        y = refbind.run_reference(st["model"], shard, st["w"], fast=st["fast"])
    else:
        y = refbind.run_port(st["model"], shard, st["w"], fast=st["fast"])
    return time.perf_counter() - t0, int(np.isfinite(y).sum())


class CpuReference:
    """The reference's CPU build of the path on `cores` worker processes (contiguous graph shards)."""

    def __init__(self, model: str, batch, cores: int):
        import multiprocessing as mp
        from flowgnn_b200.sharding import shard_of
        from oracle import refbind
        self.kind = "reference" if refbind.have_ref("ginvn" if model == "ginvn" else model) else "port"
        self.cores = cores
        self.shards = [shard_of(batch, r, cores)[0] for r in range(cores)]
        self.num_graphs = batch.num_graphs
        self.pool = mp.get_context("spawn").Pool(cores, initializer=_cpu_init, initargs=(model, self.kind, True))

    def step(self) -> float:
        t0 = time.perf_counter()
        self.pool.map(_cpu_run, self.shards, chunksize=1)
        return time.perf_counter() - t0

    def close(self):
        self.pool.close()
        self.pool.join()


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def bench_config(model, G, N, E, world):
    """The `config` object of the JSON line -- built by this one function for BOTH arms, so that they are equal key for key."""
    shape = WORKLOADS[model][0]
    return {"workload": f"{model.upper()} forward, {G} synthetic {shape}-shaped graphs per GPU per step",
            "batch": f"sum N = {N}, sum E = {E} on rank 0; trained weights shipped with the reference",
            "l2": "inputs larger than L2 (activations 2 x %.0f MB per GPU)" % (N * ALGO[model][0] * 4 / 1e6),
            "parallelism": f"graphs sharded by index over {world} GPU(s), no data-path collective"}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    model = args.model
    cores = host_cores()
    per_core = args.cpu_graphs_per_core
    sample = make_workload(model, per_core * cores, base_graphs=4096)
    ref = CpuReference(model, sample, cores)
    for _ in range(args.warmup):
        ref.step()
    t = [ref.step() for _ in range(args.steps)]
    ref.close()
    total = sum(t)
    value = sample.num_graphs * args.steps / total
    shape, full = WORKLOADS[model]
    # the B200 arm's batch (rank 0), only to name the same configuration: the timed sample above is its first graphs
    G_full = args.graphs or full
    whole = make_workload(model, G_full, seed_offset=0, base_graphs=args.base_graphs)
    sample_txt = (f"first {sample.num_graphs} graphs ({per_core} per core) of the synthetic {shape}-shaped workload per step, "
                  f"one process per core, timed around <MODEL>_compute_graphs only")
    line = {
        "impl": "reference", "metric": metric_name(model), "value": value, "unit": "graphs/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": bench_config(model, G_full, whole.total_nodes, whole.total_edges, int(os.environ.get("WORLD_SIZE", "1"))),
        "reference_sample_graphs_per_step": sample.num_graphs,
        "cpu_baseline": {"value": value, "unit": "graphs/s", "cores": cores, "kind": ref.kind, "sample": sample_txt},
        "e2e": {"value": value, "unit": "graphs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def pinned_copy(a):
    import torch
    t = torch.empty(a.shape, dtype=torch.from_numpy(a[:0]).dtype, pin_memory=True)
    v = t.numpy()
    v[...] = a
    return t, v



def bind_to_gpu_numa(dev_index: int):
    """Pin this process to the CPU cores of the GPU's NUMA node (if the cpuset allows any of them), BEFORE any pinned host
    memory is allocated: first-touch then places the staging buffers next to the GPU's PCIe root.  Returns a record for the
    JSON line.  (r1: all 8 ranks sat on NUMA node 0 and the end-to-end arm lost 33 % at 8 GPUs.)"""
    import torch
    rec = {"numa_node": None, "cpus_before": None, "cpus_after": None}
    try:
        allowed = os.sched_getaffinity(0)
        rec["cpus_before"] = len(allowed)
        pr = torch.cuda.get_device_properties(dev_index)
        bdf = f"{getattr(pr, 'pci_domain_id', 0):04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        base = f"/sys/bus/pci/devices/{bdf}"
        with open(base + "/numa_node") as f:
            node = int(f.read().strip())
        rec["numa_node"] = node
        if node < 0:
            return rec
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            txt = f.read().strip()
        cpus = set()
        for part in txt.split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        want = cpus & allowed
        if want and want != allowed:
            os.sched_setaffinity(0, want)
        rec["cpus_after"] = len(os.sched_getaffinity(0))
        rec["local_cpus_in_cpuset"] = len(want)
    except (OSError, ValueError, AttributeError) as e:
        rec["error"] = str(e)[:80]
    return rec


def measure_resident(ctx, model, steps, warmup, stream, barrier, grouped):
    """W untimed + K timed passes of the hot path over the batch resident in `ctx`; returns (ms of the K steps on this rank,
    per-layer-launch ms list, launches)."""
    import torch
    L = ALGO[model][1]
    for _ in range(warmup):
        ctx.compute(model, timed=True)
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    layer_ms = []
    for _ in range(steps):
        ctx.compute(model, timed=True)
        lm = ctx.last_layer_ms()
        layer_ms.extend([lm[0] / L] * L if grouped and len(lm) == 1 else lm)
    ev1.record(stream)
    barrier()
    return ev0.elapsed_time(ev1), layer_ms, ctx.last_launch_count * steps


def run_extra_configs(args, rank, world, local, dev, barrier, max_over_ranks, sum_over_ranks, hbm_peak):
    """BASELINE configs C3, C4 and C5 inside the same JSON line (the driver only runs `bench.py --gpus N`):
      gat_molhiv           GAT (5 layers, 4 heads) on molhiv-shaped graphs, 41,127 per GPU (weak);
      pna_molpcba_sharded  ONE batch of 437,929 molpcba-shaped graphs cut by graph index over the N ranks
                           (flowgnn_b200.sharding.shard_of -- strong scaling; NCCL only tallies);
      ginvn_hep10k         GIN-VN on hep10k-shaped graphs, 40,000 per GPU (weak), plus a batch-size sweep.
    Device-timed, max over ranks; fewer steps than the headline (each PNA step is ~50 ms on one GPU)."""
    import torch
    from flowgnn_b200.capi import Context
    from flowgnn_b200.sharding import shard_of
    from flowgnn_b200.weights import load_weights
    steps, warmup = max(3, min(args.steps, 8)), 3
    out = {}

    def one(model, batch, grouped):
        w = load_weights(model, os.path.join(GOLDEN, "weights", WEIGHT_DIRS[model]))
        ctx = Context(local)
        try:
            ctx.load_weights(model, w)
            ctx.upload(batch)
            ctx.set_option("time_layers", 2 if grouped else 1)
            stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
            ms, layer_ms, launches = measure_resident(ctx, model, steps, warmup, stream, barrier, grouped)
            y = ctx.download()
        finally:
            ctx.close()
        return ms, layer_ms, launches, y

    def end_to_end(model, batch, y_dev, graphs_all_ranks):
        """The same batch through <MODEL>_compute_graphs with the caller's arrays in ORDINARY (pageable) host memory -- what the
        reference's host owns (common/includes/xcl2/xcl2.hpp:61-76): upload, on-device load_graph, forward and download inside the
        timed call; predictions must equal the device-resident run bit for bit."""
        from flowgnn_b200.capi import ReferenceCall, last_transfer_bytes
        w = load_weights(model, os.path.join(GOLDEN, "weights", WEIGHT_DIRS[model]))
        call = ReferenceCall(model, batch, w)
        esteps = max(3, steps // 2)
        for _ in range(2):
            y = call.run()
        barrier()
        t0 = time.perf_counter()
        for _ in range(esteps):
            y = call.run()
        torch.cuda.synchronize()
        dt = max_over_ranks(time.perf_counter() - t0)
        barrier()
        same = bool(np.array_equal(y.view(np.int32), y_dev.view(np.int32)))
        if not same:
            raise SystemExit(f"bench.py: {model}: end-to-end and device-resident predictions differ")
        h2d, d2h = last_transfer_bytes()
        spec_arrays = [batch.nums_of_nodes, batch.nums_of_edges, batch.node_feature, batch.edge_list, batch.edge_attr, batch.node_eigen]
        return {"value": graphs_all_ranks * esteps / dt, "unit": "graphs/s", "ms_per_step": 1e3 * dt / esteps, "steps": esteps,
                "caller_memory": "pageable", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "caller_input_bytes_per_step": int(sum(a.nbytes for a in spec_arrays if a is not None)),
                "bit_identical_to_device_resident": same}

    # ---- C3: GAT on molhiv-shaped graphs, weak scaling -------------------------------------------------------
    model = "gat"
    batch = make_workload(model, WORKLOADS[model][1], seed_offset=rank, base_graphs=4096)
    ms, layer_ms, launches, y = one(model, batch, False)
    ms = max_over_ranks(ms)
    done = sum_over_ranks(float(batch.num_graphs))
    lb = layer_bytes(model, batch.total_nodes, batch.total_edges)
    mean_layer = max_over_ranks(float(np.mean(layer_ms[:-1])))                  # the four fused layer launches (the fifth interval is the final gather)
    out["gat_molhiv"] = {
        "metric": metric_name(model), "value": done * steps / (ms * 1e-3), "unit": "graphs/s", "scaling": "weak",
        "workload": f"GAT forward (5 layers, 4 heads x 16), {batch.num_graphs} synthetic molhiv-shaped graphs per GPU per step",
        "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": ms / steps, "gpu_launches": int(launches),
        "finite_outputs": sum_over_ranks(float(np.isfinite(y).sum())),
        "roofline": {"bound": "hbm", "kernel": LAYER_KERNEL[model], "achieved": lb / (mean_layer * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                     "frac": lb / (mean_layer * 1e-3) / 1e9 / hbm_peak, "mean_layer_ms": mean_layer, "algorithmic_bytes_per_layer": lb},
        "e2e": end_to_end(model, batch, y, done),
    }
    del batch

    # ---- C4: PNA, one molpcba-sized batch sharded by graph index ------------------------------------------
    model = "pna"
    full = make_workload(model, WORKLOADS[model][1], base_graphs=8192)          # same on every rank
    shard, g0, g1 = shard_of(full, rank, world)
    ms, layer_ms, launches, y = one(model, shard, False)
    ms = max_over_ranks(ms)
    done = sum_over_ranks(float(g1 - g0))
    finite = sum_over_ranks(float(np.isfinite(y).sum()))
    lb = layer_bytes(model, full.total_nodes, full.total_edges)                # whole job
    mean_layer = max_over_ranks(float(np.mean(layer_ms)))
    out["pna_molpcba_sharded"] = {
        "metric": metric_name(model), "value": done * steps / (ms * 1e-3), "unit": "graphs/s", "scaling": "strong",
        "workload": f"PNA forward, ONE batch of {full.num_graphs} synthetic molpcba-shaped graphs (sum N = {full.total_nodes}, "
                    f"sum E = {full.total_edges}) cut into {world} contiguous graph ranges (shard_of)",
        "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": ms / steps, "graphs_done_per_step": done, "finite_outputs": finite,
        "gpu_launches": int(launches), "layer_ms_max_over_ranks": mean_layer,
        "roofline": {"bound": "hbm", "kernel": LAYER_KERNEL[model], "achieved": lb / world / (mean_layer * 1e-3) / 1e9, "peak": hbm_peak,
                     "unit": "GB/s per GPU", "frac": lb / world / (mean_layer * 1e-3) / 1e9 / hbm_peak, "algorithmic_bytes_per_layer_whole_job": lb},
        "e2e": end_to_end(model, shard, y, done),
    }
    del full, shard

    # ---- C5: GIN-VN on hep10k-shaped graphs, weak scaling + batch-size sweep -------------------------------
    model = "ginvn"
    sweep = []
    for G in (2500, 10000, WORKLOADS[model][1]):
        batch = make_workload(model, G, seed_offset=rank, base_graphs=2048)
        ms, layer_ms, launches, y = one(model, batch, True)
        ms = max_over_ranks(ms)
        done = sum_over_ranks(float(G))
        lb = layer_bytes(model, batch.total_nodes, batch.total_edges)
        mean_layer = max_over_ranks(float(np.mean(layer_ms)))
        rec = {"graphs_per_gpu": G, "value": done * steps / (ms * 1e-3), "ms_per_step": ms / steps, "layer_ms": mean_layer,
               "roofline_frac": lb / (mean_layer * 1e-3) / 1e9 / hbm_peak, "gpu_launches": int(launches),
               "finite_outputs": sum_over_ranks(float(np.isfinite(y).sum()))}
        sweep.append(rec)
    e2e_top = end_to_end(model, batch, y, done)          # the 40,000-graph batch of the last sweep point
    top = sweep[-1]
    out["ginvn_hep10k"] = {
        "metric": metric_name(model), "value": top["value"], "unit": "graphs/s", "scaling": "weak",
        "workload": f"GIN-VN forward, {WORKLOADS[model][1]} synthetic hep10k-shaped graphs per GPU per step (virtual node added)",
        "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": top["ms_per_step"],
        "roofline": {"bound": "hbm", "kernel": "GIN layer on dense graphs", "frac": top["roofline_frac"], "peak": hbm_peak, "unit": "GB/s",
                     "mean_layer_ms": top["layer_ms"]},
        "sweep": sweep,
        "e2e": e2e_top,
    }
    return out


def run_b200_arm(args):
    import torch
    import torch.distributed as dist
    from flowgnn_b200.capi import Context, ReferenceCall
    from flowgnn_b200.dataset import Batch
    from flowgnn_b200.models import get_model
    from flowgnn_b200.weights import load_weights

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 arm has no CPU fallback (use --impl reference for the CPU build)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    affinity = bind_to_gpu_numa(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local])
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    model = args.model
    shape, default_graphs = WORKLOADS[model]
    G = args.graphs or default_graphs
    batch = make_workload(model, G, seed_offset=rank, base_graphs=args.base_graphs)
    weights = load_weights(model, os.path.join(GOLDEN, "weights", WEIGHT_DIRS[model]))
    N, E = batch.total_nodes, batch.total_edges

    # pinned host copies of the batch: what a caller of the reference entry point owns (GIN/src/host.cc:141-182)
    keep = []
    arrays = {}
    for name in ("nums_of_nodes", "nums_of_edges", "node_feature", "edge_list", "edge_attr", "node_eigen"):
        a = getattr(batch, name)
        if a is None:
            arrays[name] = None
            continue
        t, v = pinned_copy(a)
        keep.append(t)
        arrays[name] = v
    hbatch = Batch(arrays["nums_of_nodes"], arrays["nums_of_edges"], arrays["node_feature"], arrays["edge_list"], arrays["edge_attr"],
                   arrays["node_eigen"], name=batch.name)                    # views of the pinned buffers, no copies
    assert hbatch.node_feature.ctypes.data == arrays["node_feature"].ctypes.data
    spec = get_model(model)
    used = ["nums_of_nodes", "nums_of_edges", "node_feature", "edge_list"] + (["edge_attr"] if spec.uses_edge_attr else []) + \
           (["node_eigen"] if spec.uses_eigen else [])
    h2d = sum(arrays[k].nbytes for k in used)
    d2h = 4 * G

    ctx = Context(local)
    ctx.load_weights(model, weights)
    ctx.upload(hbatch)
    # GIN: one event pair around the five layer launches (events between them would break programmatic dependent launch);
    # the per-launch time is that interval / 5 and includes the (overlapped) launch gaps
    grouped = model in ("gin", "ginvn")
    ctx.set_option("time_layers", 2 if grouped else 1)
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    gpu_id = str(torch.cuda.get_device_properties(dev).uuid)
    if not gpu_id.startswith("GPU-"):
        gpu_id = "GPU-" + gpu_id

    # ---- device-resident arm ------------------------------------------------------------------------
    for _ in range(args.warmup):
        ctx.compute(model, timed=True)
    sampler = ClockSampler(gpu_id) if rank == 0 else None
    windows = []
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0 = time.perf_counter()
    ev0.record(stream)
    layer_ms = []
    for _ in range(args.steps):
        ctx.compute(model, timed=True)
        lm = ctx.last_layer_ms()
        layer_ms.extend([lm[0] / ALGO[model][1]] * ALGO[model][1] if grouped and len(lm) == 1 else lm)
    ev1.record(stream)
    barrier()
    windows.append((w0, time.perf_counter()))
    launches = ctx.last_launch_count * args.steps
    dev_ms = max_over_ranks(ev0.elapsed_time(ev1))
    total_graphs = sum_over_ranks(float(G))
    value = total_graphs * args.steps / (dev_ms * 1e-3)
    y_dev = ctx.download()

    # ---- end-to-end arm: the reference-compatible entry point with host buffers ---------------------
    call = ReferenceCall(model, hbatch, weights)         # marshalled once: each run() is the bare C-ABI call
    for _ in range(args.warmup):
        y_e2e = call.run()
    barrier()
    w0 = time.perf_counter()
    for _ in range(args.steps):
        y_e2e = call.run()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - w0
    windows.append((w0, time.perf_counter()))
    barrier()
    e2e_s = max_over_ranks(e2e_s)
    e2e_value = total_graphs * args.steps / e2e_s
    # what the call actually moved over PCIe (counted inside the library around its copies): the entry point narrows the int32 words of
    # the reference layout on the host (option host_stage) when that is faster than copying them as they are
    from flowgnn_b200.capi import last_transfer_bytes
    caller_bytes = h2d
    h2d, d2h = last_transfer_bytes()
    if not np.array_equal(y_dev.view(np.int32), y_e2e.view(np.int32)):
        raise SystemExit("bench.py: device-resident and end-to-end predictions differ")
    clocks = sampler.stop(windows) if sampler else None

    # ---- the same call with PAGEABLE caller buffers (what the reference host owns: aligned_allocator vectors,
    # common/includes/xcl2/xcl2.hpp:61-76), and again after pinning them in place through flowgnn_b200_pin_host ----
    e2e_pageable = None
    if not args.no_pageable:
        from flowgnn_b200.capi import pin_host, unpin_host
        pbatch = Batch(batch.nums_of_nodes, batch.nums_of_edges, batch.node_feature.copy(), batch.edge_list.copy(),
                       None if batch.edge_attr is None else batch.edge_attr.copy(), None if batch.node_eigen is None else batch.node_eigen.copy(),
                       name=batch.name)
        pcall = ReferenceCall(model, pbatch, weights)
        ksteps = max(3, args.steps // 2)

        def timed_calls():
            for _ in range(3):
                y = pcall.run()
            barrier()
            t0 = time.perf_counter()
            for _ in range(ksteps):
                y = pcall.run()
            torch.cuda.synchronize()
            dt = max_over_ranks(time.perf_counter() - t0)
            barrier()
            return total_graphs * ksteps / dt, y.copy()

        v_page, y_page = timed_calls()
        big = [a for a in (pbatch.node_feature, pbatch.edge_list, pbatch.edge_attr, pbatch.node_eigen) if a is not None]
        for a in big:
            pin_host(a)
        v_reg, y_reg = timed_calls()
        for a in big:
            unpin_host(a)
        if not (np.array_equal(y_page.view(np.int32), y_dev.view(np.int32)) and np.array_equal(y_reg.view(np.int32), y_dev.view(np.int32))):
            raise SystemExit("bench.py: pageable-buffer predictions differ")
        e2e_pageable = {"value": v_page, "unit": "graphs/s", "steps": ksteps,
                        "after_flowgnn_b200_pin_host": v_reg,
                        "note": "same C-ABI call; caller arrays in ordinary (pageable) memory, then page-locked in place"}

    # ---- end to end from the PACKED host layout (Part 2: flowgnn_b200_upload_batch_packed, the layout of the packed dataset files:
    # uint8 features, uint16 edge ids, uint8 bond attributes): upload + on-device load_graph + forward + download per step, pinned host
    # buffers, no host threads involved -- what a loader that keeps its dataset packed pays, and what scales with the number of GPUs
    # when the int32 words of the reference ABI saturate the host's memory system
    e2e_packed = None
    if not args.no_packed and int(batch.node_feature.max(initial=0)) < 256 and int(batch.edge_list.max(initial=0)) < 65536:
        packed = {"nf": pinned_copy(batch.node_feature.astype(np.uint8)), "el": pinned_copy(batch.edge_list.astype(np.uint16)),
                  "ea": pinned_copy(batch.edge_attr.astype(np.uint8)) if spec.uses_edge_attr else (None, None)}
        from flowgnn_b200.capi import PackedCall
        pcall_packed = PackedCall(model, hbatch, weights, packed["nf"][1], packed["el"][1], packed["ea"][1])

        for _ in range(3):
            pcall_packed.run()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            y_packed = pcall_packed.run()
        torch.cuda.synchronize()
        dt = max_over_ranks(time.perf_counter() - t0)
        barrier()
        if not np.array_equal(y_packed.view(np.int32), y_dev.view(np.int32)):
            raise SystemExit("bench.py: packed-upload predictions differ")
        pbytes, pd2h = last_transfer_bytes()
        e2e_packed = {"value": total_graphs * args.steps / dt, "unit": "graphs/s", "ms_per_step": 1e3 * dt / args.steps, "steps": args.steps,
                      "h2d_bytes_per_step": int(pbytes), "d2h_bytes_per_step": int(pd2h),
                      "note": "flowgnn_b200_compute_graphs_packed: the chunked host-pointer pipeline of the entry points fed with pinned host buffers "
                              "in the packed dataset layout (u8 / u16 / u8) instead of the reference ABI's int32 words; no host threads narrow anything"}

    # ---- roofline of the dominant kernel (per-layer CUDA events recorded inside the timed region) ---
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(peaks_path):
        with open(peaks_path) as f:
            hbm_peak, peak_src = float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        hbm_peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    lb_full = layer_bytes(model, N, E)
    lb = lb_full
    mean_layer_ms = float(np.mean(layer_ms)) if layer_ms else float("nan")
    achieved = lb / (mean_layer_ms * 1e-3) / 1e9
    if model in ("gin", "ginvn") and layer_ms:
        # the last GIN launch has the prediction head fused into its epilogue: it reads h (4 D N) and the edge records and
        # writes 4 bytes per node instead of h'.  The roofline figure is the sum of the launches' algorithmic bytes over the
        # sum of their times; "algorithmic_bytes_per_launch" and "mean_launch_ms" stay per-launch means.
        L = ALGO[model][1]
        lb_last = 4 * ALGO[model][0] * N + E * (8 + 4 * ALGO[model][2]) + 4 * N
        lb = ((L - 1) * lb_full + lb_last) / L
        achieved = lb / (mean_layer_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.isfile(tpath):
        with open(tpath) as f:
            traffic = json.load(f).get(LAYER_KERNEL[model])
    # dense graphs (>= 6 in-edges per node: hep10k) run a GIN layer as two launches, the staged gather and the node MLP
    dense = model in ("gin", "ginvn") and E >= 6 * N
    layer_kernel = "gin_gather_staged_kernel + gin_layer_fused_kernel (two launches per layer)" if dense else LAYER_KERNEL[model]
    if dense:
        traffic = None
    roofline = {"bound": "hbm", "kernel": layer_kernel, "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                "frac": achieved / hbm_peak, "traffic": traffic,
                "traffic_source": "profiles/traffic.json: dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture of this kernel, "
                                  "per launch; a committed constant, NOT measured in this run" if traffic else None,
                "peak_source": peak_src, "algorithmic_bytes_per_launch": lb,
                "mean_launch_ms": mean_layer_ms, "share_of_step": mean_layer_ms * ALGO[model][1] * args.steps / ev0.elapsed_time(ev1)}

    # The second floor of the GIN layer kernel: the tensor-core work it ISSUES.  Every 128-row tile runs GEMM1 [128 x 104] x [104 x 208]
    # and GEMM2 [128 x 208] x [208 x 128] (K and N padded to the MMA shapes, biases as an extra K column) three times (bf16 hi*hi +
    # lo*hi + hi*lo keeps the fp32 contract) -- against the measured dense bf16 rate of this GPU that is a longer floor than the HBM one.
    if model == "gin" and not dense and layer_ms and ctx.tile_count > 0:
        tf_peak = 0.0
        if os.path.isfile(peaks_path):
            with open(peaks_path) as f:
                tf_peak = float(json.load(f).get("bf16_tflops", 0.0))
        tiles = ctx.tile_count
        issued = 3 * 2.0 * (104 * 208 + 208 * 128) * 128 * tiles
        useful = 2.0 * (100 * 200 + 200 * 100) * N
        if tf_peak > 0:
            roofline["tensor"] = {"issued_flops_per_launch": issued, "fp32_equivalent_flops_per_launch": useful, "tiles": int(tiles),
                                  "achieved": issued / (mean_layer_ms * 1e-3) / 1e12, "peak": tf_peak, "unit": "TFLOP/s (dense bf16, measured)",
                                  "frac": issued / (mean_layer_ms * 1e-3) / 1e12 / tf_peak,
                                  "floor_ms": issued / (tf_peak * 1e12) * 1e3, "hbm_floor_ms": lb / (hbm_peak * 1e9) * 1e3,
                                  "note": "the kernel's longer floor: 3-product bf16 split of an fp32 MLP on padded MMA shapes; bound stays "
                                          "'hbm' for the north-star figure, this object says how far the issued tensor work is from its own peak"}

    edge_gather = None
    edge_gather_side = None
    if model in ("gin", "ginvn"):
        def mp_only_run(mode):
            ctx.set_option("mp_only", mode)
            ctx.set_option("time_layers", 1)
            for _ in range(args.warmup):
                ctx.compute(model, timed=True)
            ms = []
            for _ in range(max(5, args.steps // 2)):
                ctx.compute(model, timed=True)
                ms.extend(ctx.last_layer_ms())
            ctx.set_option("mp_only", 0)
            return float(np.mean(ms))
        # the edge-gather figure is measured on the mp_only variant of the SAME layer kernel (node transform =
        # identity, same loads and stores): gin_layer_fused_kernel with its MMA / epilogue warps idle
        t = mp_only_run(1)
        a = lb_full / (t * 1e-3) / 1e9
        edge_gather = {"kernel": "gin_layer_fused_kernel (mp_only: node transform = identity)", "achieved": a, "peak": hbm_peak, "unit": "GB/s",
                       "frac": a / hbm_peak, "mean_launch_ms": t, "algorithmic_bytes_per_launch": lb_full}
        # information only: the stand-alone gather kernels of round 1 (row-per-warp / staged), never on the product path of molecules
        t = mp_only_run(2)
        a = lb_full / (t * 1e-3) / 1e9
        gk = "gin_gather_staged_kernel" if dense else "gin_gather_kernel"
        edge_gather_side = {"kernel": gk + " (stand-alone gather kernel, information only)", "achieved": a, "peak": hbm_peak, "unit": "GB/s",
                            "frac": a / hbm_peak, "mean_launch_ms": t}
    ctx.close()

    extras = None
    if not args.no_extras and model == "gin":
        extras = run_extra_configs(args, rank, world, local, dev, barrier, max_over_ranks, sum_over_ranks, hbm_peak)

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = host_cores()
        per_core = args.cpu_graphs_per_core * 8
        sample = batch.slice(0, min(G, per_core * cores))
        ref = CpuReference(model, sample, cores)
        ref.step() if sample.num_graphs <= 64 * cores else None
        t = ref.step()
        ref.close()
        cpu_baseline = {"value": sample.num_graphs / t, "unit": "graphs/s", "cores": cores, "kind": ref.kind,
                        "sample": f"first {sample.num_graphs} graphs of this run's batch, one process per core, one pass ({t:.1f} s), "
                                  f"timed around <MODEL>_compute_graphs only"}

    if rank == 0:
        line = {
            "metric": metric_name(model), "value": value, "unit": "graphs/s", "timer": "CUDA events on the context's stream (device-timed)",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": bench_config(model, G, N, E, world),
            "e2e": {"value": e2e_value, "unit": "graphs/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "caller_input_bytes_per_step": int(caller_bytes),
                    "ms_per_step": 1e3 * e2e_s / args.steps, "timer": "host clock around the synchronous C-ABI calls",
                    "note": "host buffers in the reference's int32 layout (pinned); h2d/d2h = bytes the call moved over PCIe, counted around "
                            "the library's copies (inputs narrowed to u8/u16 by host threads inside the timed call when host_stage is on)"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roofline,
            "edge_gather": edge_gather,
            "edge_gather_side_kernel": edge_gather_side,
            "algorithmic_bytes_per_graph": graph_bytes(model, G, N, E) / G,
            "cpu_baseline": cpu_baseline,
            "e2e_pageable": e2e_pageable,
            "e2e_packed": e2e_packed,
            "extras": extras,
            "affinity": affinity,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=("b200", "reference"), default="b200")
    ap.add_argument("--model", choices=tuple(WORKLOADS), default="gin")
    ap.add_argument("--graphs", type=int, default=0, help="graphs per GPU per step (default: the model's BASELINE config)")
    ap.add_argument("--base-graphs", type=int, default=0, help="generate this many distinct graphs and tile them")
    ap.add_argument("--cpu-graphs-per-core", type=int, default=256, help="reference arm: graphs per core per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-packed", action="store_true", help="skip the packed-layout end-to-end arm (e2e_packed)")
    ap.add_argument("--no-extras", action="store_true", help="skip the C4 (PNA sharded) / C5 (GIN-VN hep10k) lines")
    ap.add_argument("--no-pageable", action="store_true", help="skip the pageable-caller end-to-end line")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.model == "pna" and not args.base_graphs:
        args.base_graphs = 32768
    # The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints "NCCL version ..." when the box sets
    # NCCL_DEBUG), so file descriptor 1 points at stderr while the arms run and sys.stdout keeps the real stdout.
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = real_stdout
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_b200_arm(args)
    sys.stdout.flush()


if __name__ == "__main__":
    main()
