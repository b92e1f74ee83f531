"""The C++ driver flowgnn_b200/host/host_b200 (the counterpart of the reference's ./host): reads the reference's own
file formats, calls the reference-compatible entry points, writes `g%d: %.8f` lines."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ALL_MODELS, GOLDEN, MODEL_WEIGHT_DIR, ROOT, assert_parity

HOST = os.path.join(ROOT, "flowgnn_b200", "host", "host_b200")


def _write_dataset(tmp_path, datasets, count=48):
    root = str(tmp_path / "molhiv")
    b = datasets["molhiv"].slice(0, count)
    b.save_reference_layout(root)
    return root, b


def test_reference_layout_round_trip(tmp_path, datasets):
    """The writer of the reference's per-graph files and the reader of flowgnn_b200.dataset agree (incl. the eigen text)."""
    from flowgnn_b200.dataset import load_dataset_dir
    root, b = _write_dataset(tmp_path, datasets, 12)
    back = load_dataset_dir(root, with_eigen=True)
    assert np.array_equal(back.nums_of_nodes, b.nums_of_nodes) and np.array_equal(back.edge_list, b.edge_list)
    assert np.array_equal(back.node_feature, b.node_feature) and np.array_equal(back.edge_attr, b.edge_attr)
    assert np.array_equal(back.node_eigen, b.node_eigen)


def test_host_driver_is_built_and_explains_itself():
    if not os.path.isfile(HOST):
        pytest.skip("host_b200 not built (run __graft_entry__.build())")
    r = subprocess.run([HOST], capture_output=True, text=True)
    assert r.returncode != 0 and "usage: host_b200" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("model", ALL_MODELS)
def test_host_driver_matches_reference_outputs(model, tmp_path, datasets, golden):
    assert os.path.isfile(HOST), "host_b200 not built"
    root, b = _write_dataset(tmp_path, datasets)
    out = str(tmp_path / "B200_output.txt")
    wdir = os.path.join(GOLDEN, "weights", MODEL_WEIGHT_DIR[model])
    r = subprocess.run([HOST, model, root, wdir, "--trials", "2", "--out", out], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr + r.stdout
    lines = open(out).read().split("\n")[:-1]
    assert len(lines) == b.num_graphs and lines[0].startswith("g1: ") and lines[-1].startswith(f"g{b.num_graphs}: ")
    got = np.array([float(x.split(": ")[1]) for x in lines], dtype=np.float32)
    assert_parity(got, golden["molhiv"][model][:b.num_graphs], what=f"host_b200 {model}")


@pytest.mark.gpu
@pytest.mark.parametrize("model", ["gin", "gcn", "gat", "pna", "dgn"])
def test_host_driver_packed_layout_gives_the_same_file(model, tmp_path, datasets):
    """--layout packed: the host narrows the batch once after loading and calls flowgnn_b200_compute_graphs_packed; the prediction
    file must be byte-identical to the one written through <MODEL>_compute_graphs."""
    assert os.path.isfile(HOST), "host_b200 not built"
    root, b = _write_dataset(tmp_path, datasets)
    wdir = os.path.join(GOLDEN, "weights", MODEL_WEIGHT_DIR[model])
    outs = {}
    for layout in ("int32", "packed"):
        out = str(tmp_path / f"out_{layout}.txt")
        r = subprocess.run([HOST, model, root, wdir, "--trials", "2", "--out", out, "--layout", layout], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr + r.stdout
        assert f"{layout} host buffers in" in r.stdout, r.stdout
        outs[layout] = open(out, "rb").read()
    assert outs["int32"] == outs["packed"] and outs["packed"].count(b"\n") == b.num_graphs


def _gpu_count():
    import torch
    return torch.cuda.device_count()


@pytest.mark.gpu
@pytest.mark.parametrize("model", ["gin", "gat", "dgn"])
def test_host_driver_shards_over_gpus_and_reads_the_packed_format(model, tmp_path, datasets, golden):
    """`--gpus N`: one host thread per GPU, contiguous graph ranges, predictions in disjoint slices of out[G], NCCL only for
    the tally (SUM of graphs, MAX of device time).  Reads the packed single-file dataset (.fgb).  The result must equal the
    reference's single-batch output -- for GAT including its missing node offset, which a range of a larger
    batch must reproduce by reading the batch's first feature rows."""
    assert os.path.isfile(HOST), "host_b200 not built"
    b = datasets["molhiv"].slice(0, 300)
    packed = str(tmp_path / "molhiv300.fgb")
    b.save_packed(packed)
    out = str(tmp_path / "B200_output.txt")
    wdir = os.path.join(GOLDEN, "weights", MODEL_WEIGHT_DIR[model])
    for gpus in sorted({1, min(2, _gpu_count())}):
        r = subprocess.run([HOST, model, packed, wdir, "--trials", "3", "--out", out, "--gpus", str(gpus)], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr + r.stdout
        assert f"300 graphs on {gpus} GPU(s)" in r.stdout and "300 graphs in" in r.stdout, r.stdout       # the NCCL tally saw every graph
        got = np.array([float(x.split(": ")[1]) for x in open(out).read().split("\n")[:-1]], dtype=np.float32)
        assert_parity(got, golden["molhiv"][model][:300], what=f"host_b200 {model} --gpus {gpus}")
    # a sub-range of the packed file, Part-1 path (no --gpus)
    r = subprocess.run([HOST, model, packed, wdir, "--trials", "1", "--out", out, "--first", "11", "--graphs", "40"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr + r.stdout
    lines = open(out).read().split("\n")[:-1]
    assert len(lines) == 40 and lines[0].startswith("g11: ")
    if model != "gat":          # GAT's offset quirk makes a sub-range read other feature rows than the full batch did
        got = np.array([float(x.split(": ")[1]) for x in lines], dtype=np.float32)
        assert_parity(got, golden["molhiv"][model][10:50], what=f"host_b200 {model} graphs 11..50")


def _nccl_worker(rank, world, port, out_path):
    import sys
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from flowgnn_b200.capi import Context
    from flowgnn_b200.dataset import load_npz
    from flowgnn_b200.sharding import gather_predictions, shard_of, tally
    from flowgnn_b200.weights import load_weights
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    batch = load_npz(os.path.join(GOLDEN, "molpcba.npz"))
    res = {}
    with Context(rank) as ctx:
        for model in ("gin", "pna"):
            w = load_weights(model, os.path.join(GOLDEN, "weights", MODEL_WEIGHT_DIR[model]))
            shard, g0, g1 = shard_of(batch, rank, world)
            ctx.load_weights(model, w)
            ctx.upload(shard)
            ms = ctx.compute(model, timed=True)
            local = ctx.download()
            total, tmax = tally(shard.num_graphs, ms, dist, dev)
            res[model] = gather_predictions(local, g0, batch.num_graphs, dist, dev)
            res[model + "_total"] = total
    if rank == 0:
        np.savez(out_path, **res)
    dist.barrier(device_ids=[rank])
    dist.destroy_process_group()


@pytest.mark.gpu
def test_two_gpu_sharded_run_reassembles_the_single_gpu_result(tmp_path, datasets, golden, weights):
    """SURVEY.md 8e on real GPUs: ONE batch cut by graph index over two ranks (flowgnn_b200.sharding.shard_of), each rank
    runs its range on its own B200, the predictions are reassembled over NCCL -- bit-identical to the single-GPU run."""
    if _gpu_count() < 2:
        pytest.skip("needs two GPUs")
    import socket
    import torch.multiprocessing as mp
    from flowgnn_b200.capi import Context
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = str(tmp_path / "r0.npz")
    mp.spawn(_nccl_worker, args=(2, port, out), nprocs=2, join=True)
    z = np.load(out)
    with Context(0) as ctx:
        for model in ("gin", "pna"):
            want = ctx.run(model, datasets["molpcba"], weights[model])
            assert int(z[model + "_total"]) == datasets["molpcba"].num_graphs
            assert np.array_equal(z[model].view(np.int32), want.view(np.int32)), model
            assert_parity(z[model], golden["molpcba"][model], what=f"{model} two-GPU sharded")


@pytest.mark.gpu
def test_run_experiments_driver(tmp_path, datasets, golden):
    """tools/run_experiments.py, the counterpart of run_experiments.sh:28-66: a `<dataset>:<model>` matrix, ms/graph table,
    per-graph outputs in the reference's `g%d: %.8f` format."""
    import sys
    packed = str(tmp_path / "molhiv200.fgb")
    datasets["molhiv"].slice(0, 200).save_packed(packed)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "run_experiments.py"), f"{packed}:gin", f"{packed}:pna", "synthetic-hep10k:64:ginvn",
                        "--trials", "3", "--out-dir", str(tmp_path)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr + r.stdout
    rows = [l for l in r.stdout.split("\n") if l.strip() and not l.startswith("dataset")]
    assert len(rows) == 3 and all(float(l.split()[-1]) > 0 for l in rows), r.stdout
    for model in ("gin", "pna"):
        path = os.path.join(str(tmp_path), f"molhiv200.fgb.{model}.B200_output.txt")
        got = np.array([float(x.split(": ")[1]) for x in open(path).read().split("\n")[:-1]], dtype=np.float32)
        assert_parity(got, golden["molhiv"][model][:200], what=f"run_experiments {model}")
