"""CPU: the N>1 path (contiguous graph shards + tally/gather) with the gloo backend, world_size 2.
There is no GPU here, so each rank's shard is evaluated by the oracle port; what is under test is the
host logic in flowgnn_b200/sharding.py."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import GOLDEN, ROOT
from oracle import refbind

pytestmark = pytest.mark.skipif(not refbind.have_port(), reason="oracle port not built")


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from flowgnn_b200.dataset import load_npz
    from flowgnn_b200.sharding import gather_predictions, shard_of, tally
    from flowgnn_b200.weights import load_weights
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    batch = load_npz(os.path.join(GOLDEN, "molhiv.npz")).slice(0, 90)
    w = load_weights("pna", os.path.join(GOLDEN, "weights", "PNA"))
    shard, g0, g1 = shard_of(batch, rank, world)
    local = refbind.run_port("pna", shard, w)
    total, tmax = tally(shard.num_graphs, 10.0 * (rank + 1), dist)
    full = gather_predictions(local, g0, batch.num_graphs, dist)
    if rank == 0:
        np.savez(out_path, full=full, total=total, tmax=tmax, g1=g1)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_reassembles_the_single_rank_result(tmp_path):
    import torch.multiprocessing as mp
    from flowgnn_b200.dataset import load_npz
    from flowgnn_b200.weights import load_weights
    out = str(tmp_path / "r0.npz")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    z = np.load(out)
    batch = load_npz(os.path.join(GOLDEN, "molhiv.npz")).slice(0, 90)
    want = refbind.run_port("pna", batch, load_weights("pna", os.path.join(GOLDEN, "weights", "PNA")))
    assert int(z["total"]) == 90 and float(z["tmax"]) == 20.0 and 0 < int(z["g1"]) < 90
    assert np.array_equal(z["full"].view(np.int32), want.view(np.int32))


def test_single_process_paths_need_no_process_group():
    from flowgnn_b200.dataset import load_npz
    from flowgnn_b200.sharding import gather_predictions, shard_of, tally
    batch = load_npz(os.path.join(GOLDEN, "molhiv.npz")).slice(0, 10)
    s, g0, g1 = shard_of(batch, 0, 1)
    assert (g0, g1) == (0, 10) and s.num_graphs == 10
    assert tally(10, 3.5) == (10, 3.5)
    assert gather_predictions(np.arange(10, dtype=np.float32), 0, 10) is not None
