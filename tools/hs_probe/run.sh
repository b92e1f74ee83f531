#!/bin/bash
# usage (GPU box): bash tools/hs_probe/run.sh -- host-side narrowing throughput by thread count, host memory bandwidth, pinned H2D rate
cd tools/hs_probe
g++ -O3 -std=c++17 -pthread narrow_probe.cc ../../flowgnn_b200/csrc/host_stage.cc -o /tmp/narrow_probe && g++ -O2 -pthread bw_probe.cc -o /tmp/bw_probe
nproc; lscpu | grep -E "Model name|^CPU\(s\)|Thread|Socket"
for t in 1 2 4 8 12 16; do /tmp/narrow_probe $t; FLOWGNN_B200_NO_AVX2=1 /tmp/narrow_probe $t; done
for t in 1 4 8 16; do /tmp/bw_probe $t; done
python - <<'PY'
import torch, time
n = 82_600_000
h = torch.empty(n, dtype=torch.uint8).pin_memory(); d = torch.empty(n, dtype=torch.uint8, device="cuda")
for sz in (n, n // 3, n // 10):
    for _ in range(3): d[:sz].copy_(h[:sz], non_blocking=True)
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(10): d[:sz].copy_(h[:sz], non_blocking=True)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 10
    print(f"pinned H2D {sz/1e6:.1f} MB: {dt*1e3:.3f} ms = {sz/dt/1e9:.1f} GB/s")
PY
