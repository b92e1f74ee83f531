#!/bin/bash
# usage (GPU box with N GPUs): bash tools/e2e_scale_probe.sh N  -- end-to-end arm of bench.py at N GPUs for the upload modes of the entry points
N=${1:-8}
run() {
  echo "== $*"
  env "$@" timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 \
      --no-extras --no-cpu-baseline --no-pageable 2>/dev/null | python -c "
import json, sys
for l in sys.stdin.read().strip().splitlines():
    if l.startswith('{'):
        d = json.loads(l)
        print('device %.1f M, e2e %.1f M graphs/s (%.3f ms/step, %d B h2d), packed %s' % (d['value'] / 1e6, d['e2e']['value'] / 1e6, d['e2e']['ms_per_step'], d['e2e']['h2d_bytes_per_step'], ('%.1f M (%.3f ms)' % (d['e2e_packed']['value'] / 1e6, d['e2e_packed']['ms_per_step'])) if d.get('e2e_packed') else None))
"
}
nproc
run FLOWGNN_B200_HOST_STAGE=0
run FLOWGNN_B200_HOST_STAGE=5
run FLOWGNN_B200_HOST_STAGE=7
