#!/bin/bash
# eight B200s of one box: the bench exactly as the driver launches it (c2), plus topology
cd "$(dirname "$0")/.."
nvidia-smi topo -m 2>&1 | head -14 > gpurun_out/topo_8gpu.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench_r02_c2_8gpu.json 2> gpurun_out/bench_r02_c2_8gpu.err
python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/bench_r02_c2_8gpu.json').read().strip().splitlines()[-1])
    print('c2 x8', round(d['ms_per_step'], 4), 'ms', round(d['value'] / 1e6, 2), 'M rays/s; e2e', round(d['e2e']['value'] / 1e6, 2), round(d['e2e']['ms_per_step'], 3), 'h2d GB/s/rank', d['e2e'].get('h2d_gbs_per_rank_all_ranks_uploading'), d['e2e'].get('numa'))
except Exception as e:
    print('c2 x8 FAILED', e, open('gpurun_out/bench_r02_c2_8gpu.err').read()[-800:])
PY
