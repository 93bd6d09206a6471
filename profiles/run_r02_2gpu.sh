#!/bin/bash
# two B200s of one box: the NCCL form of the sharding test, then the bench exactly as the driver launches it
cd "$(dirname "$0")/.."
nvidia-smi topo -m 2>&1 | head -12 > gpurun_out/topo_2gpu.txt
timeout 600 python -m pytest tests/test_gpu_sharding.py -m gpu -q --tb=short 2>&1 | tail -5
for wl in c2 c5; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --workload $wl > gpurun_out/bench_r02_${wl}_2gpu.json 2> gpurun_out/bench_r02_${wl}_2gpu.err
  python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/bench_r02_${wl}_2gpu.json').read().strip().splitlines()[-1])
    print('$wl x2', round(d['ms_per_step'], 4), 'ms', round(d['value'] / 1e6, 2), 'M rays/s; e2e', round(d['e2e']['value'] / 1e6, 2), d['e2e']['ms_per_step'], 'h2d GB/s/rank', d['e2e'].get('h2d_gbs_per_rank_all_ranks_uploading'), d['e2e'].get('numa'))
except Exception as e:
    print('$wl x2 FAILED', e, open('gpurun_out/bench_r02_${wl}_2gpu.err').read()[-800:])
PY
done
