#!/bin/bash
# Round 2, final build: the bench's other modes (precisions, the reference's two-gather formulation, the configs[2] appearance swap at its
# own size, configs[0] as a CUDA graph).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { name=$1; shift; timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras "$@" > gpurun_out/bench_r02_$name.json 2> gpurun_out/bench_r02_$name.err
  python - <<PY
import json
try:
    d = json.load(open('gpurun_out/bench_r02_$name.json'))
    print('$name', round(d['ms_per_step'], 4), 'ms', round(d['value'] / 1e6, 3), 'M rays/s', 'graph', (d.get('cuda_graph') or {}).get('ms_per_step'))
except Exception as e:
    print('$name FAILED', e, open('gpurun_out/bench_r02_$name.err').read()[-300:])
PY
}
run c2_bf16 --precision bf16
run c2_fp32 --precision fp32
run c2_two_gather --two-gather
run c2_statistics_swap --swap-statistics
run c3_statistics_swap --workload c3 --swap-statistics
run c1 --workload c1
