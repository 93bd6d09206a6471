#!/bin/bash
# Round 2: field_bwd_kernel with the next tile's tap pre-pass computed behind the first GEMM: parity + c4.
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_gpu_backward.py tests/test_gpu_run_model_bwd.py tests/test_gpu_full_size.py tests/test_gpu_advice_r01.py -m gpu -q -x 2>&1 | tail -3
timeout 600 python bench.py --workload c4 --steps 20 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/bench_bwd_c4.json 2> gpurun_out/bench_bwd_c4.err
python - <<PY
import json
try:
    d = json.load(open('gpurun_out/bench_bwd_c4.json'))
    print('c4', round(d['ms_per_step'], 4), 'ms', round(d['value'] / 1e6, 3), 'M rays/s')
except Exception as e:
    print('c4 FAILED', e, open('gpurun_out/bench_bwd_c4.err').read()[-400:])
PY
