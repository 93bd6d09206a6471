#!/bin/bash
# Round 2: warp-uniform MMA issue (elect.sync + add-only descriptors) in field_pipe2_kernel and field_bwd_kernel: parity + c2 / c4.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
S=$(date +%s)
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_backward.py tests/test_gpu_run_model_bwd.py tests/test_gpu_full_size.py tests/test_gpu_advice_r01.py -m gpu -q -x 2>&1 | tail -3
echo "tests t=$(( $(date +%s)-S ))s"
for wl in c2 c4; do
  timeout 600 python bench.py --workload $wl --steps 20 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/bench_ui_$wl.json 2> gpurun_out/bench_ui_$wl.err
  python - <<PY
import json
try:
    d = json.load(open('gpurun_out/bench_ui_$wl.json'))
    print('$wl', round(d['ms_per_step'], 4), 'ms', round(d['value'] / 1e6, 3), 'M rays/s', {k: round(v, 4) for k, v in d['stages_ms_per_step'].items()})
except Exception as e:
    print('$wl FAILED', e, open('gpurun_out/bench_ui_$wl.err').read()[-400:])
PY
done
echo "total t=$(( $(date +%s)-S ))s"
