#!/bin/bash
# Round 2: the whole generator (BASELINE configs[1] in full) — parity against the reference's CPU output, then the bench line with extras.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_generator.py -q 2>&1 | tail -15
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_gen.json 2> gpurun_out/bench_gen.err
python - <<PY
import json
try:
    d = json.load(open('gpurun_out/bench_gen.json'))
    print('c2', round(d['ms_per_step'], 4), 'ms')
    for k, v in (d.get('extras') or {}).items():
        print(' ', k, {kk: (round(vv, 3) if isinstance(vv, float) else vv) for kk, vv in v.items() if kk in ('value', 'ms_per_step', 'images_per_s', 'error')})
except Exception as e:
    print('bench FAILED', e, open('gpurun_out/bench_gen.err').read()[-600:])
PY
