#!/bin/bash
# Round 2, f3: N = 256 tiles for the fp32 (split-operand) mode: parity + layer / generator timings.
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_gpu_conv_stack.py tests/test_gpu_generator.py -q -x 2>&1 | tail -4
timeout 600 python profiles/bench_conv.py --json gpurun_out/bench_conv_r02.json 2>&1 | grep "float32\|sr8xdc\|backbone"
timeout 600 python bench.py --no-cpu-baseline --steps 10 --warmup 3 > gpurun_out/bench_gen.json 2> gpurun_out/bench_gen.err
python - <<PY
import json
d = json.load(open('gpurun_out/bench_gen.json'))
for k in ('full_generator', 'full_generator_fp16_backbone'):
    print(k, d['extras'][k].get('ms_per_step'), d['extras'][k].get('error'))
PY
