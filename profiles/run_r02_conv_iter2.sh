#!/bin/bash
# Round 2, f3: rolled MMA issue loop + batched fp32 halo loads: parity, role profile (fp32 + fp16), layer / generator timings.
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_gpu_conv_stack.py tests/test_gpu_generator.py -q -x 2>&1 | tail -3
NFE_NVCC_FLAGS="-DNFE_MC_PROFILE" python -m nerffaceediting_b200.build --force > /dev/null
python profiles/modconv_role_profile.py 256 256 256 1 fp32 8
python profiles/modconv_role_profile.py 256 256 256 1 fp16 8
python profiles/modconv_role_profile.py 128 128 512 1 fp16 8 | head -3
python profiles/modconv_role_profile.py 256 128 512 2 fp16 8 | head -3
python -m nerffaceediting_b200.build --force > /dev/null
timeout 600 python profiles/bench_conv.py --json gpurun_out/bench_conv_r02.json 2>&1 | tail -13
timeout 600 python bench.py --no-cpu-baseline --steps 10 --warmup 3 > gpurun_out/bench_gen.json 2> gpurun_out/bench_gen.err
python - <<PY
import json
d = json.load(open('gpurun_out/bench_gen.json'))
for k in ('full_generator', 'full_generator_fp16_backbone'):
    print(k, d['extras'][k].get('ms_per_step'), d['extras'][k].get('error'))
PY
