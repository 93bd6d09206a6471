#!/bin/bash
# Round 2, f3: warp-local copy-out of the staged epilogue rows (instead of per-thread bulk stores): parity, role profile, SR head and generator.
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_gpu_plugins.py tests/test_gpu_conv_stack.py tests/test_gpu_generator.py -q -x 2>&1 | tail -5
NFE_NVCC_FLAGS="-DNFE_MC_PROFILE" python -m nerffaceediting_b200.build --force > /dev/null
python profiles/modconv_role_profile.py 256 128 512 2 fp16 8 | grep -v "loader\|producer\|epilogue t0: final\|epilogue t0: in"
python profiles/modconv_role_profile.py 128 128 512 1 fp16 8 | grep -v "loader\|producer\|epilogue t0: final\|epilogue t0: in"
python profiles/modconv_role_profile.py 256 256 256 1 fp16 8 | grep -v "loader\|producer\|epilogue t0: final\|epilogue t0: in"
python profiles/modconv_role_profile.py 256 256 256 1 fp32 8 | grep -v "loader\|producer\|epilogue t0: final\|epilogue t0: in"
python -m nerffaceediting_b200.build --force > /dev/null
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r02_sr_head.csv python profiles/bench_conv.py --sr-only > /dev/null 2>&1
python profiles/launch_summary.py gpurun_out/launches_r02_sr_head.csv "fp16 SR head, batch 8" 2>/dev/null | head -9
timeout 600 python bench.py --no-cpu-baseline --steps 10 --warmup 3 > gpurun_out/bench_gen.json 2> gpurun_out/bench_gen.err
python - <<PY
import json
d = json.load(open('gpurun_out/bench_gen.json'))
for k in ('full_generator', 'full_generator_fp16_backbone'):
    print(k, d['extras'][k].get('ms_per_step'), d['extras'][k].get('error'))
print('sr', d['extras']['f3_conv_stack']['sr_head']['ms'])
print(json.dumps(d['extras']['f3_conv_stack'])[:1500])
PY
