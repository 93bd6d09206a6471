#!/bin/bash
# Round 2, final state: whole GPU suite, smoke, the default bench line (with extras), the convolution-stack bench, launch lists of the SR head and
# of the fp32 backbone, generator latency, one ncu capture of conv_gemm_kernel on its largest layer.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
S=$(date +%s)
timeout 1500 python -m pytest tests -m gpu -q --tb=short 2>&1 | grep -v "^  \|^E    +" | tail -25 > gpurun_out/gputest_r02_full.txt
tail -6 gpurun_out/gputest_r02_full.txt
echo "tests t=$(( $(date +%s)-S ))s"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/bench_r02_c2_default.json 2> gpurun_out/bench_r02_c2_default.err
python - <<PY
import json
try:
    d = json.load(open('gpurun_out/bench_r02_c2_default.json'))
    print('c2', round(d['ms_per_step'], 4), 'ms', round(d['value'] / 1e6, 3), 'M rays/s; e2e', round(d['e2e']['value'] / 1e6, 3), 'stages', {k: round(v, 4) for k, v in d['stages_ms_per_step'].items()})
    print('roofline', d['roofline'])
    print('extras', json.dumps(d.get('extras'), indent=None)[:2500])
except Exception as e:
    print('bench FAILED', e, open('gpurun_out/bench_r02_c2_default.err').read()[-600:])
PY
echo "bench t=$(( $(date +%s)-S ))s"
timeout 600 python profiles/bench_conv.py --json gpurun_out/bench_conv_r02.json 2>&1 | tail -16
timeout 300 python profiles/bench_generator_latency.py --json gpurun_out/generator_latency_r02.json 2>&1 | tail -4
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r02_sr_head.csv python profiles/bench_conv.py --sr-only > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r02_backbone_fp32.csv python profiles/bench_conv.py --backbone-only 0 > /dev/null 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:conv_gemm -s 3 -c 1 -o gpurun_out/prof_conv_r02 -f python profiles/modconv_role_profile.py 256 256 256 1 fp16 8 2>&1 | tail -2
echo "total t=$(( $(date +%s)-S ))s"
