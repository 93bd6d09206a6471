#!/bin/bash
# Round 2: launch list of the c2 step and ncu --set full captures of the production field kernel and the compositing kernel;
# then the default bench line with every extra.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
S=$(date +%s)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r02_c2.csv python bench.py --steps 2 --warmup 1 --no-extras --no-cpu-baseline > /dev/null 2>&1
echo "launch list t=$(( $(date +%s)-S ))s"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"field_pipe2_kernel|march_kernel" -s 4 -c 3 -o gpurun_out/prof_field_r02 -f python bench.py --steps 2 --warmup 1 --no-extras --no-cpu-baseline > /dev/null 2>&1
ls -la gpurun_out/prof_field_r02.ncu-rep
echo "ncu t=$(( $(date +%s)-S ))s"
timeout 900 python bench.py > gpurun_out/bench_r02_c2_default.json 2> gpurun_out/bench_r02_c2_default.err
python - <<PY
import json
try:
    d = json.load(open('gpurun_out/bench_r02_c2_default.json'))
    print('c2', round(d['ms_per_step'], 4), 'ms', round(d['value'] / 1e6, 3), 'M rays/s; e2e', round(d['e2e']['value'] / 1e6, 3))
    for k, v in (d.get('extras') or {}).items():
        print(' ', k, {kk: (round(vv, 3) if isinstance(vv, float) else vv) for kk, vv in v.items() if kk in ('value', 'ms_per_step', 'error')} if 'sr_head' not in v else v['sr_head']['ms'])
except Exception as e:
    print('bench FAILED', e, open('gpurun_out/bench_r02_c2_default.err').read()[-600:])
PY
echo "total t=$(( $(date +%s)-S ))s"
