#!/bin/bash
# Round-2 check pass on the GPU box: full GPU test suite, smoke, default bench.  Every step under its own timeout.
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests -m gpu -q --tb=short 2>&1 | grep -v "^  \|^E    +" | tail -60 > gpurun_out/gputest_r02_full.txt
tail -15 gpurun_out/gputest_r02_full.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/bench_r02_c2.json 2> gpurun_out/bench_r02_c2.err
tail -c 600 gpurun_out/bench_r02_c2.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/bench_r02_c2.json'))
print({k: d[k] for k in ('value', 'ms_per_step', 'gpu_launches')})
print('e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e'].get('h2d_gbs_per_rank_all_ranks_uploading'))
r = d['roofline']; print('roofline', r['bound'], r['frac'], r['t_min_terms_ms'], r['avg_launch_ms'], r['l2_gbs_measured'])
print('stages', d['stages_ms_per_step'])
print('extras', {k: (v.get('ms_per_step'), v.get('error')) for k, v in d.get('extras', {}).items()})
print('cpu', d.get('cpu_baseline', {}).get('value'), 'graph', d.get('cuda_graph', {}).get('ms_per_step'))
PY
