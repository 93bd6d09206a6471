#!/bin/bash
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_gpu_advice_r01.py tests/test_gpu_backward.py tests/test_gpu_losses.py tests/test_gpu_run_model_bwd.py -m gpu -q --tb=short 2>&1 | grep -v "^  \|^E    +" | tail -40 > gpurun_out/gputest_r02_b.txt
tail -12 gpurun_out/gputest_r02_b.txt
for wl in c4 c1 c3 c5; do
  timeout 600 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r02_$wl.json 2> gpurun_out/bench_r02_$wl.err
  python - <<PY
import json
try:
    d = json.load(open('gpurun_out/bench_r02_$wl.json'))
    print('$wl', round(d['ms_per_step'], 4), 'ms', round(d['value'] / 1e6, 3), 'M rays/s', {k: round(v, 4) for k, v in d['stages_ms_per_step'].items()}, 'graph', (d.get('cuda_graph') or {}).get('ms_per_step'))
except Exception as e:
    print('$wl FAILED', e, open('gpurun_out/bench_r02_$wl.err').read()[-400:])
PY
done
# compositing kernel variant prepared in round 1 and never measured: all 32 lanes load record rows
CS=nerffaceediting_b200/csrc
unset CC CXX
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-fvisibility=hidden --expt-relaxed-constexpr -DNFE_MARCH_FULL_WARP=1 -c $CS/nfe_march.cu -o $CS/_obj/nfe_march.o 2>&1 | grep -E "error"
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o nerffaceediting_b200/lib/libnfe_b200.so $CS/_obj/*.o -cudart static
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/bench_r02_c2_march_fullwarp.json 2> /dev/null
python - <<PY
import json
d = json.load(open('gpurun_out/bench_r02_c2_march_fullwarp.json')); print('march full-warp', d['ms_per_step'], d['stages_ms_per_step'])
PY
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "render or march or composite" 2>&1 | tail -2
