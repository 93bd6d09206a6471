#!/bin/bash
# A/B of field_pipe2_kernel build variants on the GPU box: rebuilds only nfe_field_pipe2.cu with the given flags, relinks, runs the
# role profile (debug build) and the c2 bench.  usage: bash profiles/run_r02_pipe2_variants.sh "name|flags" ...
cd "$(dirname "$0")/.."
CS=nerffaceediting_b200/csrc
NV="nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-fvisibility=hidden --expt-relaxed-constexpr"
relink() { nvcc -shared -gencode arch=compute_100a,code=sm_100a -o nerffaceediting_b200/lib/libnfe_b200.so $CS/_obj/*.o -cudart static; }
unset CC CXX
for spec in "$@"; do
  name="${spec%%|*}"; flags="${spec#*|}"
  echo "=== $name: $flags"
  $NV $flags -DNFE_PIPE_PROFILE -c $CS/nfe_field_pipe2.cu -o $CS/_obj/nfe_field_pipe2.o 2>&1 | grep -E "error|spill" ; relink
  timeout 120 python profiles/pipe2_role_profile.py 2>&1 | tail -4
  $NV $flags -c $CS/nfe_field_pipe2.cu -o $CS/_obj/nfe_field_pipe2.o 2>&1 | grep -E "error" ; relink
  timeout 180 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/p2_${name}.json 2> gpurun_out/p2_${name}.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/p2_${name}.json'))
    print('${name}: step %.4f ms  field %.4f ms/launch  stages %s' % (d['ms_per_step'], d['roofline']['avg_launch_ms'], {k: round(v, 4) for k, v in d['stages_ms_per_step'].items()}))
except Exception as e:
    print('${name}: FAILED', e, open('gpurun_out/p2_${name}.err').read()[-600:])
PY
done
