set -x
python -m pytest tests/test_gpu_run_model_bwd.py -m gpu -q --tb=short 2>&1 | grep -v "^  \|^E    +" | tail -60 > gpurun_out/gputest_r02_b.txt
tail -5 gpurun_out/gputest_r02_b.txt
# v2 kernel parity
python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py tests/test_gpu_backward.py -m gpu -q -x --tb=short 2>&1 | tail -30 > gpurun_out/gputest_r02_v2.txt
tail -5 gpurun_out/gputest_r02_v2.txt
for gen in 1 2; do
  NFE_FIELD_PIPE=$gen python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/ab_v2_gen${gen}.json 2> gpurun_out/ab_v2_gen${gen}.err
  python - <<PY
import json
d=json.load(open('gpurun_out/ab_v2_gen${gen}.json'))
print('gen${gen}', d['ms_per_step'], d['stages_ms_per_step'], d['roofline']['avg_launch_ms'])
PY
done
