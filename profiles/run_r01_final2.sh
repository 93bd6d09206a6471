# Round-1 closing pass on the GPU box (gpurun): tests first, then the bench lines and launch lists that profiles/ cites.
# Every command has its own timeout so a slow one cannot eat the call; outputs land in gpurun_out/ as they finish.
mkdir -p gpurun_out
S=$(date +%s)
timeout 420 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/pytest_gpu.log) t=$(( $(date +%s)-S ))s"
timeout 90 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$? $(tail -1 gpurun_out/smoke.log)"
timeout 180 python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; python profiles/summarize_bench.py < gpurun_out/bench_c2.json
timeout 120 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 120 python bench.py --workload c4 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; python profiles/summarize_bench.py < gpurun_out/bench_c4.json
timeout 240 python bench.py --workload c5 --steps 3 --warmup 3 > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; python profiles/summarize_bench.py < gpurun_out/bench_c5.json; tail -3 gpurun_out/bench_c5.err
echo "t=$(( $(date +%s)-S ))s"
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c2.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/b_c2.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"field_pipe|march_kernel|resample" -s 12 -c 4 -o gpurun_out/prof_c2 -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/p_c2.log 2>&1
echo "t=$(( $(date +%s)-S ))s"
timeout 120 python bench.py --workload c3 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err
timeout 60 python bench.py --workload c1 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_c1.json 2> gpurun_out/bench_c1.err
for v in "--precision bf16:bf16" "--precision fp32:fp32" "--two-gather:two_gather"; do
  timeout 60 python bench.py ${v%%:*} --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_c2_${v##*:}.json 2>/dev/null
done
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_c4.csv python bench.py --workload c4 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/b_c4.log 2>&1
echo "total t=$(( $(date +%s)-S ))s"
