mkdir -p gpurun_out
python -m pytest tests/test_gpu_backward.py -x -q 2>&1 | tail -15
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --workload c4 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; tail -3 gpurun_out/bench_c4.err
python profiles/summarize_bench.py < gpurun_out/bench_c4.json
python bench.py --workload c4 --steps 5 --warmup 3 --no-cpu-baseline --two-gather 2>/dev/null | python profiles/summarize_bench.py
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_c4.csv python bench.py --workload c4 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/b_c4.log 2>&1
python profiles/summarize_launches.py gpurun_out/launches_c4.csv | head -24
