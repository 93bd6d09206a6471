# Round-1 final measurement pass (run on the GPU box through gpurun): tests, smoke, benches, launch lists, ncu captures.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -2 gpurun_out/pytest_gpu.log
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; python profiles/summarize_bench.py < gpurun_out/bench_c2.json
python bench.py --workload c4 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; python profiles/summarize_bench.py < gpurun_out/bench_c4.json
python bench.py --workload c3 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err
python bench.py --workload c1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_c1.json 2> gpurun_out/bench_c1.err
python bench.py --precision bf16 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_c2_bf16.json 2>/dev/null
python bench.py --precision fp32 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_c2_fp32.json 2>/dev/null
python bench.py --two-gather --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_c2_two_gather.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c2.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/b_c2.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_c4.csv python bench.py --workload c4 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/b_c4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"field_bwd|march_bwd|normalize_bwd" -s 5 -c 5 -o gpurun_out/prof_c4_bwd -f python bench.py --workload c4 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/p_c4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"field_pipe|march_kernel|stage32|plane_stats|resample" -s 14 -c 7 -o gpurun_out/prof_c2 -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/p_c2.log 2>&1
python profiles/bench_point_queries.py > gpurun_out/point_queries.txt 2>&1; tail -4 gpurun_out/point_queries.txt
python profiles/bench_resize.py > gpurun_out/resize.txt 2>&1
lscpu | head -20 > gpurun_out/lscpu.txt
