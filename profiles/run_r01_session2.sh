set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err
cat gpurun_out/bench_c2.json
python bench.py --workload c4 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err
cat gpurun_out/bench_c4.json
python bench.py --workload c3 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err
python bench.py --workload c1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_c1.json 2> gpurun_out/bench_c1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c2.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/b_c2.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_c4.csv python bench.py --workload c4 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/b_c4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"field_bwd|march_bwd" -s 4 -c 3 -o gpurun_out/prof_c4_bwd -f python bench.py --workload c4 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/p_c4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"field_pipe|march_kernel" -s 8 -c 4 -o gpurun_out/prof_c2 -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/p_c2.log 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
lscpu | head -20 > gpurun_out/lscpu.txt
