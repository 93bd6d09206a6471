# Last call of the round: the full GPU suite and the default bench line on the final build (ring groups of 4 rows).
mkdir -p gpurun_out
timeout 150 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/pytest_gpu.log)"
timeout 60 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 90 python bench.py --no-cpu-baseline > gpurun_out/bench_c2_last.json 2> gpurun_out/bench_c2_last.err; python profiles/summarize_bench.py < gpurun_out/bench_c2_last.json
