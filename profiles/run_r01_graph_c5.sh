# Second closing pass: the CUDA-graph path (test + c1/c2 bench lines) and the c5 (256^2, 96+96, one identity) variants.
mkdir -p gpurun_out
S=$(date +%s)
timeout 200 python -m pytest tests -m gpu -x -q -k "cuda_graph or config5 or broadcast" > gpurun_out/pytest_graph.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/pytest_graph.log)"
for w in c1 c2; do
  timeout 100 python bench.py --workload $w --steps 50 --warmup 5 --no-cpu-baseline --cuda-graph > gpurun_out/bench_${w}_graph.json 2> gpurun_out/bench_${w}_graph.err; echo "$w graph rc=$?"; python profiles/summarize_bench.py < gpurun_out/bench_${w}_graph.json; tail -2 gpurun_out/bench_${w}_graph.err
done
timeout 60 python bench.py --workload c1 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_c1.json 2> gpurun_out/bench_c1.err; python profiles/summarize_bench.py < gpurun_out/bench_c1.json
echo "t=$(( $(date +%s)-S ))s"
timeout 150 python bench.py --workload c5 --steps 3 --warmup 3 > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; echo "c5 default"; python profiles/summarize_bench.py < gpurun_out/bench_c5.json
NFE_QUAD_ORDER=1 timeout 100 python bench.py --workload c5 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c5_quad.json 2> gpurun_out/bench_c5_quad.err; echo "c5 quad order"; python profiles/summarize_bench.py < gpurun_out/bench_c5_quad.json
NFE_WORKSPACE_MB=65536 timeout 100 python bench.py --workload c5 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c5_ws64g.json 2> gpurun_out/bench_c5_ws64g.err; echo "c5 workspace 64 GB"; python profiles/summarize_bench.py < gpurun_out/bench_c5_ws64g.json; tail -2 gpurun_out/bench_c5_ws64g.err
NFE_WORKSPACE_MB=2048 timeout 100 python bench.py --workload c5 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c5_ws2g.json 2> gpurun_out/bench_c5_ws2g.err; echo "c5 workspace 2 GB"; python profiles/summarize_bench.py < gpurun_out/bench_c5_ws2g.json
echo "total t=$(( $(date +%s)-S ))s"
