timeout -k 5 200 python -m pytest tests/test_gpu_backward.py -x -q 2>&1 | tail -2
for i in 1 2; do timeout -k 5 200 python bench.py --workload c4 --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python profiles/summarize_bench.py | head -1; done
