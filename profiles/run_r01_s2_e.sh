mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python __graft_entry__.py smoke 2>&1 | tail -2
for i in 1 2; do python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python profiles/summarize_bench.py; done
NFE_MARCH_NO_BULK=1 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python profiles/summarize_bench.py
python bench.py --workload c3 --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python profiles/summarize_bench.py
