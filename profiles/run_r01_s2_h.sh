python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for i in 1 2; do python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python profiles/summarize_bench.py | head -2; done
python bench.py --workload c3 --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python profiles/summarize_bench.py | head -2
python bench.py --workload c1 --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python profiles/summarize_bench.py | head -2
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --two-gather 2>/dev/null | python profiles/summarize_bench.py | head -2
