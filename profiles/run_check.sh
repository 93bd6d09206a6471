# quick GPU check with strict timeouts: smoke, c2 bench x2, full GPU tests, c3 / two-gather / bf16 lines
timeout -k 5 90 python __graft_entry__.py smoke 2>&1 | tail -1 || echo "SMOKE TIMED OUT / FAILED"
for i in 1 2; do timeout -k 5 90 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | tail -1 | python profiles/summarize_bench.py | sed -n 1,2p; done
timeout -k 5 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout -k 5 90 python bench.py --workload c3 --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python profiles/summarize_bench.py | sed -n 1,2p
timeout -k 5 90 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --two-gather 2>/dev/null | tail -1 | python profiles/summarize_bench.py | sed -n 1,2p
timeout -k 5 90 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --precision bf16 2>/dev/null | tail -1 | python profiles/summarize_bench.py | sed -n 1,2p
