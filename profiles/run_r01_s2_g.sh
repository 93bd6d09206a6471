for wl in c3 c2; do
  echo "== $wl plain"; python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python profiles/summarize_bench.py | head -2
  echo "== $wl quad";  NFE_QUAD_ORDER=1 python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python profiles/summarize_bench.py | head -2
done
