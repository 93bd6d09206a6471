# A/B of build flags on the GPU box: usage  bash profiles/run_ab.sh "<flagsA>" "<flagsB>" ...   (c2 bench, stage times)
for flags in "$@"; do
  NFE_NVCC_FLAGS="$flags" timeout 300 python -m nerffaceediting_b200.build --force > /dev/null 2>&1 || echo "BUILD FAILED: $flags"
  echo "== flags: [$flags]"
  for i in 1 2; do
    timeout -k 5 120 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | tail -1 | python profiles/summarize_bench.py | head -2
  done
done
