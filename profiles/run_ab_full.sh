# A/B of build flags on the whole c2 step (ms/step), several repetitions
for flags in "$@"; do
  NFE_NVCC_FLAGS="$flags" timeout 300 python -m nerffaceediting_b200.build --force > /dev/null 2>&1 || echo "BUILD FAILED: $flags"
  echo "== flags: [$flags]"
  for i in 1 2 3; do timeout -k 5 120 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | tail -1 | python profiles/summarize_bench.py | head -1; done
done
