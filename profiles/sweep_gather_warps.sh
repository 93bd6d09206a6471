for cfg in "8 1" "11 1" "11 0" "10 1" "6 1"; do
  set -- $cfg
  NFE_NVCC_FLAGS="-DNFE_GATHER_WARPS=$1 -DNFE_PASS_CONTIG=$2" python -m nerffaceediting_b200.build --force > /dev/null 2>&1
  echo "== gather_warps=$1 contig=$2"
  timeout 120 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --precision bf16x3 2>&1 | tail -1 | python profiles/summarize_bench.py | head -2
done
