# Sweep the number of gather warps of field_pipe_kernel (compile-time), bench c2 for each.  Run on the GPU box.
for gw in ${GWS:-5 6 7 8}; do
  NFE_NVCC_FLAGS="-DNFE_GATHER_WARPS=$gw" python -m nerffaceediting_b200.build --force > /dev/null 2>&1
  echo "== gather_warps=$gw"
  timeout -k 5 120 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python profiles/summarize_bench.py | head -2
done
