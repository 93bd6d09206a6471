# Timing ablations of field_pipe_kernel (results are wrong by construction; only the stage times matter).  Run on the GPU box.
for flags in "" "-DNFE_ABLATE_SOFTPLUS" "-DNFE_ABLATE_GATHER" "-DNFE_ABLATE_SOFTPLUS -DNFE_ABLATE_GATHER"; do
  NFE_NVCC_FLAGS="$flags" python -m nerffaceediting_b200.build --force > /dev/null 2>&1
  echo "== flags: $flags"
  timeout -k 5 120 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python profiles/summarize_bench.py | sed -n 2p
done
