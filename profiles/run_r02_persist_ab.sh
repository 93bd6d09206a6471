#!/bin/bash
# Round 2, f3: persistent convolution CTAs (conv_gemm_kernel<..., PS = true>) vs one window per CTA ($NFE_MC_PERSIST=0): parity first
# (short timeouts: a barrier bug would hang), then layer / SR head / backbone timings.
cd "$(dirname "$0")/.."
timeout 240 python -m pytest tests/test_gpu_conv_stack.py -q -x 2>&1 | tail -3
timeout 240 python -m pytest tests/test_gpu_generator.py tests/test_gpu_plugins.py -q -x 2>&1 | tail -2
for v in 0 1; do
  echo "=== NFE_MC_PERSIST=$v"
  NFE_MC_PERSIST=$v timeout 300 python profiles/bench_conv.py 2>/dev/null | cut -c1-110
done
