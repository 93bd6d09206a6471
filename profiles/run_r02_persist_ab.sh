#!/bin/bash
# Round 2, f3: persistent convolution CTAs: N <= 128 layers as persistent window pairs with two accumulator sets ($NFE_MC_PERSIST_N128=1) vs twin CTAs (=0).
# Parity first (short timeouts: a barrier bug would hang), then layer / SR head / backbone timings.
cd "$(dirname "$0")/.."
timeout 240 python -m pytest tests/test_gpu_conv_stack.py -q -x 2>&1 | tail -3
timeout 240 python -m pytest tests/test_gpu_generator.py tests/test_gpu_plugins.py -q -x 2>&1 | tail -2
for v in 0 1; do
  echo "=== NFE_MC_PERSIST_N128=$v"
  NFE_MC_PERSIST_N128=$v timeout 300 python profiles/bench_conv.py 2>/dev/null | cut -c1-110 | grep "128->128\|256->256 @ 256^2 up=1 float16\|sr8xdc\|backbone"
done
