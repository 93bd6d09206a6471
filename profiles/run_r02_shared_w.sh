#!/bin/bash
# Round 2, f3: split-operand (fp32) layers with the weights packed once for the batch, style on the activations, demodulation in the epilogue
# ($NFE_MC_SHARED_W=0: per-item packed weights): parity, then fp32 layer / SR head / backbone / generator timings.
cd "$(dirname "$0")/.."
timeout 300 python -m pytest tests/test_gpu_conv_stack.py tests/test_gpu_generator.py -q -x 2>&1 | tail -3
for v in 0 1; do
  echo "=== NFE_MC_SHARED_W=$v"
  NFE_MC_SHARED_W=$v timeout 300 python profiles/bench_conv.py 2>/dev/null | cut -c1-100 | grep "float32\|sr8xdc fp32"
  NFE_MC_SHARED_W=$v timeout 300 python profiles/bench_generator_latency.py 2>&1 | tail -1
done
