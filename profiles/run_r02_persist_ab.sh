#!/bin/bash
# Round 2, f3: split-operand (fp32) up = 2 layers: two-tap weight blocks + one window per CTA ($NFE_MC_MERGE_SPLIT=1) vs one-tap blocks + persistent CTAs (=0).
cd "$(dirname "$0")/.."
for v in 1 0; do
  echo "=== NFE_MC_MERGE_SPLIT=$v"
  NFE_MC_MERGE_SPLIT=$v timeout 300 python profiles/bench_conv.py 2>/dev/null | cut -c1-110 | grep "up=2 float32\|sr8xdc fp32\|backbone"
  NFE_MC_MERGE_SPLIT=$v timeout 300 python profiles/bench_conv.py --backbone-only 0 2>/dev/null | tail -1 | cut -c1-110
done
timeout 240 python -m pytest tests/test_gpu_conv_stack.py tests/test_gpu_generator.py -q -x 2>&1 | tail -2
