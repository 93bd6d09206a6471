#!/bin/bash
# Round 2, f3: the convolution's CTA-shape switches, one at a time against the default build (profiles/modconv_tuning_r02.txt #14-16):
#   NFE_MC_PERSIST=0        one window per CTA everywhere (no persistent CTAs)
#   NFE_MC_PERSIST_N128=0   large fp16 layers of N <= 128 as twin CTAs instead of persistent window pairs
#   NFE_MC_MERGE=0          up = 2: one tap per weight block (nine N = 128 MMAs per K step instead of six)
#   NFE_MC_MERGE_SPLIT=0    ... for split (fp32) operands only
# Parity first (short timeouts: a barrier bug would hang), then layer / SR head / backbone timings.
cd "$(dirname "$0")/.."
timeout 240 python -m pytest tests/test_gpu_conv_stack.py -q -x 2>&1 | tail -2
for sw in "" NFE_MC_PERSIST=0 NFE_MC_PERSIST_N128=0 NFE_MC_MERGE=0 NFE_MC_MERGE_SPLIT=0; do
  echo "=== ${sw:-default}"
  env $sw timeout 300 python profiles/bench_conv.py 2>/dev/null | cut -c1-110 | grep "256->256 @ 256^2 up=1\|128->128\|up=2\|sr8xdc\|backbone"
done
