#!/bin/bash
# Round 2, f3: up = 2 weight blocks of two taps (one N = 256 MMA over two adjacent phase accumulators) vs one tap per block ($NFE_MC_MERGE=0).
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_gpu_plugins.py tests/test_gpu_conv_stack.py tests/test_gpu_generator.py -q -x 2>&1 | tail -3
NFE_NVCC_FLAGS="-DNFE_MC_PROFILE" python -m nerffaceediting_b200.build --force > /dev/null
for mg in 0 1; do
  echo "=== NFE_MC_MERGE=$mg"
  NFE_MC_MERGE=$mg python profiles/modconv_role_profile.py 256 128 512 2 fp16 8 | grep -v "loader\|producer\|fast loop"
  NFE_MC_MERGE=$mg python profiles/modconv_role_profile.py 512 512 64 2 fp32 8 | grep -v "loader\|producer\|fast loop"
done
python -m nerffaceediting_b200.build --force > /dev/null
for mg in 0 1; do
  echo "=== NFE_MC_MERGE=$mg"
  NFE_MC_MERGE=$mg python profiles/bench_conv.py 2>/dev/null | grep "up=2\|sr8xdc\|backbone"
done
