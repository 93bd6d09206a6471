#!/bin/bash
# Round 2, f3: timing experiment — does the 160-byte window-row pitch (core matrices straddling 128-byte lines) slow the MMAs down?
cd "$(dirname "$0")/.."
for flags in "" "-DNFE_MC_EXP_ALIGNED"; do
  echo "=== flags: $flags"
  NFE_NVCC_FLAGS="-DNFE_MC_PROFILE $flags" python -m nerffaceediting_b200.build --force > /dev/null
  python profiles/modconv_role_profile.py 256 256 256 1 fp16 8 | grep -v "loader\|producer"
  python profiles/modconv_role_profile.py 128 128 512 1 fp16 8 | grep -v "loader\|producer"
  python profiles/modconv_role_profile.py 256 128 512 2 fp16 8 | grep -v "loader\|producer"
done
python -m nerffaceediting_b200.build --force > /dev/null
