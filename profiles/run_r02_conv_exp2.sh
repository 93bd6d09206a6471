#!/bin/bash
# timing experiments (wrong numerics): why does a split-operand (bf16) MMA take ~3x as long as an fp16 one?
cd "$(dirname "$0")/.."
for flags in "-DNFE_MC_EXP_TERMS1" "-DNFE_MC_EXP_FMT0" "-DNFE_MC_EXP_TERMS1 -DNFE_MC_EXP_FMT0"; do
  echo "=== flags: $flags"
  NFE_NVCC_FLAGS="-DNFE_MC_PROFILE $flags" python -m nerffaceediting_b200.build --force > /dev/null
  python profiles/modconv_role_profile.py 256 256 256 1 fp32 8 | grep -v "producer"
done
python -m nerffaceediting_b200.build --force > /dev/null
