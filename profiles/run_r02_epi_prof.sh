#!/bin/bash
# Round 2, f3: role profile (debug build; figures per WINDOW) of the persistent and the one-window-per-CTA kernels, production build restored afterwards.
cd "$(dirname "$0")/.."
NFE_NVCC_FLAGS="-DNFE_MC_PROFILE $EXTRA" python -m nerffaceediting_b200.build --force > /dev/null
for v in 0 1; do
echo "=== NFE_MC_PERSIST=$v"
NFE_MC_PERSIST=$v python profiles/modconv_role_profile.py 256 256 256 1 fp16 8
NFE_MC_PERSIST=$v python profiles/modconv_role_profile.py 256 128 512 2 fp16 8 | grep -v "fast loop"
NFE_MC_PERSIST=$v python profiles/modconv_role_profile.py 256 256 256 1 fp32 8 | grep -v "fast loop"
done
python -m nerffaceediting_b200.build --force > /dev/null
