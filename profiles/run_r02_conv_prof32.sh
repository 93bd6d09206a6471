#!/bin/bash
cd "$(dirname "$0")/.."
NFE_NVCC_FLAGS="-DNFE_MC_PROFILE" python -m nerffaceediting_b200.build --force > /dev/null
python profiles/modconv_role_profile.py 256 256 256 1 fp32 8
python profiles/modconv_role_profile.py 512 512 64 1 fp32 8
python profiles/modconv_role_profile.py 256 256 256 1 fp16 8
python -m nerffaceediting_b200.build --force > /dev/null
