#!/bin/bash
# Round 2, f3: role profile of the up = 2 GEMM at the two super-resolution shapes (debug build), production build restored afterwards.
cd "$(dirname "$0")/.."
NFE_NVCC_FLAGS="-DNFE_MC_PROFILE" python -m nerffaceediting_b200.build --force > /dev/null
python profiles/modconv_role_profile.py 256 128 512 2 fp16 8
python profiles/modconv_role_profile.py 32 256 256 2 fp16 8
python profiles/modconv_role_profile.py 128 128 512 1 fp16 8
python -m nerffaceediting_b200.build --force > /dev/null
