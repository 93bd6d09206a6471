#!/bin/bash
# Round 2, f3: role profile of conv_gemm_kernel (debug build) + one ncu capture of the production build on the hot layer.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
if [ "$1" != "ncu-only" ]; then
NFE_NVCC_FLAGS=-DNFE_MC_PROFILE python -m nerffaceediting_b200.build --force > /dev/null
python profiles/modconv_role_profile.py 256 256 256 1 fp16 8
python profiles/modconv_role_profile.py 128 128 512 1 fp16 8
python profiles/modconv_role_profile.py 256 128 512 2 fp16 8
python profiles/modconv_role_profile.py 512 512 64 1 fp16 8
python -m nerffaceediting_b200.build --force > /dev/null
fi
timeout 600 ncu --set full --import-source on --clock-control none -k regex:conv_gemm -s 3 -c 1 -o gpurun_out/prof_conv_r02 -f python profiles/modconv_role_profile.py 256 256 256 1 fp16 8 2>&1 | tail -3
