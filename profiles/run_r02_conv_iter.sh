#!/bin/bash
# Round 2, f3: one tuning iteration of conv_gemm_kernel: parity tests, role profile (debug build), layer bench (production build).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_conv_stack.py -q -x 2>&1 | tail -4
NFE_NVCC_FLAGS="-DNFE_MC_PROFILE $NFE_EXTRA" python -m nerffaceediting_b200.build --force > /dev/null
python profiles/modconv_role_profile.py 256 256 256 1 fp16 8
python profiles/modconv_role_profile.py 128 128 512 1 fp16 8
python profiles/modconv_role_profile.py 256 128 512 2 fp16 8
NFE_NVCC_FLAGS="$NFE_EXTRA" python -m nerffaceediting_b200.build --force > /dev/null
timeout 600 python profiles/bench_conv.py --json gpurun_out/bench_conv_r02.json 2>&1 | tail -30
