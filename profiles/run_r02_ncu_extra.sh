#!/bin/bash
# Round 2: ncu --set full captures of the backward kernels (c4 step) and of the up = 2 convolution path (GEMM with four phase accumulators +
# the tiled filter pass), for profiles/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"field_bwd_kernel|march_bwd_kernel" -s 4 -c 3 -o gpurun_out/prof_bwd_r02 -f python bench.py --workload c4 --steps 2 --warmup 1 --no-extras --no-cpu-baseline > /dev/null 2>&1
ls -la gpurun_out/prof_bwd_r02.ncu-rep
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"conv_gemm_kernel|upfir_finish_tiled" -s 4 -c 2 -o gpurun_out/prof_conv_up2_r02 -f python profiles/modconv_role_profile.py 256 128 512 2 fp16 8 > /dev/null 2>&1
ls -la gpurun_out/prof_conv_up2_r02.ncu-rep
