#!/bin/bash
# Round 2, f3: raw epilogue for the up = 2 intermediate, exit wait on the bulk stores' reads only, cp.async tile load + packed FMAs in the filter pass
# (3 vs 4 resident blocks per SM): parity, role profile of the up = 2 GEMM, SR head launch list.
cd "$(dirname "$0")/.."
NFE_NVCC_FLAGS="-DNFE_MC_PROFILE" python -m nerffaceediting_b200.build --force > /dev/null
python profiles/modconv_role_profile.py 256 128 512 2 fp16 8 | grep -v "loader\|producer"
python profiles/modconv_role_profile.py 128 128 512 1 fp16 8 | grep -v "loader\|producer"
for mb in 3 4; do
  echo "=== filter pass, min blocks $mb"
  NFE_NVCC_FLAGS="-DNFE_FIN_MIN_BLOCKS=$mb" python -m nerffaceediting_b200.build --force > /dev/null
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r02_sr_head_mb$mb.csv python profiles/bench_conv.py --sr-only > /dev/null 2>&1
  python profiles/launch_summary.py gpurun_out/launches_r02_sr_head_mb$mb.csv "fp16 SR head, batch 8" 2>/dev/null | head -7
done
python -m nerffaceediting_b200.build --force > /dev/null
timeout 600 python -m pytest tests/test_gpu_plugins.py tests/test_gpu_conv_stack.py tests/test_gpu_generator.py -q -x 2>&1 | tail -3
