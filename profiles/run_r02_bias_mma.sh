#!/bin/bash
# Round 2, field kernel: layer-1 bias + log2(e) through the tensor core (NFE_P2_BIAS_MMA) — parity, then A/B against the old epilogue.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py -x -q 2>&1 | tail -3
bash profiles/run_r02_pipe2_variants.sh "bm1|-DNFE_P2_BIAS_MMA=1" "bm0|-DNFE_P2_BIAS_MMA=0" "bm1b|-DNFE_P2_BIAS_MMA=1" 2>&1
python -m nerffaceediting_b200.build --force > /dev/null
