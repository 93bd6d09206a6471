#!/bin/bash
# Round 2, field kernel: gather inner loop with per-plane accumulation of both feature sets, item selector precomputed by the tap warp.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
S=$(date +%s)
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_advice_r01.py -x -q 2>&1 | tail -3
echo "t=$(( $(date +%s)-S ))s"
bash profiles/run_r02_pipe2_variants.sh "g2_default|" "g2_b4|-DNFE_P2_TAP_BUFS=4" "g2_cg0|-DNFE_TAP_CG=0" "g2_cg2|-DNFE_TAP_CG=2" 2>&1
echo "total t=$(( $(date +%s)-S ))s"
