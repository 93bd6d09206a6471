#!/bin/bash
# Round 2, field kernel: instruction-cache footprint cut from 110 KB to 56 KB of steady-state code (rolled epilogue / tap loops).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
S=$(date +%s)
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3
echo "t=$(( $(date +%s)-S ))s"
bash profiles/run_r02_pipe2_variants.sh "cs_default|" "cs_mma|-DNFE_P2_ROLLED_MMA=1" "cs_h2|-DNFE_P2_HIDDEN_UNROLL=2" "cs_g112|-DNFE_P2_REGS_EPI=96 -DNFE_P2_REGS_GATHER=112 -DNFE_P2_REGS_MISC=64" 2>&1
echo "total t=$(( $(date +%s)-S ))s"
