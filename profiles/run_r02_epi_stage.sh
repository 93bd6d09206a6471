#!/bin/bash
# Round 2, field kernel: all outputs staged through shared memory (branch-free epilogue) on top of the tap input prefetch.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
S=$(date +%s)
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3
NFE_QUAD_ORDER=1 timeout 300 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3
echo "t=$(( $(date +%s)-S ))s"
bash profiles/run_r02_pipe2_variants.sh "es_b3|" "es_e112|-DNFE_P2_REGS_EPI=112 -DNFE_P2_REGS_GATHER=96 -DNFE_P2_REGS_MISC=64" 2>&1
echo "total t=$(( $(date +%s)-S ))s"
