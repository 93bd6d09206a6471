#!/bin/bash
# Round 2, field kernel: tap warps fetch their inputs one chunk ahead + records always staged through shared memory.
# Parity first (plain and quad order), then 3 vs 4 tap buffers, then an ncu capture (with source) of the default build.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
S=$(date +%s)
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3
NFE_QUAD_ORDER=1 timeout 300 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3
echo "t=$(( $(date +%s)-S ))s"
bash profiles/run_r02_pipe2_variants.sh "pf_b4|-DNFE_P2_TAP_BUFS=4" "pf_b3|" 2>&1
echo "t=$(( $(date +%s)-S ))s"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"field_pipe2" -s 6 -c 2 -o gpurun_out/prof_p2_c -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/p_p2_c.log 2>&1
echo "total t=$(( $(date +%s)-S ))s"
