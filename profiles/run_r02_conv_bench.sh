#!/bin/bash
# Round 2, f3: parity tests again (tolerance fixes), then the convolution-stack bench and one ncu capture of the hot layer.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_plugins.py tests/test_gpu_conv_stack.py -q 2>&1 | tail -8
timeout 600 python profiles/bench_conv.py --json gpurun_out/bench_conv_r02.json 2>&1 | tail -30
