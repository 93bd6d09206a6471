#!/bin/bash
# Round 2, f3: first GPU run of the plugin and convolution-stack parity tests.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_plugins.py -q 2>&1 | tail -15
timeout 420 python -m pytest tests/test_gpu_conv_stack.py -q 2>&1 | tail -40
