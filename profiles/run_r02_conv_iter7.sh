#!/bin/bash
# Round 2, f3: parity of the conv stack (with the odd-group shapes) and timings after the division-free copy-out.
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_gpu_plugins.py tests/test_gpu_conv_stack.py tests/test_gpu_generator.py -q -x 2>&1 | tail -3
python profiles/bench_conv.py 2>/dev/null | tail -25
