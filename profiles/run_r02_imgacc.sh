#!/bin/bash
# Round 2, f3: fused skip-image accumulation (nfe_image_accumulate): parity, then backbone / SR head / generator timings.
cd "$(dirname "$0")/.."
timeout 300 python -m pytest tests/test_gpu_plugins.py tests/test_gpu_conv_stack.py tests/test_gpu_generator.py -q -x 2>&1 | tail -2
timeout 300 python profiles/bench_conv.py 2>/dev/null | cut -c1-110 | grep "sr8xdc\|backbone"
timeout 300 python profiles/bench_conv.py --backbone-only 0 2>/dev/null | cut -c1-110 | tail -2
timeout 300 python profiles/bench_generator_latency.py 2>&1 | tail -2
