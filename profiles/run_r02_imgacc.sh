#!/bin/bash
# Round 2, f3: fused skip-image accumulation (nfe_image_accumulate): parity, then backbone / generator timings and the kernel's launch times.
cd "$(dirname "$0")/.."
timeout 300 python -m pytest tests/test_gpu_plugins.py tests/test_gpu_generator.py -q -x 2>&1 | tail -2
timeout 300 python profiles/bench_conv.py --backbone-only 0 2>/dev/null | cut -c1-110 | tail -1
timeout 300 python profiles/bench_generator_latency.py 2>&1 | tail -2
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_imgacc.csv -k regex:image_accumulate python profiles/bench_conv.py --backbone-only 0 > /dev/null 2>&1
python profiles/launch_summary.py gpurun_out/launches_imgacc.csv "image_accumulate" 2>/dev/null | head -4
