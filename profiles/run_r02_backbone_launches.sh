#!/bin/bash
cd "$(dirname "$0")/.."
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r02_backbone_fp32.csv python profiles/bench_conv.py --backbone-only 0 > /dev/null 2>&1
python profiles/launch_summary.py gpurun_out/launches_r02_backbone_fp32.csv "fp32 backbone, batch 8" | head -24
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r02_sr_head.csv python profiles/bench_conv.py --sr-only > /dev/null 2>&1
python profiles/launch_summary.py gpurun_out/launches_r02_sr_head.csv "fp16 SR head, batch 8" | head -16
