mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --workload c4 --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python profiles/summarize_bench.py
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_c4.csv python bench.py --workload c4 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/b_c4.log 2>&1
python profiles/summarize_launches.py gpurun_out/launches_c4.csv > gpurun_out/launches_c4_summary.txt; head -8 gpurun_out/launches_c4_summary.txt
for prec in bf16 fp32; do echo "== precision $prec"; python bench.py --steps 20 --warmup 5 --no-cpu-baseline --precision $prec 2>/dev/null | python profiles/summarize_bench.py; done
echo "== two-gather"; python bench.py --steps 20 --warmup 5 --no-cpu-baseline --two-gather 2>/dev/null | python profiles/summarize_bench.py
