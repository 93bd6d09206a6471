# A/B: coarse compositing weights formed inside resample_kernel (one launch) against march_kernel<false> + resample_kernel (NFE_SPLIT_COARSE=1)
mkdir -p gpurun_out
S=$(date +%s)
B="python bench.py --steps 40 --warmup 5 --no-cpu-baseline"
show() { python - "$1" <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); st = d["stages_ms_per_step"]
print("  %-30s step %.4f ms  march_coarse %.4f resample %.4f march_final %.4f  graph %.4f  launches/step %d" % (sys.argv[1].split('/')[-1], d["ms_per_step"],
      st.get("march_coarse", 0), st["resample"], st["march_final"], d.get("cuda_graph", {}).get("ms_per_step", 0), d["gpu_launches"] // d["steps"]))
PY
}
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_fused_coarse.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/pytest_fused_coarse.log)"
NFE_SPLIT_COARSE=1 timeout 300 python -m pytest tests -m gpu -x -q -k "render or resample or stage_taps" > gpurun_out/pytest_split.log 2>&1; echo "pytest (split) rc=$? $(tail -1 gpurun_out/pytest_split.log)"
for i in 1 2; do
  NFE_SPLIT_COARSE=1 timeout 100 $B > gpurun_out/ab_split_$i.json 2>>gpurun_out/ab_fc.err; show gpurun_out/ab_split_$i.json
  timeout 100 $B > gpurun_out/ab_fused_$i.json 2>>gpurun_out/ab_fc.err; show gpurun_out/ab_fused_$i.json
done
timeout 100 $B --workload c1 --steps 50 > gpurun_out/ab_fused_c1.json 2>>gpurun_out/ab_fc.err; show gpurun_out/ab_fused_c1.json
timeout 100 $B --workload c3 --steps 5 > gpurun_out/ab_fused_c3.json 2>>gpurun_out/ab_fc.err; show gpurun_out/ab_fused_c3.json
tail -3 gpurun_out/ab_fc.err
echo "total t=$(( $(date +%s)-S ))s"
