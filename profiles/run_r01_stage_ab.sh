# A/B of stage32_kernel<true> (normalise + stage): descending walk after plane_stats (L2 reuse) and streaming stores for the NCHW copy.
mkdir -p gpurun_out
S=$(date +%s)
B="python bench.py --steps 40 --warmup 5 --no-cpu-baseline"
show() { python - "$1" <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); st = d["stages_ms_per_step"]
rest = d["ms_per_step"] - sum(st.values())
print("  %-30s step %.4f ms  outside the render stages (rays+stats+stage32+gaps) %.4f ms  graph %.4f" % (sys.argv[1].split('/')[-1], d["ms_per_step"], rest, d.get("cuda_graph", {}).get("ms_per_step", 0)))
PY
}
for V in "0 0" "1 0" "0 1" "1 1" "0 0" "1 1"; do
  set -- $V
  NFE_NVCC_FLAGS="-DNFE_STAGE_REVERSE=$1 -DNFE_STAGE_STREAM=$2" python -m nerffaceediting_b200.build --force > /dev/null 2>gpurun_out/build_stage.err || { echo "build failed"; tail -3 gpurun_out/build_stage.err; continue; }
  f=gpurun_out/ab_stage_r$1_s$2_$(date +%s).json
  timeout 100 $B > $f 2>>gpurun_out/ab_stage.err; show $f
done
echo "t=$(( $(date +%s)-S ))s"
NFE_NVCC_FLAGS="-DNFE_STAGE_REVERSE=1 -DNFE_STAGE_STREAM=1" python -m nerffaceediting_b200.build --force > /dev/null
timeout 100 $B --workload c3 --steps 5 > gpurun_out/ab_stage_c3_r1_s1.json 2>>gpurun_out/ab_stage.err; show gpurun_out/ab_stage_c3_r1_s1.json
timeout 200 python -m pytest tests -m gpu -x -q -k "plane or stage or normal or render_small or single_gather" > gpurun_out/pytest_stage.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/pytest_stage.log)"
tail -3 gpurun_out/ab_stage.err
echo "total t=$(( $(date +%s)-S ))s"
