# Last A/B of the round: ring groups of 4 rows (two cp.async per lane and group: half the wait/commit/loop overhead) against groups of 2 rows.
mkdir -p gpurun_out
S=$(date +%s)
B="python bench.py --steps 40 --warmup 5 --no-cpu-baseline"
show() { python - "$1" <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); st = d["stages_ms_per_step"]
print("  %-30s step %.4f ms  march_final %.4f  resample %.4f  graph %.4f" % (sys.argv[1].split('/')[-1], d["ms_per_step"], st["march_final"], st["resample"], d.get("cuda_graph", {}).get("ms_per_step", 0)))
PY
}
timeout 150 python -m pytest tests -m gpu -x -q -k "variant or cuda_graph" > gpurun_out/pytest_variants.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/pytest_variants.log)"
timeout 60 $B > gpurun_out/ab_pairs1_g8.json 2>>gpurun_out/ab_pairs.err; show gpurun_out/ab_pairs1_g8.json
for V in "2 4" "2 6"; do
  set -- $V
  NFE_NVCC_FLAGS="-DNFE_MARCH_GROUP_PAIRS=$1 -DNFE_MARCH_RING_GROUPS=$2" python -m nerffaceediting_b200.build --force > /dev/null 2>gpurun_out/build_pairs.err || { echo "build failed"; tail -3 gpurun_out/build_pairs.err; continue; }
  timeout 60 $B > gpurun_out/ab_pairs$1_g$2.json 2>>gpurun_out/ab_pairs.err; show gpurun_out/ab_pairs$1_g$2.json
  [ "$V" = "2 4" ] && { timeout 100 python -m pytest tests -m gpu -x -q -k "render_small or config2 or high_sample or ray_marchers" > gpurun_out/pytest_pairs.log 2>&1; echo "pytest (pairs=2) rc=$? $(tail -1 gpurun_out/pytest_pairs.log)"; }
done
tail -3 gpurun_out/ab_pairs.err
echo "total t=$(( $(date +%s)-S ))s"
