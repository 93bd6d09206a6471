# A/B of the cp.async record ring in march_kernel<true> (merge + composite) against the register-staged loads, plus the c2 precision variants.
mkdir -p gpurun_out
S=$(date +%s)
B="python bench.py --steps 30 --warmup 5 --no-cpu-baseline"
show() { python - "$1" <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); st = d["stages_ms_per_step"]
print("  %-28s step %.3f ms  march_final %.4f  march_coarse %.4f  field %.3f/%.3f  graph %s" % (sys.argv[1].split('/')[-1], d["ms_per_step"], st["march_final"], st.get("march_coarse", 0),
      st["field_coarse"], st["field_fine"], d.get("cuda_graph", {}).get("ms_per_step")))
PY
}
NFE_MARCH_RING=1 timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_ring.log 2>&1; echo "pytest (ring on) rc=$? $(tail -1 gpurun_out/pytest_ring.log)"
timeout 100 $B > gpurun_out/ab_c2_ring0.json 2>gpurun_out/ab.err; show gpurun_out/ab_c2_ring0.json
NFE_MARCH_RING=1 timeout 100 $B > gpurun_out/ab_c2_ring8.json 2>>gpurun_out/ab.err; show gpurun_out/ab_c2_ring8.json
NFE_MARCH_RING=1 timeout 100 $B --workload c3 --steps 5 > gpurun_out/ab_c3_ring8.json 2>>gpurun_out/ab.err; show gpurun_out/ab_c3_ring8.json
NFE_MARCH_RING=1 timeout 100 $B --workload c5 --steps 3 --warmup 3 > gpurun_out/ab_c5_ring8.json 2>>gpurun_out/ab.err; show gpurun_out/ab_c5_ring8.json
echo "t=$(( $(date +%s)-S ))s"
for G in 12 16 4; do
  NFE_NVCC_FLAGS="-DNFE_MARCH_RING_GROUPS=$G" python -m nerffaceediting_b200.build --force > /dev/null 2>gpurun_out/build_$G.err || { echo "build $G failed"; tail -3 gpurun_out/build_$G.err; continue; }
  NFE_MARCH_RING=1 timeout 100 $B > gpurun_out/ab_c2_ring$G.json 2>>gpurun_out/ab.err; show gpurun_out/ab_c2_ring$G.json
done
echo "t=$(( $(date +%s)-S ))s"
python -m nerffaceediting_b200.build --force > /dev/null
timeout 100 $B --precision bf16 > gpurun_out/bench_c2_bf16.json 2>>gpurun_out/ab.err; show gpurun_out/bench_c2_bf16.json
timeout 100 $B --precision fp32 > gpurun_out/bench_c2_fp32.json 2>>gpurun_out/ab.err; show gpurun_out/bench_c2_fp32.json
timeout 100 $B --two-gather > gpurun_out/bench_c2_two_gather.json 2>>gpurun_out/ab.err; show gpurun_out/bench_c2_two_gather.json
tail -5 gpurun_out/ab.err
echo "total t=$(( $(date +%s)-S ))s"
