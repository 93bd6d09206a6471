# A/B: field_pipe_kernel epilogue sends the hidden activations to tensor memory in halves of 32 units (NFE_EPI_HALVES=1: 117 registers) vs all 64 at once (128)
mkdir -p gpurun_out
S=$(date +%s)
B="python bench.py --steps 40 --warmup 5 --no-cpu-baseline"
show() { python - "$1" <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); st = d["stages_ms_per_step"]
print("  %-30s step %.4f ms  field %.4f / %.4f  graph %.4f" % (sys.argv[1].split('/')[-1], d["ms_per_step"], st["field_coarse"], st["field_fine"], d.get("cuda_graph", {}).get("ms_per_step", 0)))
PY
}
timeout 100 $B > gpurun_out/ab_epi0_1.json 2>>gpurun_out/ab_epi.err; show gpurun_out/ab_epi0_1.json
NFE_NVCC_FLAGS="-DNFE_EPI_HALVES=1" python -m nerffaceediting_b200.build --force > /dev/null 2>gpurun_out/build_epi.err || { echo "build failed"; tail -3 gpurun_out/build_epi.err; }
timeout 100 $B > gpurun_out/ab_epi1_1.json 2>>gpurun_out/ab_epi.err; show gpurun_out/ab_epi1_1.json
timeout 100 $B --precision bf16 > gpurun_out/ab_epi1_bf16.json 2>>gpurun_out/ab_epi.err; show gpurun_out/ab_epi1_bf16.json
timeout 100 $B --workload c1 > gpurun_out/ab_epi1_c1.json 2>>gpurun_out/ab_epi.err; show gpurun_out/ab_epi1_c1.json
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_epi1.log 2>&1; echo "pytest (halves) rc=$? $(tail -1 gpurun_out/pytest_epi1.log)"
timeout 100 $B > gpurun_out/ab_epi1_2.json 2>>gpurun_out/ab_epi.err; show gpurun_out/ab_epi1_2.json
python -m nerffaceediting_b200.build --force > /dev/null
timeout 100 $B > gpurun_out/ab_epi0_2.json 2>>gpurun_out/ab_epi.err; show gpurun_out/ab_epi0_2.json
tail -3 gpurun_out/ab_epi.err
echo "total t=$(( $(date +%s)-S ))s"
