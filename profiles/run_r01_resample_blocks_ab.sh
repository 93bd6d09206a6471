# resample_kernel (coarse weights + inverse CDF) occupancy hint: 3 / 6 resident blocks per SM against the shipped 4 (0.0492 ms at c2 on every box)
mkdir -p gpurun_out
for MB in 3 6; do
  NFE_NVCC_FLAGS="-DNFE_RESAMPLE_MIN_BLOCKS=$MB" python -m nerffaceediting_b200.build --force > /dev/null 2>&1 || { echo "build failed"; continue; }
  timeout 40 python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/ab_resample_mb$MB.json 2>/dev/null
  python - gpurun_out/ab_resample_mb$MB.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); st = d["stages_ms_per_step"]
print("  %-28s step %.4f ms  resample %.4f  march_final %.4f" % (sys.argv[1].split('/')[-1], d["ms_per_step"], st["resample"], st["march_final"]))
PY
done
