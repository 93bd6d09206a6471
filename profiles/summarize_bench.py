import sys, json
for line in sys.stdin:
    line = line.strip()
    if not line.startswith("{"): 
        continue
    d = json.loads(line)
    print("value %.3g rays/s  ms/step %.3f  e2e_ms %.3f  launches %s" % (d["value"], d["ms_per_step"], d["e2e"].get("ms_per_step", 0), d.get("gpu_launches")))
    print("  stages", {k: round(v, 3) for k, v in d.get("stages_ms_per_step", {}).items()})
    r = d.get("roofline", {})
    print("  roofline achieved %.0f GB/s frac %.2f mlp_tflops %.1f share %.2f" % (r.get("achieved", 0), r.get("frac", 0), r.get("mlp_tflops", 0), r.get("share_of_step", 0)), d.get("clocks"))
