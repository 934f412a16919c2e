import json, sys
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print("value %.1f Msamples/s  ms/step %.2f  e2e %.1f  launches %d clocks %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"], d["clocks"]))
for k, v in d["kernels"].items():
    print("  %-9s %s" % (k, v))
print("  pll:", d["pll"])
print("  roofline:", {k: d["roofline"][k] for k in ("achieved", "frac", "issue_frac", "share_of_step") if k in d["roofline"]})
if d.get("cpu_baseline"): print("  cpu:", d["cpu_baseline"])
