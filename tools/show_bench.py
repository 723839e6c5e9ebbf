import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("batched: value %.4g e2e %.4g ms/step %.1f launches %s clocks %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["gpu_launches"], d["clocks"]))
print("  roofline", {k: d["roofline"][k] for k in ("achieved", "frac", "ms_per_launch")})
if d.get("cpu_baseline"): print("  cpu", d["cpu_baseline"])
if d.get("bowl_1m"): print("bowl_1m:", {k: d["bowl_1m"][k] for k in ("value", "ms", "frac_of_nominal_8TBs")})
l = d.get("large_n")
if l:
    print("large_n:", {k: l.get(k) for k in ("value", "ms_per_update", "updates_per_s", "cuda_graph", "last_status", "last_status_sweep_phase", "reduce_mode", "collective")})
    print("  roofline", {k: l["roofline"][k] for k in ("achieved", "frac", "frac_of_nominal_8TBs", "ms_per_launch")})
    print("  e2e", l.get("e2e")); print("  cpu", l.get("cpu_baseline"))
