import json, sys
for p in sys.argv[1:]:
    try:
        j = json.loads(open(p).read().strip().splitlines()[-1])
    except Exception as e:
        print(p, "unreadable:", e); continue
    print(p, "| value %.1f %s | ms/step %.3f | e2e %s | launches %s" % (
        j["value"], j["unit"], j["ms_per_step"], (j.get("e2e") or {}).get("value"), j.get("gpu_launches")))
    print("   stages:", j.get("stages_ms"))
    r = j.get("roofline") or {}
    print("   roofline:", r.get("kernel"), "frac %.3f" % r.get("frac", 0), {k: round(v["GB/s"]) for k, v in (r.get("all") or {}).items()})
    print("   clocks:", j.get("clocks"), "cpu:", j.get("cpu_baseline"))
