import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("inference", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1))
t = d.get("train")
if t:
    print("train", {k: v for k, v in t.items() if k not in ("kernels", "config", "clocks", "cpu_baseline")})
    for k in t["kernels"]:
        print("   ", k)
