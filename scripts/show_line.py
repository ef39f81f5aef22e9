"""One-line digest of a bench.py JSON line: python scripts/show_line.py file.json"""
import json
import sys

for path in sys.argv[1:]:
    try:
        d = json.loads(open(path).read().strip().splitlines()[-1])
    except Exception as e:   # noqa: BLE001
        print(path, "unreadable:", e)
        continue
    t = d.get("train") or {}
    r = d.get("roofline") or {}
    print(f"{path}: {d.get('value')} {d.get('unit')} ({d.get('ms_per_step')} ms), e2e {d.get('e2e', {}).get('value')}, "
          f"roofline {r.get('kernel')} {r.get('frac')}, step_frac {d.get('step_frac')}, train {t.get('value')} steps/s "
          f"({t.get('ms')} ms, frac {t.get('frac')}, step_frac {t.get('step_frac')}), clocks {d.get('clocks')}")
