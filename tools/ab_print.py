"""Print the headline numbers of bench JSON files (A/B runs): python tools/ab_print.py file.json ..."""
import json
import sys

for f in sys.argv[1:]:
    try:
        d = json.load(open(f))
    except Exception as e:
        print(f, "unreadable:", e)
        continue
    x = d.get("extra", {})
    print("%-28s step %.3f ms  e2e %.0f  tdnn %.3f ms  fresh %.3f ms  clocks %s %s" % (
        f.split("/")[-1], d["ms_per_step"], d["e2e"]["value"], x.get("tdnn_xvector", {}).get("ms_per_step", 0),
        x.get("fresh_geometry", {}).get("ms_per_step", 0), d["clocks"]["sm_mhz"], d["clocks"]["reasons"]))
