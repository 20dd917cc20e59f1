"""Per-kernel totals of the LAST extraction step of an `ncu --metrics gpu__time_duration.sum --csv` launch list (a step
starts at its geometry kernels / front-end launch).  Usage: python tools/step_summary.py launches.csv"""
import sys
from collections import OrderedDict
sys.path.insert(0, __file__.rsplit("/", 1)[0])
from parse_launches import load

rows = load(sys.argv[1])
names = [r[0] for r in rows]
start = max(i for i, n in enumerate(names) if n.startswith("frontend_kernel"))
while start > 0 and names[start - 1].startswith(("planmeta_kernel", "plan_activate_kernel")):
    start -= 1
last = rows[start:]
tot = sum(r[1] for r in last)
agg = OrderedDict()
for k, v, g in last:
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += v
print("launches in the last step: %d   total %.1f us  (ncu --metrics gpu__time_duration.sum --clock-control none: cold-cache, serialised)" % (len(last), tot))
for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("  %-48s x%-3d %9.1f us  %5.1f%%" % (k[:48], n, v, 100 * v / tot))
