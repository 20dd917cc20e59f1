"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-launch table of the last pass
and totals per kernel.  Usage: python tools/parse_launches.py file.csv [launches_per_pass]"""
import csv
import re
import sys
from collections import OrderedDict


def load(fn):
    with open(fn) as f:
        lines = [l for l in f if not l.startswith("==")]
    return [(re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").replace("skb::", ""), float(r["Metric Value"].replace(",", "")) / 1000.0,
             r["Grid Size"]) for r in csv.DictReader(lines)]


def main():
    rows = load(sys.argv[1])
    per = int(sys.argv[2]) if len(sys.argv) > 2 else len(rows) // 2
    last = rows[-per:]
    tot = sum(r[1] for r in last)
    agg = OrderedDict()
    for k, v, g in last:
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
    print("launches in pass: %d   total %.1f us" % (per, tot))
    for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("  %-40s x%-3d %9.1f us  %5.1f%%" % (k[:40], n, v, 100 * v / tot))
    if "-v" in sys.argv:
        for i, (k, v, g) in enumerate(last):
            print(i, k[:50], g, "%.1f us" % v)


if __name__ == "__main__":
    main()
