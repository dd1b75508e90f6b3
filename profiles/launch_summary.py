"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list:  python profiles/launch_summary.py <csv> [title]"""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0].isdigit() and r[12] == "gpu__time_duration.sum"]
tot = collections.OrderedDict()
for r in rows:
    name = r[4].split("(")[0].replace("void ", "")
    t = float(r[14].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0}[r[13]]
    a = tot.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += t
total = sum(v[1] for v in tot.values())
print("%s: total kernel time %.2f ms, %d launches" % (sys.argv[2] if len(sys.argv) > 2 else sys.argv[1], total, len(rows)))
for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print("%-66s %4d %9.3f ms %5.1f%%" % (k[:66], v[0], v[1], 100 * v[1] / total))
