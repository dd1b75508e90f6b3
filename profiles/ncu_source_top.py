"""Top stall sites of one kernel from an ncu report (source page, SASS view).

    python profiles/ncu_source_top.py <report.ncu-rep> [N]

Prints the N SASS instructions with the most warp-stall samples, the dominant stall reason of each, and totals per reason
and per opcode class -- the per-instruction view the summary metrics cannot give."""
import csv
import subprocess
import sys
from collections import Counter

rep = sys.argv[1]
N = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
recs = []
for r in rows[2:]:
    if len(r) < len(hdr): continue
    try:
        smp = int(r[ix["# Samples"]])
    except ValueError:
        continue
    st = {h: int(r[ix[h]] or 0) for h in stall_cols}
    recs.append((smp, int(r[ix["Instructions Executed"]] or 0), r[ix["Source"]].strip(), st, len(recs)))
tot = sum(s for s, _, _, _, _ in recs)
print("total samples %d over %d instructions, %d warp-instructions executed" % (tot, len(recs), sum(e for _, e, _, _, _ in recs)))
by_reason = Counter()
for s, e, src, st, _ in recs:
    for k, v in st.items(): by_reason[k] += v
print("by reason:", ", ".join("%s %.1f%%" % (k.replace("stall_", ""), 100.0 * v / max(1, sum(by_reason.values()))) for k, v in by_reason.most_common(8)))
by_op = Counter(); ex_op = Counter()
for s, e, src, st, _ in recs:
    op = src.split()[0] if not src.startswith("@") else src.split()[1]
    op = op.split(".")[0]
    by_op[op] += s; ex_op[op] += e
print("by opcode (samples%, executed): " + ", ".join("%s %.1f%% (%d)" % (k, 100.0 * v / tot, ex_op[k]) for k, v in by_op.most_common(14)))
print("%6s %6s %9s  %-14s %s" % ("idx", "smp%", "executed", "top stall", "SASS"))
for s, e, src, st, i in sorted(recs, key=lambda t: -t[0])[:N]:
    top = max(st.items(), key=lambda kv: kv[1])
    print("%6d %5.1f%% %9d  %-14s %s" % (i, 100.0 * s / tot, e, top[0].replace("stall_", ""), src[:110]))
