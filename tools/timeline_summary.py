"""Summarise a tools/timeline.py CSV: per-step span, busy time per stream, union busy time, idle gaps,
concurrency histogram and per-kernel totals (second replay only)."""
import csv
import re
import sys
from collections import defaultdict

rows = list(csv.DictReader(open(sys.argv[1])))
ev = [(float(r["start_us"]), float(r["dur_us"]), r["stream"], r["name"]) for r in rows]
# split into the two replays at the adam kernel
adam = [i for i, e in enumerate(ev) if "adam_kernel" in e[3]]
if len(adam) >= 2:
    ev = ev[adam[0] + 1:adam[1] + 1]
t0 = ev[0][0]
t1 = max(s + d for s, d, _, _ in ev)
print("step span %.1f us, %d kernels" % (t1 - t0, len(ev)))
streams = defaultdict(float)
for s, d, st, n in ev:
    streams[st] += d
for st, b in sorted(streams.items(), key=lambda x: -x[1]):
    print("  stream %-6s busy %8.1f us (%.1f%%)" % (st, b, 100 * b / (t1 - t0)))
# union / concurrency
pts = []
for s, d, st, n in ev:
    pts.append((s, 1))
    pts.append((s + d, -1))
pts.sort()
conc = defaultdict(float)
cur, last = 0, t0
for t, dlt in pts:
    conc[cur] += t - last
    last = t
    cur += dlt
for c in sorted(conc):
    print("  %d kernels running: %8.1f us (%.1f%%)" % (c, conc[c], 100 * conc[c] / (t1 - t0)))


def short(n):
    n = re.sub(r"^void\s+", "", n)
    n = re.sub(r"b200u::", "", n)
    m = re.match(r"(\w+)<(.*?)>", n)
    if m and "gemm_tc" in n:
        return "gemm_tc<%s>" % m.group(2).replace(" ", "")[:40]
    return n.split("(")[0][:60]


tot = defaultdict(lambda: [0, 0.0])
for s, d, st, n in ev:
    k = short(n)
    tot[k][0] += 1
    tot[k][1] += d
print("per kernel (sum of durations, overlapping kernels counted fully):")
for k, (c, d) in sorted(tot.items(), key=lambda x: -x[1][1])[:40]:
    print("  %-62s n=%4d tot=%8.1f us avg=%7.2f" % (k, c, d, d / c))
print("  sum of all durations: %.1f us" % sum(d for _, d in tot.values()))
