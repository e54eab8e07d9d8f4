"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel."""
import collections, csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
h = rows[hdr]; ki = h.index('Kernel Name'); vi = h.index('Metric Value'); ui = h.index('Metric Unit')
agg = collections.OrderedDict(); n = 0
for r in rows[hdr + 1:]:
    if len(r) <= vi: continue
    v = float(r[vi].replace(',', '')); u = r[ui]
    v = v / 1000.0 if u in ('ns', 'nsecond') else (v * 1000.0 if u in ('ms', 'msecond') else v)
    name = re.sub(r'b200u::', '', r[ki]); name = re.sub(r'\(.*', '', name); name = re.sub(r'^void ', '', name)
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v; n += 1
tot = sum(a[1] for a in agg.values())
print("launches=%d total=%.1f us" % (n, tot))
for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 30]:
    print("%-78s n=%4d tot=%9.1f us avg=%8.2f us %5.1f%%" % (k[:78], c, t, t / c, 100 * t / tot))
