"""Top stall locations of one launch of an .ncu-rep (source page): python tools/ncu_stalls.py rep [launch_index] [top_n]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; idx = int(sys.argv[2]) if len(sys.argv) > 2 else 0; topn = int(sys.argv[3]) if len(sys.argv) > 3 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(idx), "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
# sections: "Kernel Name" row, header row, data rows
secs, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "data": []}
        secs.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = r
    elif cur is not None and len(r) == len(cur["hdr"]):
        cur["data"].append(r)
s = secs[0]
hdr, data = s["hdr"], s["data"]
print(s["name"][:150])
isrc, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not" not in h]
tot = sum(int(r[isamp] or 0) for r in data)
print("samples", tot, "sass instructions", len(data), "warp-instructions executed", sum(int(r[iex] or 0) for r in data))
agg = {}
for r in data:
    for i in stall:
        agg[hdr[i]] = agg.get(hdr[i], 0) + int(r[i] or 0)
print("  ".join("%s=%.1f%%" % (k[6:], 100 * v / max(tot, 1)) for k, v in sorted(agg.items(), key=lambda x: -x[1])[:9]))
for r in sorted(data, key=lambda r: -int(r[isamp] or 0))[:topn]:
    st = sorted([(int(r[i] or 0), hdr[i][6:]) for i in stall], reverse=True)[:2]
    print("%6s %5.1f%% ex=%9s %-64s %s" % (r[isamp], 100 * int(r[isamp]) / max(tot, 1), r[iex], r[isrc][:64], st))
