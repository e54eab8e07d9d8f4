"""Summarise an .ncu-rep (raw page): per launch kernel name, duration, DRAM bytes read / written, tensor-pipe %,
warps active %, registers. Optional names (in launch order) label the rows; --json writes the mean DRAM traffic
per launch (bench.py roofline.traffic reads profiles/r02_gemm_dram_traffic.json).
    python tools/ncu_summary.py prof.ncu-rep [--names a,b,c] [--json out.json]
"""
import csv
import io
import json
import subprocess
import sys

rep = sys.argv[1]
names = None
out_json = None
for i, a in enumerate(sys.argv):
    if a == "--names":
        names = sys.argv[i + 1].split(",")
    if a == "--json":
        out_json = sys.argv[i + 1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
data = rows[2:]


def col(name):
    return hdr.index(name) if name in hdr else None


cols = {"kernel": col("Kernel Name"), "dur": col("gpu__time_duration.sum"), "rd": col("dram__bytes_read.sum"),
        "wr": col("dram__bytes_write.sum"), "tensor": col("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
        "warps": col("sm__warps_active.avg.pct_of_peak_sustained_active"), "regs": col("launch__registers_per_thread"),
        "issue": col("smsp__issue_active.avg.pct_of_peak_sustained_active")}
units = rows[1]


def scale(name):
    i = col(name)
    u = units[i] if i is not None else ""
    return {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}.get(u, 1.0)


srd, swr = scale("dram__bytes_read.sum"), scale("dram__bytes_write.sum")
print("%-22s %-44s %9s %10s %10s %8s %7s %6s %5s" % ("shape", "kernel", "dur_us", "dram_rd_MB", "dram_wr_MB", "tensor%", "warps%", "issue%", "regs"))
tot, n = 0.0, 0
per = {}
for i, r in enumerate(data):
    k = r[cols["kernel"]][:44]
    rd = float(r[cols["rd"]]) * srd
    wr = float(r[cols["wr"]]) * swr
    dur = float(r[cols["dur"]])
    du = units[cols["dur"]]
    dur_us = dur * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3}.get(du, 1.0)
    nm = names[i] if names and i < len(names) else ""
    print("%-22s %-44s %9.2f %10.2f %10.2f %8s %7s %6s %5s" % (nm, k, dur_us, rd / 1e6, wr / 1e6,
          r[cols["tensor"]][:6] if cols["tensor"] is not None else "", r[cols["warps"]][:6] if cols["warps"] is not None else "",
          r[cols["issue"]][:6] if cols["issue"] is not None else "", r[cols["regs"]] if cols["regs"] is not None else ""))
    tot += rd + wr
    n += 1
    per[nm or str(i)] = {"dur_us_ncu": round(dur_us, 2), "dram_read_bytes": rd, "dram_write_bytes": wr}
if out_json and n:
    json.dump({"source": rep, "mean_bytes_per_launch": tot / n, "launches": n, "per_launch": per}, open(out_json, "w"), indent=1)
    print("wrote", out_json, "mean bytes per launch %.3e" % (tot / n))
