"""Table of an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` capture of
tools/probe_hbm.py (no GPU needed): one row per launch of this library's kernels - duration, DRAM bytes read / written,
GB/s at the DRAM pins.  The probe's own flush launches (at::*) are dropped.
Usage: python tools/summarize_hbm_ncu.py [profiles/r02b_hbm_kernels_ncu.csv]"""
import csv, os, re, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r02b_hbm_kernels_ncu.csv")
rows = list(csv.reader(open(path)))
h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[h]
ix = {k: i for i, k in enumerate(hdr)}
scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "us": 1.0, "ms": 1e3, "ns": 1e-3, "s": 1e6}
per, order = {}, []
for r in rows[h + 1:]:
    if len(r) < len(hdr):
        continue
    i = r[ix["ID"]]
    if i not in per:
        per[i] = {"name": r[ix["Kernel Name"]]}
        order.append(i)
    per[i][r[ix["Metric Name"]]] = float(r[ix["Metric Value"]].replace(",", "")) * scale.get(r[ix["Metric Unit"]], 1.0)
clean = lambda n: re.sub(r"\(.*", "", n).replace("void ", "").replace("dinox::", "")
print(f"{'kernel':56s} {'us':>7s} {'dram read MB':>13s} {'dram write MB':>14s} {'GB/s':>8s}")
for i in order:
    d = per[i]
    n = clean(d["name"])
    if n.startswith("at::"):
        continue
    us, rd, wr = d["gpu__time_duration.sum"], d["dram__bytes_read.sum"] / 1e6, d["dram__bytes_write.sum"] / 1e6
    print(f"{n[:56]:56s} {us:7.1f} {rd:13.1f} {wr:14.1f} {(rd + wr) / us * 1e3:8.0f}")
