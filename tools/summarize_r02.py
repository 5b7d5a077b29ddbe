"""Round-2 profile summaries from the raw captures under gpurun_out/ (no GPU needed):
  profiles/r02_step_breakdown.txt   one steady-state micro-step cut out of the ncu launch list of `python bench.py`
  profiles/r02_traffic.json         DRAM / L2 bytes, tensor-op utilisation per big kernel (ncu --set full capture)
Usage: python tools/summarize_r02.py [launch_csv] [ncu_rep ...]"""
import csv, json, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
launch_csv = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "r02_launches.csv")
reps = sys.argv[2:]
clean = lambda n: re.sub(r"\(CUtensorMap.*", "", n).replace("void ", "").replace("dinox::", "").replace("gemm::", "").replace("(int)", "").replace("(bool)", "")
if os.path.exists(launch_csv):
    rows = list(csv.reader(open(launch_csv)))
    h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[h]; ix = {k: i for i, k in enumerate(hdr)}
    L = [(clean(r[ix["Kernel Name"]]), float(r[ix["Metric Value"]]) / 1000) for r in rows[h + 1:] if len(r) >= len(hdr)]
    hg = [i for i, (n, v) in enumerate(L) if "EpiGradR" in n]
    s, e = hg[-3], hg[-2]
    agg = {}
    for n, v in L[s:e]:
        a = agg.setdefault(n[:78], [0, 0.0]); a[0] += 1; a[1] += v
    tot = sum(v for _, v in L[s:e])
    ours = sum(c for n, (c, v) in agg.items() if not n.startswith("at::") and "nccl" not in n.lower())
    out = [f"# one steady-state micro-step (between two pass-2 launches) from {os.path.basename(launch_csv)}",
           "# (ncu --metrics gpu__time_duration.sum --clock-control none over `python bench.py --steps 4 --warmup 3`)",
           "# ncu serialises launches and runs them cold: compare SHARES, not absolute times",
           f"# launches {e - s} ({ours} of this library, {e - s - ours} framework), sum of device times {tot:.1f} us", ""]
    for n, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"{v:9.1f} us  {100 * v / tot:5.1f} %  x{c:<3d} {n}")
    open(os.path.join(ROOT, "profiles", "r02_step_breakdown.txt"), "w").write("\n".join(out) + "\n")
    print("\n".join(out))
scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}
tr = {}
key_of = [("EpiStatsT", "head_stats_student"), ("EpiTeachQT", "head_teacher"), ("EpiGradRT", "head_grad"), ("EpiGradT", "head_grad_recompute")]
for rep in reps:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(raw.splitlines())); hd, un = rr[0], rr[1]
    n_store = 0
    for r in rr[2:]:
        d = dict(zip(hd, r)); u = dict(zip(hd, un))
        name = d["Kernel Name"]
        key = next((k for pat, k in key_of if pat in name), None)
        if key is None and "EpiStore" in name:
            key = ["gemm_dW2", "gemm_dH"][n_store] if n_store < 2 else None
            n_store += 1
        if key is None or key in tr:
            continue
        b = lambda k: float(d[k]) * scale[u[k]]
        dur = float(d["gpu__time_duration.sum"]) * {"us": 1, "ms": 1000}[u["gpu__time_duration.sum"]]
        tr[key] = {"kernel": clean(name)[:90], "dram_bytes_read": b("dram__bytes_read.sum"), "dram_bytes_write": b("dram__bytes_write.sum"),
                   "l2_to_sm_bytes": b("l1tex__m_xbar2l1tex_read_bytes.sum"), "duration_us_under_ncu": dur,
                   "dram_gbs_under_ncu": (b("dram__bytes_read.sum") + b("dram__bytes_write.sum")) / dur / 1e3,
                   "tensor_op_pct_of_hw_peak": float(d["sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed"]),
                   "l2_hit_pct": float(d["lts__t_sector_hit_rate.pct"]),
                   "sm_clock_mhz_under_ncu": float(d["sm__cycles_elapsed.max"]) / dur, "report": os.path.basename(rep)}
if tr:
    json.dump({"source": "ncu --set full --clock-control none --import-source on, tools/probe_r02.py once (C2 shapes)",
               "pass2_mode": "readback", "kernels": tr}, open(os.path.join(ROOT, "profiles", "r02_traffic.json"), "w"), indent=1)
    for k, v in tr.items():
        print(k, {a: (round(b, 1) if isinstance(b, float) else b) for a, b in v.items() if "bytes" not in a and a not in ("kernel", "report")},
              "dram MB", round((v["dram_bytes_read"] + v["dram_bytes_write"]) / 1e6))
