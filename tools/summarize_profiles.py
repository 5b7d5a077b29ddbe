"""Regenerate profiles/r01_step_breakdown.txt and profiles/r01_traffic.json from the raw captures
(profiles/r01_launches_bench.csv, gpurun_out/prof_round.ncu-rep).  Usage: python tools/summarize_profiles.py"""
import csv, json, os, re, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rows = list(csv.reader(open(os.path.join(ROOT, "profiles", "r01_launches_bench.csv"))))
h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[h]; ix = {k: i for i, k in enumerate(hdr)}
clean = lambda n: re.sub(r"\(.*", "", n).replace("void ", "").replace("dinox::", "").replace("gemm::", "")
L = [(clean(r[ix["Kernel Name"]]), float(r[ix["Metric Value"]]) / 1000) for r in rows[h + 1:] if len(r) >= len(hdr)]
hg = [i for i, (n, v) in enumerate(L) if "EpiGradT" in n]
s, e = hg[-3], hg[-2]
agg = {}
for n, v in L[s:e]:
    a = agg.setdefault(n[:70], [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(v for _, v in L[s:e])
out = ["# one steady-state micro-step (between two head_grad launches) from r01_launches_bench.csv",
       "# ncu serialises launches and runs them cold: compare SHARES, not absolute times",
       f"# launches {e - s}, sum of device times {tot:.1f} us", ""]
for n, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append(f"{v:9.1f} us  {100 * v / tot:5.1f} %  x{c:<3d} {n}")
open(os.path.join(ROOT, "profiles", "r01_step_breakdown.txt"), "w").write("\n".join(out) + "\n")
print("\n".join(out[:12]))
rep = os.path.join(ROOT, "gpurun_out", "prof_round.ncu-rep")
if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(raw.splitlines())); hd, un = rr[0], rr[1]
    names = ["head_stats_student", "head_stats_teacher_patch", "head_grad", "gemm_dW2", "gemm_dH"]
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}
    tr = {}
    for nm, r in zip(names, rr[2:]):
        d = dict(zip(hd, r)); u = dict(zip(hd, un))
        b = lambda k: float(d[k]) * scale[u[k]]
        tr[nm] = {"dram_bytes_read": b("dram__bytes_read.sum"), "dram_bytes_write": b("dram__bytes_write.sum"),
                  "l2_to_sm_bytes": b("l1tex__m_xbar2l1tex_read_bytes.sum"),
                  "duration_us_under_ncu": float(d["gpu__time_duration.sum"]) * {"us": 1, "ms": 1000}[u["gpu__time_duration.sum"]],
                  "tensor_op_pct_of_hw_peak": float(d["sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed"]),
                  "tensor_pipe_active_pct": float(d["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]),
                  "sm_clock_mhz_under_ncu": float(d["sm__cycles_elapsed.max"]) / (float(d["gpu__time_duration.sum"]) * {"us": 1, "ms": 1000}[u["gpu__time_duration.sum"]])}
    json.dump({"source": "ncu --set full --clock-control none, tools/probe_prof.py (C2 shapes), gpurun_out/prof_round.ncu-rep",
               "kernels": tr}, open(os.path.join(ROOT, "profiles", "r01_traffic.json"), "w"), indent=1)
    for k, v in tr.items():
        print(k, {a: round(b, 1) for a, b in v.items() if "bytes" not in a}, "dram MB", round((v["dram_bytes_read"] + v["dram_bytes_write"]) / 1e6))
