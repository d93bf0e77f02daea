"""Summarise gpurun_out/<R>_launches.csv and gpurun_out/<R>_hot.ncu-rep into profiles/ (tracked)."""
import csv, io, json, os, subprocess, sys
R = sys.argv[1] if len(sys.argv) > 1 else "r01"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out"); P = os.path.join(ROOT, "profiles")
os.makedirs(P, exist_ok=True)

# ---- launch list: per-kernel totals and shares
rows = [r for r in csv.reader(l for l in open(os.path.join(G, f"{R}_launches.csv")) if not l.startswith("=="))]
hdr = rows[0]; ik = hdr.index("Kernel Name"); iv = hdr.index("Metric Value"); iu = hdr.index("Metric Unit")
agg = {}
for r in rows[1:]:
    if len(r) <= iv: continue
    name = r[ik].split("(")[0].replace("void ", "").replace("fftconv::", "")
    v = float(r[iv].replace(",", "")); u = r[iu]
    v_us = v / 1000.0 if u in ("ns", "nsecond") else (v * 1000.0 if u in ("ms", "msecond") else v)
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v_us
tot = sum(a[1] for a in agg.values())
with open(os.path.join(P, f"{R}_launches_summary.md"), "w") as f:
    f.write(f"# {R}: ncu launch list of `python bench.py --steps 2 --warmup 1 --no-cpu --no-extras`\n\n"
            "`ncu --metrics gpu__time_duration.sum --clock-control none` (cold-cache, serialised: compare SHARES).\n"
            "Includes the warm-up, timed, profiling and e2e legs of bench.py.\n\n| kernel | launches | total us | share |\n|---|---|---|---|\n")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"| {k} | {n} | {t:.1f} | {t / tot:.3f} |\n")
print(open(os.path.join(P, f"{R}_launches_summary.md")).read())

# ---- full capture: key counters per kernel
rep = os.path.join(G, f"{R}_hot.ncu-rep")
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_bytes.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__sass_inst_executed_op_local_ld.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "dram__bytes.sum.per_second"]
traffic = {}
with open(os.path.join(P, f"{R}_hot_kernels_ncu.md"), "w") as f:
    f.write(f"# {R}: `ncu --set full --clock-control none` of the hot kernels (one launch each, C2 workload, one chunk of 1000 templates)\n\n")
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "").replace("fftconv::", "")
        f.write(f"## {name}\n\n| metric | value | unit |\n|---|---|---|\n")
        vals = {}
        for w in want:
            if w in hdr:
                i = hdr.index(w); vals[w] = r[i]; f.write(f"| {w} | {r[i]} | {units[i]} |\n")
        st = sorted(((float(r[i]), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""))
                     for i, h in enumerate(hdr) if "smsp__average_warps_issue_stalled" in h and h.endswith("per_issue_active.ratio")
                     and "not_issued" not in h), reverse=True)[:6]
        f.write("\nTop stall reasons (warps per issue-active cycle): " + ", ".join(f"{n} {v:.2f}" for v, n in st) + "\n\n")
        def mb(x, u): return float(x.replace(",", "")) * {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1}[u]
        ir, iw = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        key = name.split("<")[0]
        key = {"os_inverse_tma": "os_inverse", "os_inverse_z": "os_inverse"}.get(key, key)       # bench.py names kernels by role
        traffic[key] = mb(r[ir], units[ir]) + mb(r[iw], units[iw])
json.dump(traffic, open(os.path.join(P, "traffic.json"), "w"), indent=1)
print(open(os.path.join(P, f"{R}_hot_kernels_ncu.md")).read()[:3000])
