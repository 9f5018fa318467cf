"""Turn the raw ncu outputs in gpurun_out/ into the committed summaries under profiles/.
usage: python tools/make_profiles.py <round-tag>"""
import csv, io, json, os, subprocess, sys
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out_dir = os.path.join(root, "profiles"); os.makedirs(out_dir, exist_ok=True)
src = os.path.join(root, "gpurun_out")

# 1. launch list: per-kernel totals and shares
rows = [r for r in csv.reader(open(os.path.join(src, "launches_%s.csv" % tag))) if len(r) > 5]
hdr_i = [i for i, r in enumerate(rows) if r[0] == "ID"][0]
h, data = rows[hdr_i], rows[hdr_i + 1:]
ki, vi = h.index("Kernel Name"), h.index("Metric Value")
agg = {}
for r in data:
    agg.setdefault(r[ki].split("(")[0], []).append(float(r[vi].replace(",", "")))
total = sum(sum(v) for v in agg.values())
lines = ["# ncu launch list of `python bench.py --steps 2 --warmup 3 --cpu-frames 0 --frames-per-step 128`",
         "# (ncu --metrics gpu__time_duration.sum --clock-control none -c 700; serialised, cold cache: compare SHARES)",
         "%-60s %6s %12s %8s %10s" % ("kernel", "n", "total_us", "share", "avg_us")]
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    lines.append("%-60s %6d %12.1f %7.2f%% %10.1f" % (k[:60], len(v), sum(v) / 1e3, 100 * sum(v) / total, sum(v) / len(v) / 1e3))
open(os.path.join(out_dir, "%s_launches_summary.txt" % tag), "w").write("\n".join(lines) + "\n")
with open(os.path.join(out_dir, "%s_launches.csv" % tag), "w") as f:
    w = csv.writer(f); w.writerow(["id", "kernel", "block", "grid", "duration_ns"])
    bi, gi, ii = h.index("Block Size"), h.index("Grid Size"), h.index("ID")
    for r in data:
        w.writerow([r[ii], r[ki].split("(")[0], r[bi], r[gi], r[vi]])
print("\n".join(lines))

# 2. full capture of the message kernel
rep = os.path.join(src, "prof_%s_bench_iter.ncu-rep" % tag)
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, universal_newlines=True).stdout
rr = list(csv.reader(io.StringIO(raw))); hdr, units, dd = rr[0], rr[1], rr[2:]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum",
        "lts__t_sector_hit_rate.pct", "gpc__cycles_elapsed.avg.per_second", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
summary = ["# ncu --set full --clock-control none -k regex:k_iter_tma -s 120 -c 2 (same bench command; matches k_iter_tmast)"]
vals = {}
for wname in want:
    for i, hh in enumerate(hdr):
        if hh == wname:
            vals[wname] = [r[i] for r in dd]
            summary.append("%-64s %-16s %s" % (wname, units[i], vals[wname]))
for i, hh in enumerate(hdr):
    if "issue_stalled" in hh and hh.endswith("per_issue_active.ratio"):
        summary.append("%-64s %s" % (hh.replace("smsp__average_warps_issue_stalled_", "stall:"), [r[i] for r in dd]))
open(os.path.join(out_dir, "%s_k_iter_tma_ncu_full.txt" % tag), "w").write("\n".join(summary) + "\n")
print("\n".join(summary[:20]))

# 3. DRAM traffic per launch for bench.py's roofline.traffic
def gb(s):
    return float(s) * 1e9
i_r, i_w = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
per = [float(r[i_r]) * scale[units[i_r]] + float(r[i_w]) * scale[units[i_w]] for r in dd]
gi = hdr.index("launch__grid_size")
frames_per_launch = 64       # bench.py --frames-per-step 128: one 128-frame chunk = two 64-frame groups
json.dump({"dram_bytes_per_launch": sum(per) / len(per), "frames_per_launch": frames_per_launch,
           "grid_size": dd[0][gi],
           "source": "profiles/%s_k_iter_tma_ncu_full.txt (dram__bytes_read.sum + dram__bytes_write.sum, mean of %d launches)" % (tag, len(per))},
          open(os.path.join(out_dir, "roofline_traffic.json"), "w"), indent=1)
print(open(os.path.join(out_dir, "roofline_traffic.json")).read())
