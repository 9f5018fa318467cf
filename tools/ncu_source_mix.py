"""Dynamic instruction mix and top stall sites from `ncu --page source --csv` output."""
import csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
steps = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
hdr = rows[1]
# (a report with several captured launches repeats the header block per launch: keep the numeric rows)
data = [r for r in rows[2:] if len(r) == len(hdr) and r[hdr.index('Instructions Executed')].strip().isdigit()]
ia, ie, iss = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('Warp Stall Sampling (All Samples)')
tot = sum(int(r[ie]) for r in data); totS = sum(int(r[iss]) for r in data)
print("total warp instructions", tot, "per step %.1f" % (tot / steps), "samples", totS)
ops = {}
for r in data:
    m = re.match(r'\s*(@!?U?P\d\s+)?([A-Z0-9_]+)', r[ia])
    op = m.group(2) if m else '?'
    e = ops.setdefault(op, [0, 0]); e[0] += int(r[ie]); e[1] += int(r[iss])
for k, v in sorted(ops.items(), key=lambda kv: -kv[1][0])[:30]:
    print("%-12s %11d %8.1f %6.1f%% stall %5.1f%%" % (k, v[0], v[0] / steps, 100 * v[0] / tot, 100 * v[1] / totS))
print()
for r in sorted(data, key=lambda r: -int(r[iss]))[:12]:
    print("%6s %5.1f%%  %s" % (r[iss], 100 * int(r[iss]) / totS, r[ia].strip()[:90]))
