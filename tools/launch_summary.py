"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list: python tools/launch_summary.py list.csv > summary.csv"""
import collections, csv, re, sys
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]; ix = {h: i for i, h in enumerate(hdr)}
tot = collections.defaultdict(lambda: [0, 0.0, ""])
for r in rows[1:]:
    name = re.sub(r"\(.*", "", r[ix["Kernel Name"]]).replace("void ", "")
    t = tot[name]; t[0] += 1; t[1] += float(r[ix["Metric Value"]]); t[2] = r[ix["Grid Size"]] + "x" + r[ix["Block Size"]]
total = sum(v[1] for v in tot.values())
w = csv.writer(sys.stdout)
w.writerow(["kernel", "launches", "total_ms", "share_of_gpu_time", "mean_us", "last_grid_x_block"])
for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    w.writerow([k, v[0], f"{v[1] / 1e6:.3f}", f"{v[1] / total:.4f}", f"{v[1] / v[0] / 1e3:.1f}", v[2]])
