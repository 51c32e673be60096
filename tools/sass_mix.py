"""Executed-instruction mix by SASS opcode from an `ncu --page source --csv` export: python tools/sass_mix.py file.csv [top]"""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]; ix = {h: i for i, h in enumerate(hdr)}
cnt, samp, static = collections.Counter(), collections.Counter(), collections.Counter()
for r in rows[hi + 1:]:
    if len(r) < len(hdr) - 5: continue
    parts = r[ix["Source"]].split()
    op = parts[1] if parts[0].startswith("@") else parts[0]
    op = "SHFL" if op.startswith("SHFL") else op.split(".")[0] + ("" if not op.startswith("IMAD.MOV") else ".MOV")
    n, s = int(r[ix["Instructions Executed"]]), int(r[ix["# Samples"]])
    cnt[op] += n; samp[op] += s; static[op] += 1
tot, ts = sum(cnt.values()), sum(samp.values())
print("total warp instructions", tot, "samples", ts)
for op, n in cnt.most_common(top):
    print(f"{op:10s} {n / tot * 100:6.2f}% inst  {samp[op] / ts * 100:6.2f}% samples  static {static[op]}")
