"""Static SASS evidence per kernel of the built library (no GPU needed): instruction count and the Blackwell-specific mnemonics
(tcgen05.mma -> UTC*MMA, tcgen05.ld -> LDTM, TMA -> UTMALDG, FP64 DFMA ...).  python tools/sass_evidence.py > profiles/rNN_sass_mix.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "n_hexane_pyrolysis_surrogate_reactor_model_b200", "libcrnn_pfr_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
kern, name = collections.OrderedDict(), None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = m.group(1); kern[name] = collections.Counter(); continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and name:
        kern[name][m.group(1).split(".")[0]] += 1
KEY = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "HMMA", "DFMA", "DMUL", "DADD", "DSETP", "FFMA", "MUFU", "LDS", "STS", "SHFL", "LDG", "STG", "ATOMG", "I2F", "F2I", "BAR"]
print(f"# static SASS of {os.path.relpath(lib, ROOT)} (cuobjdump -sass, sm_100a): instructions per kernel and selected mnemonics")
print("# tcgen05.mma = UTCHMMA (kind::tf32/f16), tcgen05.ld = LDTM, cp.async.bulk.tensor (TMA) = UTMALDG, mbarrier = SYNCS")
for n, c in sorted(kern.items(), key=lambda kv: -sum(kv[1].values())):
    tot = sum(c.values())
    sel = "  ".join(f"{k} {c[k]}" for k in KEY if c[k])
    print(f"{tot:7d}  {demangle(n)[:110]}\n         {sel}")
