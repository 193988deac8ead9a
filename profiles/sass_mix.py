"""Per-kernel dynamic instruction mix from an ncu report's source page (SASS view): executed warp-instructions per
opcode, and the stall samples per opcode.  Usage: python profiles/sass_mix.py report.ncu-rep [kernel-substring]"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
want = sys.argv[2] if len(sys.argv) > 2 else ""
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
kern, hdr, seen = None, None, set()
stats = None
def flush():
    if kern is None or stats is None: return
    tot = sum(v[0] for v in stats.values()); samp = sum(v[1] for v in stats.values())
    print(f"== {kern[:110]}\n   executed warp-instructions {tot:,}  stall samples {samp:,}")
    for op, (n, s) in sorted(stats.items(), key=lambda kv: -kv[1][0])[:28]:
        print(f"   {op:28s} {n:12,} {100*n/max(tot,1):5.1f}%   samples {100*s/max(samp,1):5.1f}%")
for r in rows:
    if len(r) >= 2 and r[0] == "Kernel Name":
        flush()
        kern = r[1] if (want in r[1] and r[1] not in seen) else None
        if kern: seen.add(kern)
        stats = collections.defaultdict(lambda: [0, 0]); hdr = None
        continue
    if kern is None: continue
    if r and r[0] == "Address":
        hdr = {h: i for i, h in enumerate(r)}; continue
    if hdr is None or len(r) < len(hdr): continue
    src = r[hdr["Source"]].strip()
    toks = src.split()
    if not toks: continue
    op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
    op = ".".join(op.split(".")[:2]) if op.startswith(("LDS", "STS", "LDG", "STG", "MUFU", "BAR", "HSET2", "F2FP")) else op.split(".")[0]
    try:
        n = int(r[hdr["Instructions Executed"]]); s = int(r[hdr["# Samples"]])
    except ValueError:
        continue
    stats[op][0] += n; stats[op][1] += s
flush()
