"""Static SASS evidence per kernel of the built library: how many tcgen05.mma (UTC*MMA), tcgen05.ld/st (LDTM/STTM), TMA
(UTMALDG / UTMASTG / UBLKCP), tcgen05.commit / mbarrier (UTCBAR / SYNCS) and legacy-tensor (HMMA) instructions each
kernel contains, plus registers (from the ptxas logs when present).
usage: python profiles/sass_summary.py [lib.so] > profiles/sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "graphchem_b200", "lib", "libgraphchem_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
dem = {}
names = re.findall(r"Function : (\S+)", sass)
if names:
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    dem = dict(zip(names, out))
pat = collections.OrderedDict([("UTC*MMA", r"\bUTC\w*MMA"), ("LDTM", r"\bLDTM"), ("STTM", r"\bSTTM"), ("UTMALDG", r"\bUTMALDG"),
                               ("UTMASTG", r"\bUTMASTG"), ("UBLKCP", r"\bUBLKCP"), ("UTCBAR", r"\bUTCBAR"), ("SYNCS", r"\bSYNCS"),
                               ("HMMA", r"\bHMMA"), ("MUFU", r"\bMUFU"), ("instr", r"^\s+/\*[0-9a-f]{4,}\*/")])
rows, cur, cnt = [], None, None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        if cur:
            rows.append((cur, cnt))
        cur, cnt = m.group(1), collections.Counter()
        continue
    if cur:
        for k, p in pat.items():
            if re.search(p, line):
                cnt[k] += 1
if cur:
    rows.append((cur, cnt))
print(f"static SASS instruction counts per kernel of {os.path.relpath(lib, ROOT)} (cuobjdump -sass; sm_100a)")
print(f"{'kernel':96s} " + " ".join(f"{k:>8s}" for k in pat))
tot = collections.Counter()
for name, c in sorted(rows, key=lambda r: -(r[1]['UTC*MMA'] * 1000 + r[1]['UTMALDG'] * 10 + r[1]['LDTM'])):
    d = dem.get(name, name)
    d = re.sub(r"\(.*", "", d).replace("void ", "").replace("gcb::", "")
    tot.update(c)
    print(f"{d[:96]:96s} " + " ".join(f"{c[k]:8d}" for k in pat))
print(f"{'TOTAL':96s} " + " ".join(f"{tot[k]:8d}" for k in pat))
