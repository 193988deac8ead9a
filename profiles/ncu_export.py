"""Run on the GPU box after an `ncu --set full --import-source on -o REP` capture of a whole step: the report
itself exceeds what gpurun copies back, so keep (1) the raw page of every launch (metrics per launch) and (2) the
source page of the FIRST instance of every kernel, reduced to SASS text / executed instructions / stall samples.
usage: python profiles/ncu_export.py REP.ncu-rep OUT_PREFIX"""
import csv
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
open(out + "_raw.csv", "w").write(raw)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
keep = ["Source", "Warp Stall Sampling (All Samples)", "Warp Stall Sampling (Not-issued Samples)", "Instructions Executed",
        "Avg. Threads Executed", "L1 Conflicts Shared N-Way", "L2 Theoretical Sectors Global Excessive"]
seen, cur, idx = set(), None, None
with open(out + "_source.csv", "w", newline="") as f:
    w = csv.writer(f)
    for r in csv.reader(src.splitlines()):
        if r and r[0] == "Kernel Name":
            cur = r[1] if r[1] not in seen else None
            if cur:
                seen.add(cur)
                w.writerow(["Kernel Name", cur])
            idx = None
            continue
        if cur is None or not r:
            continue
        if idx is None:
            idx = [r.index(k) for k in keep if k in r]
            w.writerow([r[i] for i in idx])
            continue
        w.writerow([r[i] for i in idx])
print("kernels:", len(seen))
