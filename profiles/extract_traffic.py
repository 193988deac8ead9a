"""ncu --set full report(s) -> profiles/rN_traffic.json (names as bench.py / gcb_timing_report use them): per kernel class the average DRAM traffic
(dram__bytes_read.sum + dram__bytes_write.sum) and duration per launch.
usage: python profiles/extract_traffic.py out.json rep1.ncu-rep|raw1.csv [...]"""
import csv
import io
import json
import subprocess
import sys

UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
TIME = {"ns": 1e-3, "us": 1.0, "ms": 1e3}
out = {}
for rep in sys.argv[2:]:
    # a .csv argument is the `ncu -i rep --page raw --csv` dump made on the GPU box (reports of a whole step
    # exceed what gpurun copies back)
    txt = open(rep).read() if rep.endswith(".csv") else subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    h, units = rows[0], rows[1]
    ki, ri, wi, ti = (h.index("Kernel Name"), h.index("dram__bytes_read.sum"), h.index("dram__bytes_write.sum"),
                      h.index("gpu__time_duration.sum"))
    for r in rows[2:]:
        full = r[ki]
        name = full.split("<")[0].split("(")[0].replace("void ", "").split("::")[-1].strip()
        if "TileCommon" in full:  # gcb::tile kernels are timed under a tile_ prefix (bench.py / gcb_timing_report)
            name = "tile_" + name
        elif name in ("gemm_nt_tc_kernel", "gemm_tn_tc_kernel"):  # one timing name per template instance
            import re
            m = re.search(r"<\(?(?:int\))?\s*(\d+),\s*(?:\(int\))?\s*(\d+)(?:,\s*(?:\(bool\))?\s*(\w+))?", full)
            a, b, res = m.group(1), m.group(2), m.group(3)
            name = (f"gemm_nt_tc_k{a}n{b}" + ("_res" if res in ("1", "true") else "")) if name.startswith("gemm_nt") else f"gemm_tn_tc_n{a}k{b}"
        elif name == "readout2_fwd_kernel":
            name = "readout_fwd_kernel"
        elif name == "readout2_bwd_kernel":
            name = "readout_bwd_kernel"
        elif name == "fused_rows_tc_kernel":
            name = "fused_bond_fwd_kernel" if "<0" in full else "fused_atom_bwd_kernel"
        b = float(r[ri]) * UNIT[units[ri]] + float(r[wi]) * UNIT[units[wi]]
        e = out.setdefault(name, {"launches": 0, "dram_bytes": 0.0, "us": 0.0, "source": []})
        e["launches"] += 1
        e["dram_bytes"] += b
        e["us"] += float(r[ti]) * TIME[units[ti]]
        if rep not in e["source"]:
            e["source"].append(rep)
for e in out.values():
    e["dram_bytes_per_launch"] = e["dram_bytes"] / e["launches"]
    e["us_per_launch"] = e["us"] / e["launches"]
json.dump(out, open(sys.argv[1], "w"), indent=1)
for k, e in sorted(out.items()):
    print(f"{k:32s} n={e['launches']:2d} {e['dram_bytes_per_launch'] / 1e6:8.1f} MB/launch {e['us_per_launch']:8.1f} us")
