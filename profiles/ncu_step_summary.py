"""`ncu -i rep --page raw --csv` dump of one train step (profiles/ncu_export.py) -> one line per kernel class: time, DRAM
bytes read / written, DRAM and tensor-pipe utilisation, resident warps, issue utilisation, executed instructions,
registers, shared memory, grid, block, launches, share of the serialised step.
usage: python profiles/ncu_step_summary.py RAW.csv [title] > profiles/rN_ncu_full_summary.txt"""
import csv
import sys
from collections import OrderedDict

UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
TIME = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}
rows = list(csv.reader(open(sys.argv[1])))
h, units = rows[0], rows[1]
col = {k: i for i, k in enumerate(h)}


def val(r, k, scale=None):
    v = r[col[k]].replace(",", "")
    x = float(v) if v not in ("", "n/a") else 0.0
    if scale:
        x *= scale.get(units[col[k]], 1.0)
    return x


agg = OrderedDict()
for r in rows[2:]:
    name = r[col["Kernel Name"]]
    a = agg.setdefault(name, {"n": 0, "us": 0.0, "rd": 0.0, "wr": 0.0, "dram": 0.0, "tensor": 0.0, "warps": 0.0, "issue": 0.0,
                              "inst": 0.0, "regs": 0.0, "smem": 0.0, "grid": 0.0, "block": 0.0})
    a["n"] += 1
    a["us"] += val(r, "gpu__time_duration.sum", TIME)
    a["rd"] += val(r, "dram__bytes_read.sum", UNIT)
    a["wr"] += val(r, "dram__bytes_write.sum", UNIT)
    a["dram"] += val(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")
    a["tensor"] += val(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")
    a["warps"] += val(r, "sm__warps_active.avg.pct_of_peak_sustained_active")
    a["issue"] += val(r, "smsp__issue_active.avg.pct_of_peak_sustained_active")
    a["inst"] += val(r, "smsp__inst_executed.sum")
    a["regs"] += val(r, "launch__registers_per_thread")
    a["smem"] += val(r, "launch__shared_mem_per_block_dynamic") + val(r, "launch__shared_mem_per_block_static")
    a["grid"] += val(r, "launch__grid_size")
    a["block"] += val(r, "launch__block_size")
total = sum(a["us"] for a in agg.values())
print(sys.argv[2] if len(sys.argv) > 2 else sys.argv[1])
print("per-launch averages; cold cache and serialised under ncu: compare SHARES, not absolute times")
print(f"{'kernel':<62s}{'us':>8s}{'rd MB':>9s}{'wr MB':>9s}{'dram%':>8s}{'tensor%':>9s}{'warps%':>8s}{'issue%':>8s}{'Minst':>8s}{'regs':>6s}{'smemKB':>8s}{'grid':>7s}{'block':>7s}")
for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
    n = a["n"]
    print(f"{name[:60]:<62s}{a['us'] / n:8.1f}{a['rd'] / n / 1e6:9.1f}{a['wr'] / n / 1e6:9.1f}{a['dram'] / n:8.1f}{a['tensor'] / n:9.1f}"
          f"{a['warps'] / n:8.1f}{a['issue'] / n:8.1f}{a['inst'] / n / 1e6:8.1f}{a['regs'] / n:6.0f}{a['smem'] / n:8.1f}{a['grid'] / n:7.0f}{a['block'] / n:7.0f}"
          f"   x{n}  share {100 * a['us'] / total:4.1f}%")
rd, wr = sum(a["rd"] for a in agg.values()), sum(a["wr"] for a in agg.values())
print(f"TOTAL serialised {total:.1f} us over {sum(a['n'] for a in agg.values())} launches; DRAM read {rd / 1e9:.2f} GB + write {wr / 1e9:.2f} GB = {(rd + wr) / 1e9:.2f} GB")
