"""A/B harness for the GCB_* environment switches (one B200): runs `bench.py --no-cpu-baseline` once per flag set in
its own process (the switches are read once per process), interleaved over `--rounds` rounds so that clock / thermal
drift hits every variant alike, and prints one JSON line per variant: resident and end-to-end molecules/s, ms/step,
the streaming legs.  Example:
    python profiles/micro/ab_flags.py --sets "" GCB_NO_PDL=1 "GCB_TILE_Q_ONLY=1 GCB_TN_DEFER=1" --steps 40 --rounds 2
"""
import argparse
import json
import os
import statistics
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def run_once(flags: str, steps: int, warmup: int, batch: int):
    env = dict(os.environ)
    for kv in flags.split():
        k, _, v = kv.partition("=")
        env[k] = v or "1"
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--no-cpu-baseline", "--steps", str(steps), "--warmup", str(warmup),
           "--batch", str(batch)]
    res = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    if res.returncode != 0:
        return {"error": (res.stderr or res.stdout)[-400:]}
    line = [ln for ln in res.stdout.splitlines() if ln.startswith("{")][-1]
    return json.loads(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sets", nargs="+", default=[""], help='flag sets, e.g. "" "GCB_NO_PDL=1" "A=1 B=1"')
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--batch", type=int, default=8192)
    ap.add_argument("--rounds", type=int, default=2)
    args = ap.parse_args()
    runs = {s: [] for s in args.sets}
    for _ in range(args.rounds):
        for s in args.sets:
            runs[s].append(run_once(s, args.steps, args.warmup, args.batch))
    for s, rs in runs.items():
        ok = [r for r in rs if "error" not in r]
        out = {"flags": s or "(default)", "runs": len(rs), "failed": len(rs) - len(ok)}
        if ok:
            out.update({
                "ms_per_step": [round(r["ms_per_step"], 4) for r in ok],
                "ms_per_step_median": round(statistics.median(r["ms_per_step"] for r in ok), 4),
                "value_median": round(statistics.median(r["value"] for r in ok)),
                "e2e_median": round(statistics.median(r["e2e"]["value"] for r in ok)),
                "e2e_stream": [round(r["e2e_stream"]["value"]) for r in ok if r.get("e2e_stream") and "value" in r["e2e_stream"]],
                "e2e_stream_device_pack": [round(r["e2e_stream_device_pack"]["value"]) for r in ok
                                           if r.get("e2e_stream_device_pack") and "value" in r["e2e_stream_device_pack"]],
                "sm_mhz": [r.get("clocks", {}).get("sm_mhz") for r in ok],
            })
        else:
            out["error"] = rs[-1].get("error")
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
