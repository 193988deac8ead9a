#!/bin/bash
# Interleaved A/B of GCB_* flag sets through step_time.py: ab.sh OUT.jsonl ROUNDS "FLAGS_A" "FLAGS_B" ...   ("-" = no flags)
out=$1; rounds=$2; shift 2
for r in $(seq 1 $rounds); do
  for f in "$@"; do
    if [ "$f" = "-" ]; then env GCB_X=0 python profiles/micro/step_time.py $AB_ARGS >> $out 2>> ${out%.jsonl}.err
    else env $f python profiles/micro/step_time.py $AB_ARGS >> $out 2>> ${out%.jsonl}.err; fi
  done
done
python - "$out" <<'P'
import json, sys
for ln in open(sys.argv[1]):
    d = json.loads(ln)
    print({k: v for k, v in d.get("flags", {}).items() if k != "GCB_X"}, d.get("train_ms_per_step"), d.get("infer_ms_per_step"), d.get("launches_per_step"))
P
