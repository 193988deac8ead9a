#!/bin/bash
# gpurun_out/$1_* (profiles/final_capture.sh) -> the tracked round-2 artefacts under profiles/
p=gpurun_out/${1:-r2g}
set -e
for f in train infer reference; do tail -n 1 ${p}_bench_$f.json > profiles/r2_bench_${f}_final.json; done
cp ${p}_bench_launches.csv profiles/r2_bench_launches.csv
python profiles/summarize_launches.py profiles/r2_bench_launches.csv > profiles/r2_bench_launches.txt
python profiles/ncu_step_summary.py ${p}_step_raw.csv "ncu --set full --clock-control none --import-source on --profile-from-start off python profiles/prof_step.py 3 (one warm un-graphed train step at the final round-2 code: batch 8192, D=128, L=4, bf16; one B200)" > profiles/r2_final_ncu_full_summary.txt
python profiles/extract_traffic.py profiles/r2_traffic.json ${p}_step_raw.csv > /dev/null
cp ${p}_gputests.log profiles/r2_gputests_final.log
head -3 profiles/r2_bench_launches.txt; head -6 profiles/r2_final_ncu_full_summary.txt | cut -c1-180
