#!/bin/bash
# Final captures of a round on one B200 (run through gpurun): GPU tests, the three bench lines, the ncu launch list of
# bench.py itself, one `ncu --set full` capture of a warm un-graphed train step exported to CSV.  Output: gpurun_out/$1_*
p=gpurun_out/${1:-final}
(time python -m pytest tests -q -m gpu 2>&1 | tail -4) > ${p}_gputests.log 2>&1
python bench.py > ${p}_bench_train.json 2> ${p}_bench_train.err
python bench.py --config infer > ${p}_bench_infer.json 2> ${p}_bench_infer.err
python bench.py --impl reference > ${p}_bench_reference.json 2> ${p}_bench_reference.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file ${p}_bench_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > ${p}_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -f -o ${p}_step \
    python profiles/prof_step.py 3 > ${p}_ncu_step.log 2>&1
python profiles/ncu_export.py ${p}_step.ncu-rep ${p}_step > /dev/null 2>&1
ls -la ${p}_step.ncu-rep
rm -f ${p}_step.ncu-rep
cat ${p}_gputests.log
for f in train infer reference; do python - <<P
import json
d=json.loads(open("${p}_bench_$f.json").read().strip().splitlines()[-1])
print("$f", d.get("value"), (d.get("e2e") or {}).get("value"), (d.get("roofline") or {}).get("kernel"), (d.get("roofline") or {}).get("frac"), ((d.get("roofline") or {}).get("step") or {}).get("frac"), d.get("launches_per_step"))
P
done
