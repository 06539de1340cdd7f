#!/bin/bash
# ncu evidence, round 2: launch list of the last training step / inference pass + --set full of the tensor-core kernels.  $1 = tag
TAG=${1:-r02}
mkdir -p gpurun_out
python scripts/prof_step_once.py train > gpurun_out/plain_$TAG.log 2>&1 || { tail -5 gpurun_out/plain_$TAG.log; exit 1; }
# count launches per step: total launches / 5 steps (warm-up steps have the same launch count)
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_all_$TAG.csv python scripts/prof_step_once.py train > gpurun_out/ncu_launch_$TAG.log 2>&1
echo "launch list exit $?"
python - <<PY
import csv
rows=[r for r in csv.reader(l for l in open("gpurun_out/launches_all_$TAG.csv") if not l.startswith("=="))]
hdr=rows[0]; body=[r for r in rows[1:] if len(r)==len(hdr)]
n=len(body); per=n//5
open("gpurun_out/launches_$TAG.csv","w").write("\n".join(",".join('"%s"'%c for c in r) for r in [hdr]+body[n-per:])+"\n")
print("launches total", n, "per step", per)
PY
python scripts/prof_step_once.py train > gpurun_out/plain2_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"tc_fchain|tc_chain_bwd|tc_wgrad" -s 28 -c 7 -o gpurun_out/prof_$TAG -f python scripts/prof_step_once.py train > gpurun_out/ncu_full_$TAG.log 2>&1
echo "full capture (train) exit $?"
python scripts/prof_step_once.py infer > gpurun_out/plain3_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"tc_fchain" -s 3 -c 1 -o gpurun_out/prof_${TAG}inf -f python scripts/prof_step_once.py infer > gpurun_out/ncu_fullinf_$TAG.log 2>&1
echo "full capture (infer) exit $?"
