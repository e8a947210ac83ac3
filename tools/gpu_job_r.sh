#!/bin/bash
# round-2 GPU job R: vgg after the shared layout-conversion helper; clean per-launch list of one C2 step (3 steps, last third)
mkdir -p gpurun_out/r2r
O=gpurun_out/r2r
timeout 600 python -m pytest tests/test_vgg_gpu.py tests/test_measured_gpu.py tests/test_step_gpu.py tests/test_kernels_gpu.py -m gpu -q -x > $O/tests.log 2>&1; echo "vgg/measured/step/kernel tests rc=$?" >> $O/rc.txt
tail -3 $O/tests.log >> $O/rc.txt
timeout 400 python bench.py --config C3 --steps 10 --warmup 3 --skip-cpu > $O/bench_C3.json 2> $O/bench_C3.err; echo "bench C3 rc=$?" >> $O/rc.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_C2.csv python tools/profile_step.py --steps 3 > $O/launches_C2.out 2>&1
python tools/summarize_launches.py $O/launches_C2.csv 3 > $O/launches_C2.txt 2>&1
timeout 400 python tools/profile_step.py --steps 3 --calls --backbone vgg_64 --channels 3 --batch 128 > $O/calls_C3.txt 2>&1
cat $O/rc.txt
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2r/bench_C3.json"))
print("C3", round(d["ms_per_step"], 3), round(d["value"]), "e2e", round(d["e2e"]["value"]), "roof", round(d["roofline"]["frac"], 3), d["clocks"])
PY
head -12 gpurun_out/r2r/launches_C2.txt
