#!/bin/bash
# round-2 GPU job M: parallel align / finalize / cluster reparam kernels: tests + bench; stock-autocast gradient context
mkdir -p gpurun_out/r2m
O=gpurun_out/r2m
timeout 1200 python -m pytest tests -m gpu -q -x > $O/tests_all.log 2>&1; echo "all gpu tests rc=$?" >> $O/rc.txt
tail -4 $O/tests_all.log >> $O/rc.txt
timeout 300 python bench.py --config C2 --steps 20 --warmup 5 --skip-cpu --skip-library > $O/bench_C2.json 2> $O/bench_C2.err; echo "bench C2 rc=$?" >> $O/rc.txt
timeout 300 python bench.py --config C5 --steps 20 --warmup 5 --skip-cpu --skip-library > $O/bench_C5.json 2> $O/bench_C5.err; echo "bench C5 rc=$?" >> $O/rc.txt
timeout 300 python tools/autocast_cosine.py --backbone vgg --seq 6 --batch 32 > $O/autocast_vgg.txt 2>&1; echo "autocast vgg rc=$?" >> $O/rc.txt
timeout 300 python tools/autocast_cosine.py --backbone dcgan --channels 1 --seq 30 --batch 16 > $O/autocast_dcgan.txt 2>&1; echo "autocast dcgan rc=$?" >> $O/rc.txt
timeout 300 python tools/profile_step.py --steps 3 --calls > $O/calls_C2.txt 2>&1
cat $O/rc.txt; cat $O/autocast_vgg.txt $O/autocast_dcgan.txt | tail -30
python - <<'PY'
import json
for c in ("C2", "C5"):
    try:
        d = json.load(open(f"gpurun_out/r2m/bench_{c}.json"))
        det = d["e2e"].get("detail") or {}
        print(c, "device", round(d["ms_per_step"], 3), "e2e ms", round(det.get("ms_per_step", 0), 3), "e2e", round(d["e2e"]["value"]), "value", round(d["value"]), d.get("phases_ms"))
    except Exception as e:
        print(c, "failed", e)
PY
