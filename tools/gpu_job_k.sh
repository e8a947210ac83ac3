#!/bin/bash
# round-2 GPU job K: early loss read-back (zero-copy publish + host polling): tests and e2e A/B
mkdir -p gpurun_out/r2k
O=gpurun_out/r2k
timeout 900 python -m pytest tests/test_dropin_gpu.py tests/test_data_gpu.py tests/test_generate_gpu.py tests/test_measured_gpu.py -m gpu -q -x > $O/tests.log 2>&1; echo "tests rc=$?" >> $O/rc.txt
tail -4 $O/tests.log >> $O/rc.txt
for v in 0 1; do
  P2PVG_EARLY_LOSS=$v timeout 300 python bench.py --config C2 --steps 20 --warmup 5 --skip-cpu --skip-library --skip-phases > $O/bench_C2_early$v.json 2> $O/bench_C2_early$v.err; echo "bench C2 early=$v rc=$?" >> $O/rc.txt
done
P2PVG_EARLY_LOSS=0 timeout 300 python bench.py --config C5 --steps 20 --warmup 5 --skip-cpu --skip-library --skip-phases > $O/bench_C5_early0.json 2> $O/bench_C5_early0.err
timeout 300 python bench.py --config C5 --steps 20 --warmup 5 --skip-cpu --skip-library --skip-phases > $O/bench_C5_early1.json 2> $O/bench_C5_early1.err
cat $O/rc.txt
python - <<'PY'
import json
for c in ("C2_early0", "C2_early1", "C5_early0", "C5_early1"):
    try:
        d = json.load(open(f"gpurun_out/r2k/bench_{c}.json"))
        det = d["e2e"].get("detail") or {}
        print(c, "device", round(d["ms_per_step"], 3), "e2e ms", round(det.get("ms_per_step", 0), 3), "resident", round(det.get("ms_per_step_batch_resident", 0), 3),
              "e2e", round(d["e2e"]["value"]), "value", round(d["value"]))
    except Exception as e:
        print(c, "failed", e)
PY
