#!/bin/bash
# round-2 GPU job J: full GPU test suite, smoke, final bench lines of every config (after the 48-row R=512 slabs, the dual
# frame-layout conversion and the forward() host-side reordering)
mkdir -p gpurun_out/r2j
O=gpurun_out/r2j
timeout 1200 python -m pytest tests -m gpu -q > $O/tests_all.log 2>&1; echo "all gpu tests rc=$?" >> $O/rc.txt
tail -4 $O/tests_all.log >> $O/rc.txt
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/rc.txt
timeout 400 python bench.py --config C2 --steps 20 --warmup 5 > $O/bench_C2.json 2> $O/bench_C2.err; echo "bench C2 rc=$?" >> $O/rc.txt
timeout 300 python bench.py --config C2 --skip-prob 0.5 --steps 20 --warmup 3 --skip-cpu --skip-library > $O/bench_C2_skip.json 2> $O/bench_C2_skip.err; echo "bench C2 skip rc=$?" >> $O/rc.txt
timeout 400 python bench.py --config C3 --steps 10 --warmup 3 --skip-cpu > $O/bench_C3.json 2> $O/bench_C3.err; echo "bench C3 rc=$?" >> $O/rc.txt
timeout 300 python bench.py --config C4 --steps 10 --warmup 3 --skip-cpu > $O/bench_C4.json 2> $O/bench_C4.err; echo "bench C4 rc=$?" >> $O/rc.txt
timeout 300 python bench.py --config C5 --steps 10 --warmup 3 --skip-cpu > $O/bench_C5.json 2> $O/bench_C5.err; echo "bench C5 rc=$?" >> $O/rc.txt
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref_C2.json 2> $O/bench_ref_C2.err; echo "bench reference arm rc=$?" >> $O/rc.txt
cat $O/rc.txt
python - <<'PY'
import json
for c in ("C2", "C2_skip", "C3", "C4", "C5"):
    try:
        d = json.load(open(f"gpurun_out/r2j/bench_{c}.json"))
        det = d["e2e"].get("detail") or {}
        print(c, round(d["ms_per_step"], 3), round(d["value"]), "e2e", round(d["e2e"]["value"]), "e2e ms", round(det.get("ms_per_step", 0), 2),
              "resident", round(det.get("ms_per_step_batch_resident", 0), 2), "roof", round(d["roofline"]["frac"], 3),
              "lstm", round((d.get("roofline_lstm") or {}).get("frac", 0), 3), d["clocks"]["sm_mhz"], d["clocks"]["reasons"],
              "lib", (d.get("library_baseline") or {}).get("ms_per_step"), "cpu", (d.get("cpu_baseline") or {}).get("value"))
    except Exception as e:
        print(c, "failed", e)
PY
