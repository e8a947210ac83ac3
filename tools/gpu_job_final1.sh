#!/bin/bash
# round-2 final 1-GPU job: full GPU test suite, smoke, bench lines of every config + reference arm, fresh ncu evidence
mkdir -p gpurun_out/r2z
O=gpurun_out/r2z
timeout 1200 python -m pytest tests -m gpu -q > $O/tests_all.log 2>&1; echo "all gpu tests rc=$?" >> $O/rc.txt
tail -4 $O/tests_all.log >> $O/rc.txt
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/rc.txt
timeout 400 python bench.py --config C2 --steps 20 --warmup 5 > $O/bench_C2.json 2> $O/bench_C2.err; echo "bench C2 rc=$?" >> $O/rc.txt
timeout 300 python bench.py --config C2 --skip-prob 0.5 --steps 20 --warmup 3 --skip-cpu --skip-library > $O/bench_C2_skip.json 2> $O/bench_C2_skip.err; echo "bench C2 skip rc=$?" >> $O/rc.txt
timeout 400 python bench.py --config C3 --steps 10 --warmup 3 --skip-cpu > $O/bench_C3.json 2> $O/bench_C3.err; echo "bench C3 rc=$?" >> $O/rc.txt
timeout 300 python bench.py --config C4 --steps 10 --warmup 3 --skip-cpu > $O/bench_C4.json 2> $O/bench_C4.err; echo "bench C4 rc=$?" >> $O/rc.txt
timeout 300 python bench.py --config C5 --steps 10 --warmup 3 --skip-cpu > $O/bench_C5.json 2> $O/bench_C5.err; echo "bench C5 rc=$?" >> $O/rc.txt
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref_C2.json 2> $O/bench_ref_C2.err; echo "bench reference arm rc=$?" >> $O/rc.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_C2.csv python tools/profile_step.py --steps 2 > $O/launches_C2.out 2>&1
python tools/summarize_launches.py $O/launches_C2.csv 2 > $O/launches_C2.txt 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"lstm_cl" -c 4 -o $O/lstm_full python tools/bench_lstm_scan_once.py > $O/lstm_full.out 2>&1
timeout 300 python tools/profile_step.py --steps 3 --calls > $O/calls_C2.txt 2>&1
timeout 300 python tools/profile_step.py --steps 3 --calls --backbone h36m_mlp --rnn 512 --seq 60 > $O/calls_C5.txt 2>&1
timeout 300 python tools/profile_step.py --steps 3 --calls --backbone dcgan_128 --channels 3 --batch 64 > $O/calls_C4.txt 2>&1
R=512 timeout 200 python tools/bench_lstm_scan.py 2>&1 | grep "tf32=1" > $O/scan_R512.txt
cat $O/rc.txt
python - <<'PY'
import json
for c in ("C2", "C2_skip", "C3", "C4", "C5"):
    try:
        d = json.load(open(f"gpurun_out/r2z/bench_{c}.json"))
        det = d["e2e"].get("detail") or {}
        print(c, round(d["ms_per_step"], 3), round(d["value"]), "e2e", round(d["e2e"]["value"]), "e2e ms", round(det.get("ms_per_step", 0), 2),
              "roof", round(d["roofline"]["frac"], 3), "lstm", round((d.get("roofline_lstm") or {}).get("frac", 0), 3), d["clocks"]["sm_mhz"], d["clocks"]["reasons"],
              "h2d", det.get("h2d_gbps_gpu_idle"), "lib", (d.get("library_baseline") or {}).get("ms_per_step"), "cpu", (d.get("cpu_baseline") or {}).get("value"))
    except Exception as e:
        print(c, "failed", e)
PY
