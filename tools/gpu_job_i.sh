#!/bin/bash
# round-2 GPU job I: 48-row slabs of the R=512 forward scan, dual frame-layout conversion, e2e host/H2D split
mkdir -p gpurun_out/r2i
O=gpurun_out/r2i
timeout 600 python -m pytest tests/test_lstm_scan_gpu.py tests/test_kernels_gpu.py -m gpu -q -x > $O/tests_a.log 2>&1; echo "scan+kernel tests rc=$?" >> $O/rc.txt
tail -3 $O/tests_a.log >> $O/rc.txt
timeout 600 python -m pytest tests/test_measured_gpu.py tests/test_step_gpu.py -m gpu -q -x > $O/tests_b.log 2>&1; echo "measured+step tests rc=$?" >> $O/rc.txt
tail -3 $O/tests_b.log >> $O/rc.txt
echo "== R=512 scans (us/step): slab tiles picked by the wave model, then forced 1 / 2 / 3" > $O/scan.txt
R=512 timeout 200 python tools/bench_lstm_scan.py 2>&1 | grep -v "tf32=0" >> $O/scan.txt
for mt in 1 2 3; do echo "-- P2PVG_LSTM512_MT=$mt" >> $O/scan.txt; R=512 P2PVG_LSTM512_MT=$mt timeout 200 python tools/bench_lstm_scan.py 2>&1 | grep "tf32=1" >> $O/scan.txt; done
ts() { timeout 300 python tools/time_step.py "$@" 2>&1 | tail -1; }
echo "== C5 step: forced 32-row slabs (round-2 default so far) vs the wave model" >> $O/scan.txt
P2PVG_LSTM512_MT=2 ts --steps 10 --backbone h36m_mlp --rnn 512 --seq 60 >> $O/scan.txt
ts --steps 10 --backbone h36m_mlp --rnn 512 --seq 60 >> $O/scan.txt
timeout 300 python bench.py --config C2 --steps 20 --warmup 5 --skip-library > $O/bench_C2.json 2> $O/bench_C2.err; echo "bench C2 rc=$?" >> $O/rc.txt
timeout 300 python bench.py --config C4 --steps 10 --warmup 3 --skip-cpu --skip-library > $O/bench_C4.json 2> $O/bench_C4.err; echo "bench C4 rc=$?" >> $O/rc.txt
timeout 300 python bench.py --config C5 --steps 10 --warmup 3 --skip-cpu --skip-library > $O/bench_C5.json 2> $O/bench_C5.err; echo "bench C5 rc=$?" >> $O/rc.txt
cat $O/rc.txt; cat $O/scan.txt
python - <<'PY'
import json
for c in ("C2", "C4", "C5"):
    try:
        d = json.load(open(f"gpurun_out/r2i/bench_{c}.json"))
        print(c, round(d["ms_per_step"], 3), round(d["value"]), "e2e", round(d["e2e"]["value"]), d["e2e"].get("detail"))
    except Exception as e:
        print(c, "failed", e)
PY
