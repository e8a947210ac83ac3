#!/bin/bash
# round-2 GPU job O: R=512 backward scan with 32 / 48-row slabs (weight half in tensor memory)
mkdir -p gpurun_out/r2o
O=gpurun_out/r2o
timeout 300 python -m pytest tests/test_lstm_scan_gpu.py -m gpu -q -x > $O/tests_scan.log 2>&1; echo "scan tests rc=$?" >> $O/rc.txt
tail -15 $O/tests_scan.log >> $O/rc.txt
echo "== R=512 scans (us/step): backward slab tiles by the wave model, then forced 1 / 2 / 3" > $O/scan.txt
R=512 timeout 120 python tools/bench_lstm_scan.py 2>&1 | grep "tf32=1\|Error\|error" >> $O/scan.txt
for mt in 1 2 3; do echo "-- P2PVG_LSTM512_BWD_MT=$mt" >> $O/scan.txt; R=512 P2PVG_LSTM512_BWD_MT=$mt timeout 120 python tools/bench_lstm_scan.py 2>&1 | grep "tf32=1\|Error\|error" >> $O/scan.txt; done
ts() { timeout 200 python tools/time_step.py "$@" 2>&1 | tail -1; }
echo "== C5 step: backward 16-row slabs vs the wave model" >> $O/scan.txt
P2PVG_LSTM512_BWD_MT=1 ts --steps 10 --backbone h36m_mlp --rnn 512 --seq 60 >> $O/scan.txt
ts --steps 10 --backbone h36m_mlp --rnn 512 --seq 60 >> $O/scan.txt
timeout 300 python -m pytest tests/test_measured_gpu.py tests/test_mlp_gpu.py -m gpu -q -x > $O/tests_b.log 2>&1; echo "measured + mlp tests rc=$?" >> $O/rc.txt
tail -3 $O/tests_b.log >> $O/rc.txt
cat $O/rc.txt; cat $O/scan.txt
