#!/bin/bash
# round-2 GPU job H: cluster occupancy, R=256 slab-size A/B, per-call profiles of C4 / C5
mkdir -p gpurun_out/r2h
O=gpurun_out/r2h
echo "== R=256 scans, default slab rows (32 above B=128)" > $O/scan.txt
R=256 timeout 200 python tools/bench_lstm_scan.py 2>&1 | grep -v "tf32=0" >> $O/scan.txt
echo "== R=256 scans, 16-row slabs at every batch size (P2PVG_LSTM_MT2_ABOVE=100000)" >> $O/scan.txt
R=256 P2PVG_LSTM_MT2_ABOVE=100000 timeout 200 python tools/bench_lstm_scan.py 2>&1 | grep -v "tf32=0" >> $O/scan.txt
ts() { timeout 300 python tools/time_step.py "$@" 2>&1 | tail -1; }
echo "== C2 step with 16-row slabs everywhere" >> $O/scan.txt
ts --steps 10 >> $O/scan.txt
P2PVG_LSTM_MT2_ABOVE=100000 ts --steps 10 >> $O/scan.txt
timeout 300 python tools/profile_step.py --steps 3 --calls --backbone dcgan_128 --channels 3 --batch 64 > $O/calls_C4.txt 2>&1
timeout 300 python tools/profile_step.py --steps 3 --calls --backbone h36m_mlp --rnn 512 --seq 60 > $O/calls_C5.txt 2>&1
cat $O/scan.txt
