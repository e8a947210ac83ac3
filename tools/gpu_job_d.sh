#!/bin/bash
# round-2 GPU job D: R=512 cluster-16 LSTM scans (tests + C5), convt4 on C4, bench lines
mkdir -p gpurun_out/r2d
O=gpurun_out/r2d
timeout 600 python -m pytest tests/test_lstm_scan_gpu.py -q -x > $O/tests_lstm.log 2>&1; echo "lstm tests rc=$?" >> $O/rc.txt
timeout 900 python -m pytest tests/test_measured_gpu.py tests/test_mlp_gpu.py -q -s > $O/tests_measured.log 2>&1; echo "measured+mlp tests rc=$?" >> $O/rc.txt
ts() { timeout 300 python tools/time_step.py "$@" 2>&1 | tail -1; }
echo "== C5 cluster-16 scans off / on" >> $O/ab.txt
P2PVG_FUSED_SCAN=0 ts --steps 10 --backbone h36m_mlp --rnn 512 --seq 60 >> $O/ab.txt
ts --steps 10 --backbone h36m_mlp --rnn 512 --seq 60 >> $O/ab.txt
echo "== scan kernels alone, R=512 (us/step)" >> $O/ab.txt
R=512 timeout 300 python tools/bench_lstm_scan.py 2>&1 | grep "tf32=1" >> $O/ab.txt
echo "== C4 convt4 max cn 0 / 64" >> $O/ab.txt
for v in 0 64; do P2PVG_CONVT4_MAX_CN=$v ts --steps 10 --backbone dcgan_128 --channels 3 --batch 64 >> $O/ab.txt; done
timeout 300 python bench.py --config C5 --steps 10 --warmup 3 --skip-cpu > $O/bench_C5.json 2> $O/bench_C5.err; echo "bench C5 rc=$?" >> $O/rc.txt
timeout 300 python bench.py --config C2 --steps 10 --warmup 3 --skip-cpu --skip-library > $O/bench_C2.json 2> $O/bench_C2.err; echo "bench C2 rc=$?" >> $O/rc.txt
cat $O/rc.txt; cat $O/ab.txt; tail -5 $O/tests_lstm.log
