#!/bin/bash
# round-2 GPU job F: validate vectorised vgg kernels + bf16 addend (tests), A/B, C3 / C2 / C4 bench lines
mkdir -p gpurun_out/r2f
O=gpurun_out/r2f
timeout 900 python -m pytest tests -m gpu -q -x > $O/tests_all.log 2>&1; echo "all gpu tests rc=$?" >> $O/rc.txt
tail -3 $O/tests_all.log >> $O/rc.txt
ts() { timeout 300 python tools/time_step.py "$@" 2>&1 | tail -1; }
echo "== bf16 addend off / on (C2)" >> $O/ab.txt
for v in 0 1; do P2PVG_ADDEND_BF16=$v ts --steps 10 >> $O/ab.txt; done
echo "== bf16 addend off / on (C3 vgg_64 B=128), vectorised data movement in both" >> $O/ab.txt
for v in 0 1; do P2PVG_ADDEND_BF16=$v ts --steps 5 --backbone vgg_64 --channels 3 --batch 128 >> $O/ab.txt; done
timeout 400 python bench.py --config C3 --steps 5 --warmup 3 --skip-cpu > $O/bench_C3.json 2> $O/bench_C3.err; echo "bench C3 rc=$?" >> $O/rc.txt
timeout 300 python bench.py --config C2 --steps 20 --warmup 5 > $O/bench_C2.json 2> $O/bench_C2.err; echo "bench C2 rc=$?" >> $O/rc.txt
timeout 300 python bench.py --config C4 --steps 10 --warmup 3 --skip-cpu > $O/bench_C4.json 2> $O/bench_C4.err; echo "bench C4 rc=$?" >> $O/rc.txt
timeout 400 python tools/profile_step.py --steps 3 --calls --backbone vgg_64 --channels 3 --batch 128 > $O/calls_C3.txt 2>&1
cat $O/rc.txt; cat $O/ab.txt
