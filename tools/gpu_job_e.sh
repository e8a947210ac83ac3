#!/bin/bash
# round-2 GPU job E: full suite, remaining A/B runs, bench lines of every config
mkdir -p gpurun_out/r2e
O=gpurun_out/r2e
timeout 900 python -m pytest tests -m gpu -q > $O/tests_all.log 2>&1; echo "all gpu tests rc=$?" >> $O/rc.txt
tail -3 $O/tests_all.log >> $O/rc.txt
ts() { timeout 300 python tools/time_step.py "$@" 2>&1 | tail -1; }
echo "== fused ConvT phases (wide MMAs): max Cn 0 / 64 / 128 / 256 (C2)" >> $O/ab.txt
for v in 0 64 128 256; do P2PVG_CONVT4_MAX_CN=$v ts --steps 10 >> $O/ab.txt; done
echo "== fused last layer off / on (C2)" >> $O/ab.txt
for v in 0 1; do P2PVG_FUSE_LAST=$v ts --steps 10 >> $O/ab.txt; done
echo "== C4 fused ConvT phases 0 / 64 / 128" >> $O/ab.txt
for v in 0 64 128; do P2PVG_CONVT4_MAX_CN=$v ts --steps 10 --backbone dcgan_128 --channels 3 --batch 64 >> $O/ab.txt; done
timeout 600 python bench.py --config C2 --steps 20 --warmup 5 > $O/bench_C2.json 2> $O/bench_C2.err; echo "bench C2 rc=$?" >> $O/rc.txt
timeout 300 python bench.py --config C2 --skip-prob 0.5 --steps 20 --warmup 3 --skip-cpu --skip-library > $O/bench_C2_skip.json 2> $O/bench_C2_skip.err; echo "bench C2 skip rc=$?" >> $O/rc.txt
timeout 400 python bench.py --config C3 --steps 5 --warmup 3 --skip-cpu > $O/bench_C3.json 2> $O/bench_C3.err; echo "bench C3 rc=$?" >> $O/rc.txt
timeout 300 python bench.py --config C4 --steps 10 --warmup 3 --skip-cpu > $O/bench_C4.json 2> $O/bench_C4.err; echo "bench C4 rc=$?" >> $O/rc.txt
timeout 300 python bench.py --config C5 --steps 10 --warmup 3 --skip-cpu > $O/bench_C5.json 2> $O/bench_C5.err; echo "bench C5 rc=$?" >> $O/rc.txt
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref_C2.json 2> $O/bench_ref_C2.err; echo "bench ref rc=$?" >> $O/rc.txt
cat $O/rc.txt; cat $O/ab.txt
