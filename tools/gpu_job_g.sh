#!/bin/bash
# round-2 GPU job G: final validation (tests), A/B of the last kernels, final bench lines of every config
mkdir -p gpurun_out/r2g
O=gpurun_out/r2g
timeout 900 python -m pytest tests -m gpu -q -x > $O/tests_all.log 2>&1; echo "all gpu tests rc=$?" >> $O/rc.txt
tail -3 $O/tests_all.log >> $O/rc.txt
ts() { timeout 300 python tools/time_step.py "$@" 2>&1 | tail -1; }
echo "== C3 vgg_64: weight-gradient operand swap off/on, resident 3x3 weights off/on" >> $O/ab.txt
P2PVG_K1_SWAP=0 P2PVG_CONV_BRES=0 ts --steps 5 --backbone vgg_64 --channels 3 --batch 128 >> $O/ab.txt
P2PVG_K1_SWAP=1 P2PVG_CONV_BRES=0 ts --steps 5 --backbone vgg_64 --channels 3 --batch 128 >> $O/ab.txt
P2PVG_K1_SWAP=0 P2PVG_CONV_BRES=1 ts --steps 5 --backbone vgg_64 --channels 3 --batch 128 >> $O/ab.txt
P2PVG_K1_SWAP=1 P2PVG_CONV_BRES=1 ts --steps 5 --backbone vgg_64 --channels 3 --batch 128 >> $O/ab.txt
echo "== C4 dcgan_128: fused last layer (3 channels) off / on" >> $O/ab.txt
for v in 0 1; do P2PVG_FUSE_LAST=$v ts --steps 10 --backbone dcgan_128 --channels 3 --batch 64 >> $O/ab.txt; done
timeout 300 python bench.py --config C2 --steps 20 --warmup 5 > $O/bench_C2.json 2> $O/bench_C2.err; echo "bench C2 rc=$?" >> $O/rc.txt
timeout 300 python bench.py --config C2 --skip-prob 0.5 --steps 20 --warmup 3 --skip-cpu --skip-library > $O/bench_C2_skip.json 2> $O/bench_C2_skip.err; echo "bench C2 skip rc=$?" >> $O/rc.txt
timeout 400 python bench.py --config C3 --steps 10 --warmup 3 --skip-cpu > $O/bench_C3.json 2> $O/bench_C3.err; echo "bench C3 rc=$?" >> $O/rc.txt
timeout 300 python bench.py --config C4 --steps 10 --warmup 3 --skip-cpu > $O/bench_C4.json 2> $O/bench_C4.err; echo "bench C4 rc=$?" >> $O/rc.txt
timeout 300 python bench.py --config C5 --steps 10 --warmup 3 --skip-cpu > $O/bench_C5.json 2> $O/bench_C5.err; echo "bench C5 rc=$?" >> $O/rc.txt
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/rc.txt
cat $O/rc.txt; cat $O/ab.txt
