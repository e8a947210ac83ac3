#!/bin/bash
# round-2 GPU job B: new parity tests, fused BatchNorm statistics A/B, C2 bench line
mkdir -p gpurun_out/r2b
O=gpurun_out/r2b
timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_measured_gpu.py > $O/tests_all.log 2>&1; echo "all tests rc=$?" >> $O/rc.txt
timeout 900 python -m pytest tests/test_measured_gpu.py -q -s > $O/tests_measured.log 2>&1; echo "measured tests rc=$?" >> $O/rc.txt
for fuse in 0 1; do
  P2PVG_BN_FUSE=$fuse timeout 300 python tools/time_step.py --steps 10 >> $O/ab_fuse.txt 2>&1
  P2PVG_BN_FUSE=$fuse timeout 300 python tools/time_step.py --steps 5 --backbone vgg_64 --channels 3 --batch 128 >> $O/ab_fuse.txt 2>&1
done
timeout 600 python bench.py --config C2 --steps 10 --warmup 3 > $O/bench_C2.json 2> $O/bench_C2.err; echo "bench C2 rc=$?" >> $O/rc.txt
timeout 300 python bench.py --config C2 --skip-prob 0.5 --steps 20 --warmup 3 --skip-cpu --skip-library > $O/bench_C2_skip.json 2> $O/bench_C2_skip.err; echo "bench C2 skip rc=$?" >> $O/rc.txt
cat $O/rc.txt; cat $O/ab_fuse.txt
