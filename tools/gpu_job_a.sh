#!/bin/bash
# round-2 GPU job A: full GPU test suite, bench of every BASELINE config, eager call/phase breakdown of C2
mkdir -p gpurun_out/r2a
O=gpurun_out/r2a
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_measured_gpu.py > $O/tests_old.log 2>&1; echo "old tests rc=$?" >> $O/rc.txt
timeout 900 python -m pytest tests/test_measured_gpu.py -q -s > $O/tests_measured.log 2>&1; echo "measured tests rc=$?" >> $O/rc.txt
timeout 600 python bench.py --config C2 --steps 10 --warmup 3 > $O/bench_C2.json 2> $O/bench_C2.err; echo "bench C2 rc=$?" >> $O/rc.txt
timeout 300 python bench.py --config C2 --skip-prob 0.5 --steps 10 --warmup 3 --skip-cpu --skip-library > $O/bench_C2_skip.json 2> $O/bench_C2_skip.err; echo "bench C2 skip rc=$?" >> $O/rc.txt
timeout 400 python bench.py --config C3 --steps 5 --warmup 3 --skip-cpu > $O/bench_C3.json 2> $O/bench_C3.err; echo "bench C3 rc=$?" >> $O/rc.txt
timeout 300 python bench.py --config C4 --steps 10 --warmup 3 --skip-cpu > $O/bench_C4.json 2> $O/bench_C4.err; echo "bench C4 rc=$?" >> $O/rc.txt
timeout 300 python bench.py --config C5 --steps 10 --warmup 3 --skip-cpu > $O/bench_C5.json 2> $O/bench_C5.err; echo "bench C5 rc=$?" >> $O/rc.txt
timeout 300 python tools/profile_step.py --steps 3 --calls > $O/calls_C2.txt 2>&1; echo "calls rc=$?" >> $O/rc.txt
cat $O/rc.txt
