#!/bin/bash
# round-2 GPU job C: concurrency lanes + selective BatchNorm-statistics fusion A/B, tests
mkdir -p gpurun_out/r2c
O=gpurun_out/r2c
timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_measured_gpu.py > $O/tests_all.log 2>&1; echo "all tests rc=$?" >> $O/rc.txt
timeout 900 python -m pytest tests/test_measured_gpu.py -q -s > $O/tests_measured.log 2>&1; echo "measured tests rc=$?" >> $O/rc.txt
ts() { timeout 300 python tools/time_step.py "$@" 2>&1 | tail -1; }
echo "== concurrency off / on (C2)" >> $O/ab.txt
P2PVG_CONCURRENT=0 ts --steps 10 >> $O/ab.txt
P2PVG_CONCURRENT=1 ts --steps 10 >> $O/ab.txt
echo "== concurrency off / on (C5: h36m R=512 B=256 T=60)" >> $O/ab.txt
P2PVG_CONCURRENT=0 ts --steps 10 --backbone h36m_mlp --rnn 512 --seq 60 >> $O/ab.txt
P2PVG_CONCURRENT=1 ts --steps 10 --backbone h36m_mlp --rnn 512 --seq 60 >> $O/ab.txt
echo "== BN statistics fusion: off, min 0 (always), 131072, 262144 (C2)" >> $O/ab.txt
P2PVG_BN_FUSE=0 ts --steps 10 >> $O/ab.txt
P2PVG_BN_FUSE_MIN=0 ts --steps 10 >> $O/ab.txt
P2PVG_BN_FUSE_MIN=131072 ts --steps 10 >> $O/ab.txt
P2PVG_BN_FUSE_MIN=262144 ts --steps 10 >> $O/ab.txt
echo "== same for vgg_64 (C3): off, 0, 131072, 262144" >> $O/ab.txt
for v in off 0 131072 262144; do
  if [ $v = off ]; then P2PVG_BN_FUSE=0 ts --steps 5 --backbone vgg_64 --channels 3 --batch 128 >> $O/ab.txt; else P2PVG_BN_FUSE_MIN=$v ts --steps 5 --backbone vgg_64 --channels 3 --batch 128 >> $O/ab.txt; fi
done
echo "== fused ConvT phases: max Cn 0 (off), 64, 128 (default), 256 (C2)" >> $O/ab.txt
for v in 0 64 128 256; do P2PVG_CONVT4_MAX_CN=$v ts --steps 10 >> $O/ab.txt; done
echo "== fused ConvT phases off / on (C4 dcgan_128 B=64)" >> $O/ab.txt
for v in 0 128; do P2PVG_CONVT4_MAX_CN=$v ts --steps 10 --backbone dcgan_128 --channels 3 --batch 64 >> $O/ab.txt; done
echo "== overlap heavy forks (P2PVG_OVERLAP=1) (C2)" >> $O/ab.txt
P2PVG_OVERLAP=1 ts --steps 10 >> $O/ab.txt
timeout 300 python bench.py --config C2 --steps 10 --warmup 3 --skip-cpu --skip-library > $O/bench_C2.json 2> $O/bench_C2.err; echo "bench C2 rc=$?" >> $O/rc.txt
cat $O/rc.txt; cat $O/ab.txt
