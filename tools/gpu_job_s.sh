#!/bin/bash
# round-2 GPU job S: shared-memory-staged last layer + loss: tests, A/B on C2 / C4
mkdir -p gpurun_out/r2s
O=gpurun_out/r2s
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_step_gpu.py tests/test_measured_gpu.py tests/test_dropin_gpu.py -m gpu -q -x > $O/tests.log 2>&1; echo "tests rc=$?" >> $O/rc.txt
tail -3 $O/tests.log >> $O/rc.txt
ts() { timeout 200 python tools/time_step.py "$@" 2>&1 | tail -1; }
echo "== last layer + loss: direct gather (P2PVG_C1LOSS_TILED=0) / shared-memory bands (1): C2, then C4" > $O/ab.txt
for v in 0 1; do P2PVG_C1LOSS_TILED=$v ts --steps 20 >> $O/ab.txt; done
for v in 0 1; do P2PVG_C1LOSS_TILED=$v ts --steps 10 --backbone dcgan_128 --channels 3 --batch 64 >> $O/ab.txt; done
timeout 300 python tools/profile_step.py --steps 3 --calls --backbone dcgan_128 --channels 3 --batch 64 2>&1 | grep "convt_c1_loss" > $O/c1loss_C4.txt
timeout 300 python tools/profile_step.py --steps 3 --calls 2>&1 | grep "convt_c1_loss" > $O/c1loss_C2.txt
cat $O/rc.txt $O/ab.txt $O/c1loss_C4.txt $O/c1loss_C2.txt
