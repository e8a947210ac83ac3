#!/bin/bash
# compute-sanitizer (synccheck, racecheck, memcheck) over the kernel-level GPU tests: mbarrier / TMEM / cluster-barrier kernels
mkdir -p gpurun_out/r2san
O=gpurun_out/r2san
SEL='tests/test_conv_gemm_gpu.py tests/test_lstm_scan_gpu.py::test_scan_fwd_bwd[5-3-64-True] tests/test_lstm_scan_gpu.py::test_scan_fwd_bwd[4-100-128-True] tests/test_lstm_scan_gpu.py::test_scan_fwd_bwd[6-40-512-True] tests/test_kernels_gpu.py'
for tool in synccheck racecheck memcheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python -m pytest $SEL -q -x -p no:cacheprovider > $O/$tool.log 2>&1
  echo "$tool rc=$?" >> $O/rc.txt
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" $O/$tool.log | tail -3 >> $O/rc.txt
done
cat $O/rc.txt
