#!/bin/bash
# backward slab-size invariance (tensor-memory scan vs shared-memory scan, bitwise) + synccheck / racecheck over the other late kernels
mkdir -p gpurun_out/r2w
O=gpurun_out/r2w
timeout 120 python -m pytest tests/test_lstm_scan_gpu.py -m gpu -q -x > $O/tests_scan.log 2>&1; echo "scan tests (incl. backward slab invariance) rc=$?" >> $O/rc.txt
tail -2 $O/tests_scan.log >> $O/rc.txt
SEL='tests/test_kernels_gpu.py::test_permute4_flat_cast tests/test_kernels_gpu.py::test_nchw_to_nhwc_dual tests/test_kernels_gpu.py::test_lstm_pointwise_and_reparam tests/test_kernels_gpu.py::test_concat_gather_align_colsum_act tests/test_kernels_gpu.py::test_sigmoid_mse_finalize_adam tests/test_lstm_scan_gpu.py::test_scan512_slab_size_invariance'
for tool in synccheck racecheck; do
  timeout 150 compute-sanitizer --tool $tool --print-limit 10 python -m pytest $SEL -q -x -p no:cacheprovider > $O/$tool.log 2>&1
  echo "$tool (late kernels without the tensor-memory scan) rc=$?" >> $O/rc.txt
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" $O/$tool.log | tail -2 >> $O/rc.txt
done
cat $O/rc.txt
