#!/bin/bash
# sanitizer diagnostics for the tensor-memory backward scan (synccheck report "barrier missing init at shared 0x0")
mkdir -p gpurun_out/r2x
O=gpurun_out/r2x
T1='tests/test_lstm_scan_gpu.py::test_scan_fwd_bwd[6-40-512-True]'
T2='tests/test_lstm_scan_gpu.py::test_scan_fwd_bwd[5-128-512-True]'
TG='tests/test_conv_gemm_gpu.py'
timeout 120 compute-sanitizer --tool synccheck --print-limit 2 --show-backtrace no python -m pytest "$T1" -q -x -p no:cacheprovider > $O/sync_mt1.log 2>&1; echo "synccheck 16-row slabs rc=$?" >> $O/rc.txt
grep -E "ERROR SUMMARY|passed|failed" $O/sync_mt1.log | tail -2 >> $O/rc.txt
timeout 120 compute-sanitizer --tool synccheck --print-limit 2 --show-backtrace no python -m pytest "$T2" -q -x -p no:cacheprovider > $O/sync_mt2.log 2>&1; echo "synccheck 32-row slabs (tensor memory) rc=$?" >> $O/rc.txt
grep -E "ERROR SUMMARY|passed|failed" $O/sync_mt2.log | tail -2 >> $O/rc.txt
P2PVG_LSTM512_BWD_MT=1 timeout 120 compute-sanitizer --tool synccheck --print-limit 2 --show-backtrace no python -m pytest "$T2" -q -x -p no:cacheprovider > $O/sync_mt2_forced1.log 2>&1; echo "synccheck same test, 16-row slabs forced rc=$?" >> $O/rc.txt
grep -E "ERROR SUMMARY|passed|failed" $O/sync_mt2_forced1.log | tail -2 >> $O/rc.txt
timeout 150 compute-sanitizer --tool racecheck --print-limit 5 python -m pytest "$T2" -q -x -p no:cacheprovider > $O/race_mt2.log 2>&1; echo "racecheck 32-row slabs rc=$?" >> $O/rc.txt
grep -E "RACECHECK SUMMARY|passed|failed|Error" $O/race_mt2.log | tail -3 >> $O/rc.txt
cat $O/rc.txt; grep -v "Host Frame" $O/sync_mt2.log | head -20
