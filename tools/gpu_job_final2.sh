#!/bin/bash
# round-2 last GPU job: the driver's own sequence on the final commit (GPU tests, smoke, default bench invocation + reference arm),
# then compute-sanitizer over the tests of the kernels added late in the round
mkdir -p gpurun_out/r2y
O=gpurun_out/r2y
timeout 900 python -m pytest tests -x -q -m gpu > $O/tests_all.log 2>&1; echo "pytest -m gpu rc=$?" >> $O/rc.txt
tail -3 $O/tests_all.log >> $O/rc.txt
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/rc.txt
timeout 300 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; echo "bench --impl reference rc=$?" >> $O/rc.txt
timeout 400 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" >> $O/rc.txt
SEL='tests/test_lstm_scan_gpu.py::test_scan_fwd_bwd[5-128-512-True] tests/test_lstm_scan_gpu.py::test_scan_fwd_bwd[7-250-512-True] tests/test_lstm_scan_gpu.py::test_scan_fwd_bwd[6-40-512-True] tests/test_kernels_gpu.py::test_permute4_flat_cast tests/test_kernels_gpu.py::test_nchw_to_nhwc_dual tests/test_kernels_gpu.py::test_lstm_pointwise_and_reparam tests/test_kernels_gpu.py::test_concat_gather_align_colsum_act tests/test_kernels_gpu.py::test_sigmoid_mse_finalize_adam'
for tool in memcheck synccheck racecheck; do
  timeout 240 compute-sanitizer --tool $tool --print-limit 20 python -m pytest $SEL -q -x -p no:cacheprovider > $O/$tool.log 2>&1
  echo "$tool rc=$?" >> $O/rc.txt
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" $O/$tool.log | tail -3 >> $O/rc.txt
done
cat $O/rc.txt
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2y/bench.json"))
print("C2", round(d["ms_per_step"], 3), round(d["value"]), "e2e", round(d["e2e"]["value"]), "roof", round(d["roofline"]["frac"], 3), d["clocks"], "launches", d["gpu_launches"])
r = json.load(open("gpurun_out/r2y/bench_ref.json"))
print("ref", r["value"], r["impl"], r["cpu_baseline"]["cores"])
PY
