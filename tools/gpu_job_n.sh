#!/bin/bash
# round-2 GPU job N: H2D rate of the box, e2e with 1 / 4 copy streams
mkdir -p gpurun_out/r2n
O=gpurun_out/r2n
timeout 300 python -m pytest tests/test_data_gpu.py -m gpu -q -x > $O/tests.log 2>&1; echo "data tests rc=$?" >> $O/rc.txt
tail -3 $O/tests.log >> $O/rc.txt
for n in 1 4; do
  P2PVG_BENCH_COPY_STREAMS=$n timeout 300 python bench.py --config C2 --steps 20 --warmup 5 --skip-cpu --skip-library --skip-phases > $O/bench_C2_cs$n.json 2> $O/bench_C2_cs$n.err; echo "bench C2 copy_streams=$n rc=$?" >> $O/rc.txt
done
cat $O/rc.txt
python - <<'PY'
import json
for n in (1, 4):
    try:
        d = json.load(open(f"gpurun_out/r2n/bench_C2_cs{n}.json"))
        det = d["e2e"]["detail"]
        print("copy_streams", n, "device", round(d["ms_per_step"], 3), "e2e ms", round(det["ms_per_step"], 3), "resident", round(det["ms_per_step_batch_resident"], 3),
              "h2d idle GB/s", det["h2d_gbps_gpu_idle"], "needed", round(det["h2d_gbps_needed"], 2))
    except Exception as e:
        print(n, "failed", e)
PY
nvidia-smi topo -m 2>/dev/null | head -8; nvidia-smi --query-gpu=pcie.link.gen.current,pcie.link.width.current,pcie.link.gen.max --format=csv
