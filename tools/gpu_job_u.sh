#!/bin/bash
# 8-GPU check of the input-path scheduling (slot refill after the step when data parallel): C2 weak only
N=8
mkdir -p gpurun_out/r2u
O=gpurun_out/r2u
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) "$@"; }
timeout 170 bash -c "$(declare -f run); N=$N; run bench.py --gpus $N --steps 20 --warmup 5 --skip-cpu --skip-library --skip-phases" > $O/bench_C2_weak.json 2> $O/bench_C2_weak.err; echo "rc=$?" >> $O/rc.txt
cat $O/rc.txt
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2u/bench_C2_weak.json").read().strip().splitlines()[-1])
det = d["e2e"]["detail"]
print("N=8 C2 weak: device", round(d["ms_per_step"], 3), "e2e", round(det["ms_per_step"], 3), "resident", round(det["ms_per_step_batch_resident"], 3), "value", round(d["value"]), "e2e", round(d["e2e"]["value"]))
PY
