#!/bin/bash
# 4-GPU line (the one world size of the driver's scaling run not exercised yet this round)
N=4
mkdir -p gpurun_out/r2v
O=gpurun_out/r2v
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) "$@"; }
timeout 100 bash -c "$(declare -f run); N=$N; run bench.py --gpus $N --steps 10 --warmup 3 --skip-library --skip-phases" > $O/bench_C2_weak.json 2> $O/bench_C2_weak.err; echo "rc=$?" >> $O/rc.txt
cat $O/rc.txt
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2v/bench_C2_weak.json").read().strip().splitlines()[-1])
print("N=4 C2 weak: device", round(d["ms_per_step"], 3), "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "cpu_baseline", d.get("cpu_baseline"))
PY
