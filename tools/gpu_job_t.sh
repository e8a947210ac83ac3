#!/bin/bash
# 2-GPU A/B: end-to-end with / without the early loss hand-over and early slot release
N=2
mkdir -p gpurun_out/r2t
O=gpurun_out/r2t
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) "$@"; }
for v in 0 1; do
  P2PVG_EARLY_LOSS=$v timeout 170 bash -c "$(declare -f run); N=$N; run bench.py --gpus $N --steps 20 --warmup 5 --skip-cpu --skip-library --skip-phases" > $O/bench_early$v.json 2> $O/bench_early$v.err; echo "early=$v rc=$?" >> $O/rc.txt
done
cat $O/rc.txt
python - <<'PY'
import json
for v in (0, 1):
    try:
        d = json.loads(open(f"gpurun_out/r2t/bench_early{v}.json").read().strip().splitlines()[-1])
        det = d["e2e"]["detail"]
        print("early", v, "device", round(d["ms_per_step"], 3), "e2e", round(det["ms_per_step"], 3), "resident", round(det["ms_per_step_batch_resident"], 3), "enqueue", round(det["host_enqueue_ms_per_step"], 3))
    except Exception as e:
        print(v, "failed", e)
PY
