#!/bin/bash
# multi-GPU job: tools/gpu_job_mgpu.sh N   (DP check + bench lines at N GPUs of one box)
N=${1:-2}
mkdir -p gpurun_out/r2m$N
O=gpurun_out/r2m$N
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) "$@"; }
timeout 170 bash -c "$(declare -f run); N=$N; run tools/dp_check.py" > $O/dp_check.log 2>&1; echo "dp_check rc=$?" >> $O/rc.txt
tail -4 $O/dp_check.log >> $O/rc.txt
timeout 170 bash -c "$(declare -f run); N=$N; run bench.py --gpus $N --steps 20 --warmup 5 --skip-cpu" > $O/bench_C2_weak.json 2> $O/bench_C2_weak.err; echo "bench C2 weak rc=$?" >> $O/rc.txt
timeout 170 bash -c "$(declare -f run); N=$N; run bench.py --gpus $N --steps 20 --warmup 5 --skip-cpu --strong" > $O/bench_C2_strong.json 2> $O/bench_C2_strong.err; echo "bench C2 strong rc=$?" >> $O/rc.txt
if [ "$N" = "8" ]; then
  timeout 170 bash -c "$(declare -f run); N=$N; run bench.py --config C4 --gpus $N --steps 20 --warmup 5 --skip-cpu" > $O/bench_C4.json 2> $O/bench_C4.err; echo "bench C4 rc=$?" >> $O/rc.txt
  timeout 170 bash -c "$(declare -f run); N=$N; run bench.py --config C5 --gpus $N --steps 20 --warmup 5 --skip-cpu" > $O/bench_C5.json 2> $O/bench_C5.err; echo "bench C5 rc=$?" >> $O/rc.txt
  [ -z "$MGPU_MODE_B" ] || P2PVG_UPDATE_MODE=B timeout 170 bash -c "$(declare -f run); N=$N; run bench.py --gpus $N --steps 20 --warmup 5 --skip-cpu" > $O/bench_C2_weak_modeB.json 2> $O/bench_C2_weak_modeB.err; echo "bench C2 mode B rc=$?" >> $O/rc.txt
fi
cat $O/rc.txt
for f in $O/bench_*.json; do echo $f; python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("  ", d["n_gpus"], "GPUs", d["config"]["global_batch"], "global batch:", round(d["ms_per_step"], 3), "ms/step", round(d["value"]), "frames/s; e2e", round(d["e2e"]["value"]))
except Exception as e:
    print("  ERR", e)
PY
done
