#!/usr/bin/env python
"""Benchmark of the p2pvg training hot path (BASELINE.json): frames/s of one P2PModel.forward train step
(forward + both backwards + five Adam updates) on synthetic MovingMNIST-shaped batches.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One JSON line on stdout (rank 0).  `value` = device-timed throughput with the batch resident in HBM;
`e2e` = the same through the public drop-in API (p2pvg_b200.data.DevicePrefetcher feeding
models.p2p_model.P2PModel.__call__): every step's batch is copied from pinned host memory inside the timed
region (side stream, overlapping the previous step) and the four loss scalars are read back every step; `roofline` = the tcgen05 GEMM
kernel (executed FLOPs / CUDA-event time of its launches, measured in an instrumented pass inside this
process); `cpu_baseline` = the CPU oracle (port of the reference path) on the host cores.
`--impl reference` times the reference's CPU path (the oracle port; the reference itself is pure PyTorch
and cannot travel to the GPU box) on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CFG = dict(g_dim=128, z_dim=10, rnn_size=256, channels=1, image_width=64, predictor_rnn_layers=2, posterior_rnn_layers=1,
           prior_rnn_layers=1)
METRIC = "frames/sec (train step, device-timed) MovingMNIST 64x64 seq30"
# algorithmic FLOPs per sequence per train step for dcgan_64, T=30
W_FLOP_PER_SEQ = 54.9e9


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(tflops=d["bf16_tflops_sustained"], hbm=d["hbm_gbs"], which="measured (MEASURED_PEAKS.json, sustained)")
    return dict(tflops=1590.0, hbm=6650.0, which="fallback (B200_PROFILING.md)")


class ClockSampler:
    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return None
        self.proc.terminate()
        sm, mx, reasons = [], 0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for t, line in self.rows:
            if t < t0 or t > t1:
                continue
            f = [c.strip() for c in line.split(",")]
            try:
                sm.append(float(f[0]))
                mx = max(mx, float(f[1]))
            except Exception:
                continue
            for nm, v in zip(names, f[2:]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=[], samples=0)
        return dict(sm_mhz=float(np.median(sm)), sm_max_mhz=mx, reasons=sorted(reasons), samples=len(sm))


def log(msg):
    print(f"[bench {time.strftime('%H:%M:%S')}] {msg}", file=sys.stderr, flush=True)


def host_threads():
    """Threads for the CPU reference: the reference path is ~93k small ATen ops per step; beyond
    a few dozen threads the per-op fork/join cost dominates, so cap at 32 (P2PVG_CPU_THREADS overrides)."""
    return int(os.environ.get("P2PVG_CPU_THREADS", min(os.cpu_count() or 1, 32)))


def make_opt(backbone, batch):
    return types.SimpleNamespace(dataset="mnist", backbone_net=backbone, lr=1e-3, beta1=0.9, beta=1e-4, weight_cpc=100.0,
                                 weight_align=0.5, skip_prob=0.0, n_past=1, last_frame_skip=False, batch_size=batch)


def cpu_reference_steps(T, B, steps, warmup, threads):
    """Times the CPU oracle (port of reference models/p2p_model.py:185-271, Mode A) on a [T,B] batch."""
    from oracle import p2p_oracle as O
    torch.set_num_threads(threads)
    state = O.build_state(CFG, seed=1)
    adam = {m: O.new_adam_state(state[m]) for m in O.MODULES}
    opt = O.default_opt(batch_size=B)
    gen = torch.Generator().manual_seed(1234)
    times = []
    for it in range(warmup + steps):
        x = torch.rand(T, B, 1, 64, 64, generator=gen)
        probs = np.random.RandomState(it).uniform(0, 1, T - 1)
        eps = O.draw_eps(T - 1, B, CFG["z_dim"], seed=it)
        t0 = time.perf_counter()
        O.train_step(state, adam, x, opt, 64, eps, probs, mode="A")
        dt = time.perf_counter() - t0
        log(f'cpu reference step {it}: {dt:.2f} s')
        if it >= warmup:
            times.append(dt)
    return times


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    T, B = args.seq, args.ref_batch
    threads = host_threads()
    times = cpu_reference_steps(T, B, args.steps, args.warmup, threads)
    tot = sum(times)
    val = T * B * len(times) / tot
    line = dict(metric=METRIC, value=val, unit="frames/s", n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=1e3 * tot / len(times), higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32",
                data="synthetic", impl="reference",
                config=dict(workload=f"mnist dcgan_64 seq_len {T}: sample of {B} sequences per step of the batch-256 workload",
                            skip_prob=0.0, parallelism="cpu"),
                cpu_baseline=dict(value=val, unit="frames/s", cores=threads, kind="port",
                                  sample=f"{len(times)} steps of T={T},B={B} (oracle/p2p_oracle.py, Mode A, torch CPU fp32)"),
                e2e=dict(value=val, unit="frames/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--batch", type=int, default=256, help="sequences per GPU")
    ap.add_argument("--seq", type=int, default=30)
    ap.add_argument("--ref-batch", type=int, default=16)
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--skip-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    from p2pvg_b200.models import dcgan_64
    from p2pvg_b200.models.p2p_model import P2PModel

    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    T, B = args.seq, args.batch
    os.environ["P2PVG_PRECISION"] = args.precision
    # multi-GPU: the NCCL all-reduces are captured into the step graph as well (P2PVG_DP_GRAPH=0 disables)
    dp_graph = os.environ.get("P2PVG_DP_GRAPH", "1") != "0"
    os.environ["P2PVG_GRAPH"] = "0" if (args.no_graph or (world > 1 and not dp_graph)) else "1"
    torch.manual_seed(1)
    np.random.seed(0)
    model = P2PModel(B, 1, 128, 10, 256, 1, 1, 2, opt=make_opt(dcgan_64, B)).cuda()
    model.train()
    gen = torch.Generator().manual_seed(1234 + rank)
    x_host = torch.rand(T, B, 1, 64, 64, generator=gen).pin_memory()
    x_dev = x_host.to(dev)
    eng = model.engine(64)
    if world > 1:
        eng.dist = (dist, None, world)
    K = eng.K
    use_graph = model.use_graph

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    log('model built; warm-up')
    # ---- device-timed, batch resident in HBM --------------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        eng.step(x_dev, use_graph=use_graph, return_device=True)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    n0 = K.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    tw0 = time.time()
    e0.record()
    for _ in range(args.steps):
        out = eng.step(x_dev, use_graph=use_graph, return_device=True)
    e1.record()
    barrier()
    tw1 = time.time()
    launches = K.launches - n0
    log(f'timed region done: {e0.elapsed_time(e1) / args.steps:.2f} ms/step')
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_step = ms.item() / args.steps
    clocks = sampler.stop(tw0, tw1) if rank == 0 else None
    value = T * B * world / (ms_step * 1e-3)

    # ---- end to end through the public API ----------------------------------------------------------
    # public API = DevicePrefetcher (p2pvg_b200/data.py) feeding P2PModel.__call__: every step's batch is copied from pinned
    # host memory inside the timed region (on a side stream, overlapping the previous step) and the four scalars are read back
    from p2pvg_b200.data import DevicePrefetcher

    def host_batches(n):
        for _ in range(n):
            yield x_host

    for xb in DevicePrefetcher(host_batches(2), dev):
        model(xb, 0, T - 1)
    barrier()
    e0.record()
    pf = DevicePrefetcher(host_batches(args.steps), dev)
    for xb in pf:
        losses = model(xb, 0, T - 1)
        pf.release()
    e1.record()
    barrier()
    ms2 = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
    e2e_val = T * B * world / (ms2.item() / args.steps * 1e-3)

    log(f'e2e done: {ms2.item() / args.steps:.2f} ms/step')
    # ---- roofline of the dominant kernel: instrumented eager pass ----------------------------------
    roof = None
    if rank == 0 and args.precision == "bf16":
        rec = []
        orig = K.gemm

        def timed_gemm(A, Bm, C, M, N, Kd, **kw):
            if A.dtype == torch.bfloat16:
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                orig(A, Bm, C, M, N, Kd, **kw)
                b.record()
                # block-diagonal weights (thin first / last layers): 3/4 of the executed MACs multiply structural zeros
                useful = 0.25 if Bm.data_ptr() in {t.data_ptr() for k_, t in eng._packed.items() if ".bd" in k_} else 1.0
                rec.append((a, b, useful * 2.0 * M * N * Kd, (M, N, Kd, bool(kw.get("a_mn")), bool(kw.get("b_mn")), str(C.dtype))))
            else:
                orig(A, Bm, C, M, N, Kd, **kw)

        orig_conv = K.conv_gemm

        def timed_conv(kind, a_, b_, c_, N_, H_, W_, Ck, Cn, Cm=0, **kw):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            orig_conv(kind, a_, b_, c_, N_, H_, W_, Ck, Cn, Cm=Cm, **kw)
            b.record()
            pix = float(N_) * H_ * W_   # small-map pixels; every kind does 16 taps worth of MACs per (pixel, Ck|Cm, Cn)
            fl_ = 2.0 * pix * 16 * (Cm if kind == 1 else Ck) * Cn
            rec.append((a, b, fl_, (f"conv_gemm kind{kind}", N_, H_, W_, Ck, Cn, Cm)))

        K.gemm = timed_gemm
        K.conv_gemm = timed_conv
        saved_dist, eng.dist = eng.dist, None  # rank-0-only pass: no collectives
        eng.step(x_dev, use_graph=False, return_device=True)
        torch.cuda.synchronize()
        eng.dist = saved_dist
        K.gemm = orig
        K.conv_gemm = orig_conv
        tms = sum(r[0].elapsed_time(r[1]) for r in rec)
        fl = sum(r[2] for r in rec)
        if os.environ.get("P2PVG_DUMP_GEMMS"):
            rows = [dict(shape=r[3], ms=r[0].elapsed_time(r[1]), tflops=r[2] / (r[0].elapsed_time(r[1]) * 1e-3) / 1e12) for r in rec]
            json.dump(rows, open(os.environ["P2PVG_DUMP_GEMMS"], "w"))
        pk = peaks()
        ach = fl / (tms * 1e-3) / 1e12
        roof = dict(bound="tensor",
                    kernel="conv_gemm_kernel + gemm_tc_kernel: all bf16 tcgen05 launches of the step (persistent, TMA 2-D/4-D staged, "
                           "tcgen05.mma kind::f16, double-buffered TMEM accumulators)",
                    achieved=ach, peak=pk["tflops"], unit="TFLOP/s", frac=ach / pk["tflops"], traffic=None, peak_source=pk["which"],
                    launches_per_step=len(rec), gemm_ms_per_step=tms, executed_gemm_tflop_per_step=fl / 1e12,  # useful MACs only (zeros of block-diagonal weights excluded)
                   
                    step_algorithmic_frac=(B * W_FLOP_PER_SEQ / (ms_step * 1e-3)) / 1e12 / pk["tflops"])

    # ---- CPU baseline (oracle port) on the host cores: bounded sample -------------------------------
    cpu = None
    if rank == 0 and not args.skip_cpu:
        threads = host_threads()
        log(f'cpu baseline on {threads} threads')
        times = cpu_reference_steps(T, args.ref_batch, 2, 1, threads)
        log(f'cpu baseline done: {times}')
        cpu = dict(value=T * args.ref_batch * len(times) / sum(times), unit="frames/s", cores=threads, kind="port",
                   sample=f"2 steps of T={T},B={args.ref_batch} (config C1 shape) of the oracle port, all host threads")

    if rank == 0:
        line = dict(metric=METRIC, value=value, unit="frames/s", n_gpus=world, steps=args.steps, warmup=max(args.warmup, 3),
                    ms_per_step=ms_step, higher_is_better=True, scaling="weak", vs_baseline=None,
                    dtype=("bf16" if args.precision == "bf16" else "f32"), data="synthetic",
                    config=dict(workload=f"mnist dcgan_64 seq_len {T} batch {B} per GPU (BASELINE configs[1])", global_batch=B * world,
                                seq_len=T, skip_prob=0.0, parallelism=f"dp{world}", cuda_graph=bool(use_graph),
                                l2="per-step working set (GBs of activations) far exceeds the 126 MB L2; no flush needed",
                                update_mode="A (reference two-phase update)"),
                    roofline=roof, cpu_baseline=cpu,
                    e2e=dict(value=e2e_val, unit="frames/s", h2d_bytes_per_step=int(x_host.numel() * 4), d2h_bytes_per_step=16),
                    gpu_launches=int(launches), clocks=clocks, losses=[float(v) for v in losses])
        print(json.dumps(line), flush=True)
    if world > 1:
        # tear-down must never hang the launcher: drop the CUDA graphs that hold captured NCCL kernels, give
        # destroy_process_group a bounded time, then leave without running interpreter-exit hooks
        import gc
        import threading
        barrier()
        torch.cuda.synchronize()
        eng._graphs.clear()
        gc.collect()
        sys.stdout.flush()
        sys.stderr.flush()
        t = threading.Thread(target=dist.destroy_process_group, daemon=True)
        t.start()
        t.join(15)
        os._exit(0)


if __name__ == "__main__":
    main()
