#!/usr/bin/env python
"""Multi-GPU data-parallel check (run under torchrun, one rank per GPU, NCCL): every rank trains on its own batch shard
through the captured step (side lanes + bucketed ncclAllReduce(AVG) inside the CUDA graph) and the replicas must stay
bit-identical; rank 0 also compares the averaged-gradient step with N single-GPU gradient computations.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dp_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import p2p_oracle as O  # noqa: E402  (test infrastructure: initial weights only)
from p2pvg_b200._lib import kernels_for  # noqa: E402
from p2pvg_b200.engine import TrainEngine  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
CFG = dict(g_dim=128, z_dim=10, rnn_size=256, channels=1, image_width=64, predictor_rnn_layers=2, posterior_rnn_layers=1, prior_rnn_layers=1)
T, B = 8, 16
state = O.build_state(CFG, seed=1)
opt = O.default_opt(batch_size=B)


def shard(r, it):
    g = torch.Generator().manual_seed(100 + 17 * it + r)
    return torch.rand(T, B, 1, 64, 64, generator=g), O.draw_eps(T - 1, B, 10, seed=200 + 17 * it + r)


for mode in ("A", "B"):
    eng = TrainEngine(O.clone_state(state), CFG, opt, kernels_for(dev), act_dtype=torch.bfloat16, mode=mode)
    eng.dist = (dist, None, world)
    probs = np.random.RandomState(0).uniform(0, 1, T - 1)   # identical skip mask on every rank
    for it in range(4):   # eager, capture, replay, replay
        x, eps = shard(rank, it)
        losses = eng.step(x.to(dev), probs=probs, eps=eps.to(dev), use_graph=True)
        assert np.all(np.isfinite(losses)), losses
    flat = eng.pool["flat"]
    ref = flat.clone()
    dist.broadcast(ref, 0)
    same = torch.equal(ref, flat)
    ok = torch.tensor([int(same)], device=dev)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"mode {mode}: replicas identical after 4 steps: {bool(ok.item())}; losses {losses}", flush=True)
    assert ok.item() == 1, f"mode {mode}: replicas diverged"
    dist.barrier()
# one exchanged step with lr = 0: the averaged gradient left in the arena equals the hand average of the shards' gradients
eng = TrainEngine(O.clone_state(state), CFG, dict(opt, lr=0.0), kernels_for(dev), act_dtype=torch.bfloat16, mode="B")
eng.dist = (dist, None, world)
x, eps = shard(rank, 0)
eng.step(x.to(dev), probs=probs, eps=eps.to(dev))
mine = eng.pool["grad"].clone()
solo = TrainEngine(O.clone_state(state), CFG, dict(opt, lr=0.0), kernels_for(dev), act_dtype=torch.bfloat16, mode="B")
solo.step(x.to(dev), probs=probs, eps=eps.to(dev))
g = solo.pool["grad"].clone()
dist.all_reduce(g)
g /= world
err = (mine - g).abs().max().item() / (g.abs().max().item() + 1e-30)
if rank == 0:
    print(f"averaged gradient vs hand average of the shards: max rel err {err:.3e}", flush=True)
assert err < 1e-5, err
dist.barrier()
if rank == 0:
    print("DP CHECK OK", flush=True)
torch.cuda.synchronize()
os._exit(0)
