#!/usr/bin/env python
"""CUDA-event time of the graph-replayed train step of the bench workload (quick A/B runs of env switches)."""
import argparse
import os
import sys
import types

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from p2pvg_b200.models import dcgan_64  # noqa: E402
from p2pvg_b200.models.p2p_model import P2PModel  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--seq", type=int, default=30)
args = ap.parse_args()
opt = types.SimpleNamespace(dataset="mnist", backbone_net=dcgan_64, lr=1e-3, beta1=0.9, beta=1e-4, weight_cpc=100.0, weight_align=0.5,
                            skip_prob=0.0, n_past=1, last_frame_skip=False, batch_size=args.batch)
torch.manual_seed(1)
model = P2PModel(args.batch, 1, 128, 10, 256, 1, 1, 2, opt=opt).cuda()
x = torch.rand(args.seq, args.batch, 1, 64, 64, device="cuda")
eng = model.engine(64)
for _ in range(4):
    eng.step(x, use_graph=True, return_device=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.steps):
    out = eng.step(x, use_graph=True, return_device=True)
e1.record()
torch.cuda.synchronize()
print(f"{e0.elapsed_time(e1) / args.steps:.3f} ms/step  losses {out.tolist()}")
