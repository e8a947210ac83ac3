#!/usr/bin/env python
"""CUDA-event time of the graph-replayed train step of the bench workload (quick A/B runs of env switches)."""
import argparse
import os
import sys
import types

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from p2pvg_b200.models import dcgan_64, dcgan_128, h36m_mlp, vgg_64, vgg_128  # noqa: E402
from p2pvg_b200.models.p2p_model import P2PModel  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--seq", type=int, default=30)
ap.add_argument("--backbone", default="dcgan_64", choices=["dcgan_64", "dcgan_128", "vgg_64", "vgg_128", "h36m_mlp"])
ap.add_argument("--channels", type=int, default=1)
ap.add_argument("--rnn", type=int, default=256)
args = ap.parse_args()
net = dict(dcgan_64=dcgan_64, dcgan_128=dcgan_128, vgg_64=vgg_64, vgg_128=vgg_128, h36m_mlp=h36m_mlp)[args.backbone]
pose = args.backbone == "h36m_mlp"
width = 128 if args.backbone.endswith("128") else 64
opt = types.SimpleNamespace(dataset="h36m" if pose else "mnist", backbone_net=net, lr=1e-3, beta1=0.9, beta=1e-4, weight_cpc=100.0,
                            weight_align=0.5, skip_prob=0.0, n_past=1, last_frame_skip=False, batch_size=args.batch)
torch.manual_seed(1)
model = P2PModel(args.batch, args.channels, 128, 10, args.rnn, 1, 1, 2, opt=opt).cuda()
if pose:
    x = 3 * torch.randn(args.seq, args.batch, 17, 3, device="cuda")
else:
    x = torch.rand(args.seq, args.batch, args.channels, width, width, device="cuda")
eng = model.engine(width)
for _ in range(4):
    eng.step(x, use_graph=True, return_device=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.steps):
    out = eng.step(x, use_graph=True, return_device=True)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / args.steps
print(f"{args.backbone} C={args.channels} T={args.seq} B={args.batch} R={args.rnn}: {ms:.3f} ms/step = {args.seq * args.batch / ms * 1e3:,.0f} frames/s  "
      f"peak mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB  losses {[round(v, 5) for v in out.tolist()]}")
