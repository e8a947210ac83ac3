#!/usr/bin/env python
"""Context for the bf16-mode gradient tolerances: how far do the gradients of STOCK PyTorch mixed precision
(torch.autocast(bfloat16) around the reference's forward, cuDNN / cuBLAS) drift from exact fp32 on the same step?

Runs the oracle restatement of P2PModel.forward on the GPU twice from identical weights / inputs / noise -- exact fp32 (TF32
off) and under autocast(bf16) -- and prints the per-tensor cosine statistics of the backward-#1 gradients, in the same format as
tests/test_measured_gpu.py prints them for the hand-written bf16 path.  Test / documentation tool (uses oracle/)."""
import argparse
import copy
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import p2p_oracle as O  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--backbone", default="vgg", choices=["vgg", "dcgan"])
ap.add_argument("--seq", type=int, default=6)
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--channels", type=int, default=3)
args = ap.parse_args()
dev = "cuda"
torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
cfg = dict(g_dim=128, z_dim=10, rnn_size=256, predictor_rnn_layers=2, posterior_rnn_layers=1, prior_rnn_layers=1,
           channels=args.channels, image_width=64)
width = 64
if args.backbone == "vgg":
    cfg.update(backbone="vgg")
    width = "vgg"
T, B = args.seq, args.batch
state0 = {m: {k: v.to(dev) for k, v in sd.items()} for m, sd in O.build_state(cfg, seed=1).items()}
x = torch.rand(T, B, args.channels, 64, 64, generator=torch.Generator().manual_seed(2)).to(dev)
probs = np.random.RandomState(0).uniform(0, 1, T - 1)
eps = O.draw_eps(T - 1, B, 10, seed=3).to(dev)
opt = O.default_opt(batch_size=B)


def run(autocast):
    state = copy.deepcopy(state0)
    adam = {m: O.new_adam_state(state[m]) for m in O.MODULES}
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
        out = O.train_step(state, adam, x, opt, width, eps, probs, mode="A")
    return out


ref, ac = run(False), run(True)
print(f"{args.backbone}_64 C={args.channels} T={T} B={B}: losses fp32 {np.round(ref['losses'], 6)}  autocast(bf16) {np.round(ac['losses'], 6)}")
cos = []
for m in O.MODULES:
    for k, g in ref["grads"][m].items():
        h = ac["grads"][m].get(k)
        if g is None or h is None or g.abs().max() == 0:
            continue
        c = torch.nn.functional.cosine_similarity(g.flatten().double(), h.flatten().double(), dim=0).item()
        cos.append((c, m, k))
cos.sort()
v = np.array([c for c, _, _ in cos])
print(f"gradient cosine autocast(bf16) vs fp32 over {len(v)} tensors: min {v.min():.4f}  10% {np.quantile(v, 0.1):.4f}  median {np.median(v):.4f}  "
      f"90% {np.quantile(v, 0.9):.4f}")
for c, m, k in cos[:8]:
    print(f"   {c:.4f}  {m}.{k}")
