#!/usr/bin/env python
"""One forward + one backward cluster scan at the C2 shape (S=29, B=256, R=256) and the C5 shape (S=59, B=256, R=512), for ncu."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from p2pvg_b200._lib import CudaKernels  # noqa: E402

K = CudaKernels("cuda")
for S, B, R in ((29, 256, 256), (59, 256, 512)):
    dev = "cuda"
    pre = torch.randn(S, B, 4 * R, device=dev) * 0.5
    whh = torch.randn(4 * R, R, device=dev) / R ** 0.5
    bhh = torch.randn(4 * R, device=dev) * 0.1
    gates = torch.empty(S, B, 4 * R, device=dev)
    hs = torch.zeros(S + 1, B, R, device=dev)
    cs = torch.zeros(S + 1, B, R, device=dev)
    dG = torch.empty(S, B, 4 * R, device=dev)
    dh = torch.randn(S, B, R, device=dev)
    ctr = torch.zeros(4, dtype=torch.int32, device=dev)
    K.lstm_scan_fwd(pre, whh, bhh, gates, hs, cs, S, B, R, ctr, tf32=True)
    K.lstm_scan_bwd(dh, whh, gates, cs, dG, S, B, R, ctr, tf32=True)
    torch.cuda.synchronize()
print("done")
