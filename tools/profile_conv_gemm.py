#!/usr/bin/env python
"""Representative launches of the implicit-GEMM convolution kernel (shapes of the dcgan_64 batch-256 step) for
`ncu --set full -k regex:conv_gemm_kernel`: conv forward c3, ConvT forward upc3, weight gradient c3."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from p2pvg_b200._lib import CudaKernels  # noqa: E402

K = CudaKernels("cuda")
bf = torch.bfloat16
N = 7680
for rep in range(2):
    x = torch.randn(N, 16, 16, 128, device="cuda", dtype=bf)          # big map of c3
    w = torch.randn(256, 16 * 128, device="cuda", dtype=bf) * 0.02
    y = torch.empty(N, 8, 8, 256, device="cuda", dtype=bf)
    K.conv_gemm(0, x, w, y, N, 8, 8, 128, 256)                        # kind 0: conv forward
    xs = torch.randn(N, 8, 8, 256, device="cuda", dtype=bf)
    wt = torch.randn(256, 16 * 128, device="cuda", dtype=bf) * 0.02
    yb = torch.empty(N, 16, 16, 128, device="cuda", dtype=bf)
    K.conv_gemm(2, xs, wt, yb, N, 8, 8, 256, 128)                     # kind 2: ConvT forward
    gw = torch.empty(256, 16 * 128, device="cuda")
    K.conv_gemm(1, xs, x, gw, N, 8, 8, 0, 128, Cm=256)                # kind 1: weight gradient
    torch.cuda.synchronize()
print("done")
