#!/usr/bin/env python
"""Representative launches of the tcgen05 GEMM (shapes of the dcgan_64 batch-256 train step) for
`ncu --set full -k regex:gemm_tc_kernel`:  conv forward (K-major), ConvT forward (MN-major B, output-bound),
weight gradient (MN-major A and B, split-K)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from p2pvg_b200._lib import CudaKernels  # noqa: E402

K = CudaKernels("cuda")
bf = torch.bfloat16
shapes = [  # M, N, K, a_mn, b_mn, c dtype
    (491520, 256, 2048, False, False, bf),          # encoder c3 forward
    (1966080, 1024, 128, False, True, bf),          # decoder upc4 forward (col buffer, output-bound)
    (512, 4096, 122880, True, True, torch.float32),  # decoder upc2 weight gradient (split-K)
]
for rep in range(2):
    for M, N, Kd, a_mn, b_mn, cdt in shapes:
        A = torch.randn((Kd, M) if a_mn else (M, Kd), device="cuda", dtype=bf)
        B = torch.randn((Kd, N) if b_mn else (N, Kd), device="cuda", dtype=bf)
        C = torch.empty(M, N, device="cuda", dtype=cdt)
        K.gemm(A, B, C, M, N, Kd, a_mn=a_mn, b_mn=b_mn)
        torch.cuda.synchronize()
        del A, B, C
print("done")
