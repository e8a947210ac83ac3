#!/usr/bin/env python
"""Does the stride-2 pixel gather cost more TMA time than an un-strided one?  Times the implicit weight-gradient GEMM with the
4x4/stride-2 gather (kind 1) and with the 3x3/stride-1 gather (kind 4) on the same small map, channels and pixel count; the
FLOPs differ only by the tap count (16 vs 9), so TFLOP/s is directly comparable."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from p2pvg_b200._lib import CudaKernels  # noqa: E402

K = CudaKernels("cuda")
bf = torch.bfloat16
N = 7680
for H, Cm, Cn in ((8, 256, 128), (16, 128, 64), (4, 512, 256)):
    a = torch.randn(N, H, H, Cm, device="cuda", dtype=bf)
    big = torch.randn(N, 2 * H, 2 * H, Cn, device="cuda", dtype=bf)
    same = torch.randn(N, H, H, Cn, device="cuda", dtype=bf)
    for kind, b, taps in ((1, big, 16), (4, same, 9)):
        gw = torch.empty(Cm, taps * Cn, device="cuda")
        ts = []
        for rep in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            K.conv_gemm(kind, a, b, gw, N, H, H, 0, Cn, Cm=Cm)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = min(ts[1:])
        fl = 2.0 * N * H * H * taps * Cm * Cn
        print(f"kind {kind} ({'4x4 stride-2' if kind == 1 else '3x3 stride-1'} gather) H={H} Cm={Cm} Cn={Cn}: {ms:.3f} ms  {fl / ms / 1e9:.0f} TFLOP/s")
